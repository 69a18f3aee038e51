"""B200: the LMMSE demosaicer (method 6, every refinement mode) through the C ABI and the host entry point, bit for bit against the oracle
with the tile planes zeroed in front of every tile (the mode the kernel reproduces; tests/test_cpu_lmmse.py pins the oracle to the reference)."""
import ctypes as C

import numpy as np
import pytest

import util
import lmmse_util as lu

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import ansel_b200 as ab
    ab.init()
    return ab


def cuda(ab, m, filters, mode, host=False, smoothing=0, green_eq=0):
    import torch
    h, w = m.shape
    d = ab.demosaic_data(6)
    d.lmmse_refine, d.color_smoothing, d.green_eq = mode, smoothing, green_eq
    piece = ab.make_piece(w, h, filters=filters, data=d, devid=0, processed_maximum=lu.PMAX + (1.0,))
    if host:
        out = np.full((h, w, 4), -7.0, np.float32)
        ab.check(ab.lib().b200_demosaic_process_host(C.byref(piece), m.ctypes.data, out.ctypes.data))
        return out
    d_in = torch.from_numpy(np.ascontiguousarray(m)).cuda()
    d_out = torch.full((h, w, 4), -7.0, device="cuda")
    ab.check(ab.lib().b200_demosaic_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("name", list(lu.CASES))
def test_lmmse_bit_exact(built, name, mode):
    m, f = lu.case(name)
    want = lu.oracle(m, f, mode, carry=0)
    got = cuda(built, m, f, mode)
    assert same_bits(got, want).all()
    if mode in (1, 4):
        assert same_bits(cuda(built, m, f, mode, host=True), got).all()


def test_lmmse_larger_frame_colour_smoothing_and_small_frames(built):
    ab = built
    m = util.frame_natural(1500, 1100, 3)
    f = util.BAYER["RGGB"]
    assert same_bits(cuda(ab, m, f, 1), lu.oracle(m, f, 1, carry=0)).all()
    want = util.oracle_color_smoothing(lu.oracle(m, f, 3, carry=0), 1)
    assert same_bits(cuda(ab, m, f, 3, smoothing=1)[..., :3], want[..., :3]).all()
    tiny = util.frame_natural(15, 40, 1)                      # "too small area": the output stays as found (:140-144)
    assert (cuda(ab, tiny, f, 1) == -7.0).all()


def test_lmmse_45mp_properties(built):
    """full frame: a constant mosaic comes back constant away from the frame's border (the reference's planes are zero outside the
    frame, so the outer rows and columns are darker there too: tests/test_cpu_lmmse.py pins them); a crop on the tile grid (a multiple of 112 rows / columns, even) is the same
    computation away from the crop's border"""
    ab = built
    w, h = util.SIZE_45MP
    f = util.BAYER["RGGB"]
    flat = cuda(ab, np.full((h, w), 0.375, np.float32), f, 1)
    v = flat[100, 100, 0]
    assert (flat[8:-8, 8:-8, :3] == v).all() and abs(float(v) - 0.375) < 1e-3
    edge = lu.oracle(np.full((300, 400), 0.375, np.float32), f, 1, carry=0)
    assert same_bits(flat[:8, :380, :3], edge[:8, :380, :3]).all() and same_bits(flat[:280, :8, :3], edge[:280, :8, :3]).all()
    m = util.frame_natural(w, h, 9)
    full = cuda(ab, m, f, 2)
    assert np.isfinite(full[..., :3]).all()
    y0, x0, ch, cw = 112 * 10, 112 * 20, 112 * 5, 112 * 6
    crop = cuda(ab, np.ascontiguousarray(m[y0:y0 + ch, x0:x0 + cw]), f, 2)
    inner = (slice(120, ch - 130), slice(120, cw - 130))
    assert same_bits(crop[inner], full[y0:y0 + ch, x0:x0 + cw][inner]).all()
