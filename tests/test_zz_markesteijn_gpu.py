"""B200: Markesteijn's X-Trans demosaicer (one pass, method 1025; three passes, method 1026) through the C ABI and the module adapter, bit for bit against the oracle
(itself pinned to iop/demosaic/markesteijn.c compiled in place) and the golden vectors of the reference's build."""
import ctypes as C
import os

import numpy as np
import pytest

import util
import markesteijn_util as mu
from vng_util import XTRANS

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


def cuda(ab, m, x, y, method=1025, host=False, smoothing=0):
    import torch
    h, w = m.shape
    d = ab.demosaic_data(method)
    d.color_smoothing = smoothing
    piece = ab.make_piece(w, h, filters=9, data=d, devid=0, roi_x=x, roi_y=y)
    for i in range(6):
        for j in range(6):
            piece.xtrans[i][j] = int(XTRANS[i][j])
    if host:
        out = np.full((h, w, 4), -7.0, np.float32)
        ab.check(ab.lib().b200_demosaic_process_host(C.byref(piece), m.ctypes.data, out.ctypes.data))
        return out
    d_in = torch.from_numpy(np.ascontiguousarray(m)).cuda()
    d_out = torch.full((h, w, 4), -7.0, device="cuda")
    ab.check(ab.lib().b200_demosaic_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("name", list(mu.CASES))
def test_markesteijn_bit_exact(built, name, passes):
    m, x, y = mu.case(name)
    method = 1025 if passes == 1 else 1026
    if min(m.shape) <= (12 if passes == 1 else 17):     # the reference mirrors out of its input there: refused
        with pytest.raises(built.B200Error) as e:
            cuda(built, m, x, y, method=method)
        assert e.value.code == built.B200_ERR_UNSUPPORTED
        return
    want = mu.oracle(m, x, y, passes)
    got = cuda(built, m, x, y, method=method)
    assert same_bits(got[..., :3], want[..., :3]).all() and (got[..., 3] == -7.0).all()   # lane 3 is not a result: kept as found
    assert same_bits(cuda(built, m, x, y, method=method, host=True), got).all()
    g = np.load(os.path.join(util.GOLDEN_DIR, "markesteijn.npz"))
    assert same_bits(got[..., :3], g[f"p{passes}_{name}"][..., :3]).all()


def test_markesteijn_larger_frames_second_call_and_a_dark_frame(built):
    for (w, h, x, y, seed) in ((1500, 1100, 0, 0, 3), (1203, 907, 4, 5, 4)):
        m = util.frame_natural(w, h, seed)
        want = mu.oracle(m, x, y, 1)
        for _ in range(2):                       # the second call runs on the cached plan of the frame's geometry
            assert same_bits(cuda(built, m, x, y)[..., :3], want[..., :3]).all(), (w, h)
        assert same_bits(cuda(built, m, x, y, method=1026)[..., :3], mu.oracle(m, x, y, 3)[..., :3]).all(), (w, h)
    m = np.zeros((300, 260), np.float32)         # every maximum of green is 0.0f, the loop's marker of a new pair
    m[70:180, 60:190] = 0.5
    assert same_bits(cuda(built, m, 2, 1)[..., :3], mu.oracle(m, 2, 1, 1)[..., :3]).all()


def test_markesteijn_with_colour_smoothing_and_fdc_refused(built):
    ab = built
    m, x, y = mu.case("roi2")
    want = util.oracle_color_smoothing(mu.oracle(m, x, y, 1), 2)
    got = cuda(ab, m, x, y, smoothing=2)
    assert same_bits(got[..., :3], want[..., :3]).all()
    with pytest.raises(ab.B200Error) as e:
        cuda(ab, m, x, y, method=1028)           # DT_IOP_DEMOSAIC_FDC
    assert e.value.code == ab.B200_ERR_UNSUPPORTED


def test_markesteijn_45mp_properties(built):
    """full frame: a constant mosaic comes back constant (every stage reproduces a flat field), and the kept pixels of a second call agree
    bit for bit with a crop computed on its own wherever the crop's mirrored border is out of reach"""
    import torch
    ab = built
    w, h = util.SIZE_45MP
    flat = cuda(ab, np.full((h, w), 0.375, np.float32), 0, 0)
    assert (flat[..., :3] == 0.375).all()
    m = util.frame_natural(w, h, 9)
    full = cuda(ab, m, 0, 0)
    assert np.isfinite(full[..., :3]).all()
    # a crop whose origin keeps the tile grid (a multiple of 98 rows and columns, and of 6 for the pattern: 294) is the same computation
    # in its interior: tiles away from the crop's border see the same pixels
    y0, x0, ch, cw = 294 * 3, 294 * 5, 98 * 6, 98 * 7
    crop = cuda(ab, np.ascontiguousarray(m[y0:y0 + ch, x0:x0 + cw]), x0 % 6, y0 % 6)
    inner = (slice(98, ch - 98), slice(98, cw - 98))
    assert same_bits(crop[inner][..., :3], full[y0:y0 + ch, x0:x0 + cw][inner][..., :3]).all()
    del torch
