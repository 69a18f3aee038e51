"""GPU parity: the CUDA RCD path (through the C ABI) against the oracle.

Bar (BASELINE.json north_star + SURVEY.md 8c): bit-identical colour channels on every pixel the
reference defines; the "reference-undefined" pixels (functions of memory rcd.c never initialises,
and the alpha of the outer 3-px ring) are masked and counted.
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def run_cuda_rcd(mosaic, filters, pm=(1.0, 1.0, 1.0, 1.0), roi=(0, 0), host=False, sensor_filters=None):
    import torch
    import ansel_b200 as ab
    ab.init()
    h, w = mosaic.shape
    data = ab.demosaic_data(ab.DEMOSAIC_RCD)
    piece = ab.make_piece(w, h, filters=sensor_filters if sensor_filters is not None else filters,
                          roi_x=roi[0], roi_y=roi[1], processed_maximum=pm, data=data, devid=0)
    if host:
        out = np.zeros((h, w, 4), np.float32)
        ab.check(ab.lib().b200_demosaic_process_host(piece, mosaic.ctypes.data, out.ctypes.data))
        return out
    d_in = torch.from_numpy(mosaic).cuda()
    d_out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ab.check(ab.lib().b200_demosaic_process_dev(piece, d_in.data_ptr(), d_out.data_ptr(),
                                                torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


def assert_parity(got, mosaic, filters, pm=(1.0, 1.0, 1.0)):
    want = util.oracle_rcd(mosaic, filters, pm)
    mask = util.oracle_rcd_mask(mosaic, filters, pm)
    defined = (mask & 1) == 0
    bad = (got[..., :3].view(np.int32) != want[..., :3].view(np.int32)).any(axis=2) & defined
    assert not bad.any(), f"{int(bad.sum())} defined pixels differ; first at {np.argwhere(bad)[:5].tolist()}"
    alpha_ok = (mask & 2) == 0
    assert (got[..., 3][alpha_ok] == 0.0).all()
    return int((~defined).sum())


@pytest.mark.parametrize("name", list(util.BAYER))
@pytest.mark.parametrize("size", [(16, 16), (117, 131), (206, 206), (207, 113), (640, 480), (1024, 768)])
def test_rcd_uniform_bit_exact(built, name, size):
    w, h = size
    m = util.frame_uniform(w, h, util.SEEDS[0])
    got = run_cuda_rcd(m, util.BAYER[name])
    assert_parity(got, m, util.BAYER[name])


@pytest.mark.parametrize("seed", util.SEEDS)
def test_rcd_natural_bit_exact(built, seed):
    w, h = 1500, 1000
    m = util.frame_natural(w, h, seed)
    got = run_cuda_rcd(m, util.BAYER["RGGB"])
    assert_parity(got, m, util.BAYER["RGGB"])


@pytest.mark.parametrize("kind", ["zeros", "ones", "impulses", "negative", "tiny"])
def test_rcd_edge_inputs(built, kind):
    m = util.frame_edge(333, 222, kind)
    got = run_cuda_rcd(m, util.BAYER["RGGB"])
    assert_parity(got, m, util.BAYER["RGGB"])


@pytest.mark.parametrize("roi", [(0, 0), (1, 0), (0, 1), (1, 1), (3, 5)])
def test_rcd_roi_phase(built, roi):
    """roi_in.(x,y) shifts the CFA phase (develop/imageop.c:139-142): integer, bit exact."""
    import ansel_b200 as ab
    m = util.frame_uniform(300, 200, 11)
    sensor = util.BAYER["RGGB"]
    shifted = ab.lib().b200_roi_filters(sensor, roi[0], roi[1])
    got = run_cuda_rcd(m, shifted, roi=roi, sensor_filters=sensor)
    assert_parity(got, m, shifted)


def test_rcd_processed_maximum_scaling(built):
    m = util.frame_uniform(320, 240, 3) * 1.7
    pm = (1.3, 1.7, 1.1, 1.0)
    got = run_cuda_rcd(m, util.BAYER["RGGB"], pm=pm)
    assert_parity(got, m, util.BAYER["RGGB"], pm[:3])


def test_rcd_too_small_is_untouched(built):
    """rcd.c:280-284: frames under 16 px are left as found."""
    m = util.frame_uniform(15, 40, 1)
    import torch
    import ansel_b200 as ab
    ab.init()
    data = ab.demosaic_data(ab.DEMOSAIC_RCD)
    piece = ab.make_piece(15, 40, data=data, devid=0)
    d_in = torch.from_numpy(m).cuda()
    d_out = torch.full((40, 15, 4), 7.0, dtype=torch.float32, device="cuda")
    ab.check(ab.lib().b200_demosaic_process_dev(piece, d_in.data_ptr(), d_out.data_ptr(), 0))
    torch.cuda.synchronize()
    assert (d_out.cpu().numpy() == 7.0).all()


def test_rcd_host_entry_matches_device_entry(built):
    m = util.frame_natural(800, 600, 5)
    a = run_cuda_rcd(m, util.BAYER["RGGB"], host=False)
    b = run_cuda_rcd(m, util.BAYER["RGGB"], host=True)
    assert (a.view(np.int32) == b.view(np.int32)).all()


def test_rcd_unsupported_is_loud(built):
    import ansel_b200 as ab
    ab.init()
    m = util.frame_uniform(64, 64, 1)
    data = ab.demosaic_data(1028)                       # an X-Trans method (frequency-domain chroma) on a Bayer frame
    piece = ab.make_piece(64, 64, data=data, devid=0)
    out = np.zeros((64, 64, 4), np.float32)
    rc = ab.lib().b200_demosaic_process_host(piece, m.ctypes.data, out.ctypes.data)
    assert rc == ab.B200_ERR_UNSUPPORTED
    assert (out == 0).all() and b"not built" in ab.lib().b200_last_error()


@pytest.mark.parametrize("size", [util.SIZE_24MP, util.SIZE_45MP])
def test_rcd_full_size_bit_exact(built, size):
    """BASELINE.json sizes: the oracle finishes a 45 MP frame in seconds, so compare directly."""
    w, h = size
    m = util.frame_natural(w, h, util.SEEDS[0])
    got = run_cuda_rcd(m, util.BAYER["RGGB"])
    n_undef = assert_parity(got, m, util.BAYER["RGGB"])
    assert n_undef < 8 * (w + h)
