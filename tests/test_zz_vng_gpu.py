"""GPU parity: VNG4 and the dual demosaic (RCD + VNG4, AMaZE + VNG4) through b200_demosaic_process_* against the oracle,
bit for bit.  The oracle is pinned to the reference's lines (tests/test_cpu_vng.py) and the kernels already agree with it on
the CPU.  Sorted last: written after the round's GPU budget was spent, so these tests have not run on a B200 yet."""
import ctypes as C
import os

import numpy as np
import pytest

import util
import vng_util as vu

pytestmark = pytest.mark.gpu
DUAL = 2048      # DEMOSAIC_DUAL, iop/demosaic.c:109


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def cuda_demosaic(mosaic, filters, method, x=0, y=0, host=False, dual_thrs=0.2, green_eq=0, smoothing=0):
    import torch
    import ansel_b200 as ab
    ab.init()
    h, w = mosaic.shape
    d = ab.demosaic_data(method)
    d.dual_thrs, d.green_eq, d.color_smoothing = dual_thrs, green_eq, smoothing
    piece = ab.make_piece(w, h, filters=filters, data=d, devid=0, roi_x=x, roi_y=y, wb_coeffs=vu.WB)
    if host:
        out = np.zeros((h, w, 4), np.float32)
        ab.check(ab.lib().b200_demosaic_process_host(C.byref(piece), mosaic.ctypes.data, out.ctypes.data))
        return out
    d_in = torch.from_numpy(np.ascontiguousarray(mosaic)).cuda()
    d_out = torch.zeros((h, w, 4), device="cuda")
    ab.check(ab.lib().b200_demosaic_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("name", list(vu.CASES))
def test_vng4_bit_exact(built, name):
    import ansel_b200 as ab
    m, filters, x, y = vu.case(name)
    want = vu.oracle_vng(m, filters, x, y)
    for host in (False, True):
        assert same_bits(cuda_demosaic(m, filters, ab.DEMOSAIC_VNG4, x, y, host), want).all(), host
    g = np.load(os.path.join(util.GOLDEN_DIR, "vng.npz"))
    assert same_bits(cuda_demosaic(m, filters, ab.DEMOSAIC_VNG4, x, y), g["vng_" + name]).all()


def test_vng4_larger_frames_and_green_equilibration(built):
    import ansel_b200 as ab
    for pat, f in util.BAYER.items():
        m = util.frame_natural(1300, 900, 5, filters=f)
        assert same_bits(cuda_demosaic(m, f, ab.DEMOSAIC_VNG4), vu.oracle_vng(m, f)).all(), pat
    f = util.BAYER["RGGB"]
    m = util.frame_natural(900, 600, 12)
    eq = util.oracle_green_eq(m, f, 1)                       # local average in front (demosaic.c:1137-1170), smoothing behind
    want = util.oracle_color_smoothing(vu.oracle_vng(eq, f), 1)
    assert same_bits(cuda_demosaic(m, f, ab.DEMOSAIC_VNG4, green_eq=1, smoothing=1)[..., :3], want[..., :3]).all()


@pytest.mark.parametrize("name", [n for n in vu.CASES if n != "gbrg_wide"])
def test_dual_rcd_vng4_bit_exact(built, name):
    """RCD, then the blend with VNG4 of the module's input under the detail mask (defined RCD pixels only: rcd.c leaves a
    few border sites undefined, tests/test_rcd_gpu.py)"""
    import ansel_b200 as ab
    m, filters, x, y = vu.case(name)
    m = np.nan_to_num(m)
    rf = ab.lib().b200_roi_filters(C.c_uint32(filters), x, y)
    sharp = np.ascontiguousarray(util.oracle_rcd(m, rf), np.float32)
    undefined = (util.oracle_rcd_mask(m, rf) & 1) != 0
    far = np.ones(m.shape, bool)
    if undefined.any():
        # rcd.c leaves those sites to uninitialised scratch: the oracle blend is fed what the device left there, and pixels within
        # reach of one (3x3 Scharr, then the 9x9 blur of the mask) are not compared
        import scipy.ndimage as ndi
        sharp = np.where(undefined[..., None], cuda_demosaic(m, filters, ab.DEMOSAIC_RCD, x, y), sharp)
        far = ~ndi.binary_dilation(undefined, iterations=6)
        assert far.mean() > 0.4          # the 16x16 frame keeps 46 % of its pixels out of reach
    for thr in (0.2, 1.0):
        want = vu.oracle_dual(sharp, m, filters, x, y, thr)
        got = cuda_demosaic(m, filters, ab.DEMOSAIC_RCD | DUAL, x, y, dual_thrs=thr)
        assert same_bits(got[..., :3], want[..., :3])[far].all(), thr     # lane 3 of the sharp frame is the demosaicer's business (alpha)


def test_dual_on_a_frame_without_undefined_rcd_sites_and_with_amaze(built):
    import ansel_b200 as ab
    f = util.BAYER["RGGB"]
    m = util.frame_natural(1000, 700, 4)
    for method, sharp in ((ab.DEMOSAIC_RCD, util.oracle_rcd(m, f)), (ab.DEMOSAIC_AMAZE, util.oracle_amaze(m, f))):
        sharp = np.ascontiguousarray(sharp, np.float32)
        got = cuda_demosaic(m, f, method | DUAL, dual_thrs=0.3)
        if method == ab.DEMOSAIC_RCD:
            defined = (util.oracle_rcd_mask(m, f) & 1) == 0
            sharp = np.where(defined[..., None], sharp, cuda_demosaic(m, f, method))     # undefined sites: whatever the device left there
        want = vu.oracle_dual(sharp, m, f, 0, 0, 0.3)
        far = np.ones(m.shape, bool)
        if method == ab.DEMOSAIC_RCD:                     # an undefined site spreads through the 3x3 Scharr and the 9x9 blur
            import scipy.ndimage as ndi
            far = ~ndi.binary_dilation(~defined, iterations=6)
        assert same_bits(got[..., :3], want[..., :3])[far].all(), method
    assert same_bits(cuda_demosaic(m, f, ab.DEMOSAIC_RCD | DUAL, dual_thrs=0.0)[..., :3], cuda_demosaic(m, f, ab.DEMOSAIC_RCD)[..., :3]).all()   # dual.c:52


def test_dual_refuses_other_pairs(built):
    import ansel_b200 as ab
    with pytest.raises(ab.B200Error) as e:
        cuda_demosaic(np.zeros((64, 64), np.float32), util.BAYER["RGGB"], ab.DEMOSAIC_PPG | DUAL)
    assert e.value.code == ab.B200_ERR_UNSUPPORTED


@pytest.mark.parametrize("name", list(vu.XTRANS_CASES))
def test_vng_xtrans_bit_exact(built, name):
    """X-Trans sensors: method DT_IOP_DEMOSAIC_VNG (1024) through the module; lane 3 is not a result (kept as found)"""
    import torch
    import ansel_b200 as ab
    ab.init()
    m, x, y = vu.xtrans_case(name)
    h, w = m.shape
    d = ab.demosaic_data(1024)
    piece = ab.make_piece(w, h, filters=9, data=d, devid=0, roi_x=x, roi_y=y)
    for i in range(6):
        for j in range(6):
            piece.xtrans[i][j] = int(vu.XTRANS[i][j])
    d_in = torch.from_numpy(np.ascontiguousarray(m)).cuda()
    d_out = torch.full((h, w, 4), -7.0, device="cuda")
    ab.check(ab.lib().b200_demosaic_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = d_out.cpu().numpy()
    assert same_bits(got[..., :3], vu.oracle_vng_xtrans(m, x, y)[..., :3]).all() and (got[..., 3] == -7.0).all()
    host = np.full((h, w, 4), -7.0, np.float32)
    ab.check(ab.lib().b200_demosaic_process_host(C.byref(piece), m.ctypes.data, host.ctypes.data))
    assert same_bits(host, got).all()
    d2 = ab.demosaic_data(1028)                                   # frequency-domain chroma: not built
    piece.data = C.cast(C.pointer(d2), C.c_void_p)
    assert ab.lib().b200_demosaic_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream) == ab.B200_ERR_UNSUPPORTED
