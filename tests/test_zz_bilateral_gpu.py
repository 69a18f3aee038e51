"""GPU parity: local contrast's bilateral-grid mode through b200_bilat_process_* against the oracle (= the reference with one
splat slice), bit for bit.  Sorted last: written after the round's GPU budget was spent, not yet run on a B200."""
import ctypes as C
import os

import numpy as np
import pytest

import bilateral_util as bu
import util

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def cuda_bilateral(img, ss, sr, detail, host=False, iscale=1.0, roi_scale=1.0):
    import torch
    import ansel_b200 as ab
    ab.init()
    h, w = img.shape[:2]
    data = ab.bilat_data(sigma_r=sr, sigma_s=ss, detail=detail, mode=0)
    piece = ab.make_piece(w, h, filters=0, channels=4, data=data, devid=0, scale=roi_scale)
    piece.iscale = iscale
    if host:
        out = np.full_like(img, -7.0)
        rc = ab.lib().b200_bilat_process_host(C.byref(piece), img.ctypes.data, out.ctypes.data)
        return rc, out
    d_in = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    d_out = torch.full(img.shape, -7.0, device="cuda")
    rc = ab.lib().b200_bilat_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rc, d_out.cpu().numpy()


@pytest.mark.parametrize("name", list(bu.CASES))
def test_bilateral_bit_exact(built, name):
    img, ss, sr, detail = bu.case(name)
    want = bu.oracle_bilateral(img, ss, sr, detail)[1]
    for host in (False, True):
        rc, got = cuda_bilateral(img, ss, sr, detail, host)
        assert rc == 0 and same_bits(got, want).all(), host
    g = np.load(os.path.join(util.GOLDEN_DIR, "bilateral.npz"))
    assert same_bits(got, g[name]).all()


def test_bilateral_module_scale_and_a_larger_frame(built):
    """sigma_s is divided by the module scale (iscale / roi scale), bilat.c:343-345"""
    img = np.ascontiguousarray(util.lab_scene(2000, 1300, 3))
    rc, got = cuda_bilateral(img, 24.0, 5.0, 0.5, iscale=1.0, roi_scale=0.5)          # module scale 2 -> sigma_s 12
    assert rc == 0 and same_bits(got, bu.oracle_bilateral(img, 12.0, 5.0, 0.5)[1]).all()


def test_bilateral_refuses_degenerate_grids(built):
    import ansel_b200 as ab
    rc, _ = cuda_bilateral(np.zeros((300, 3, 4), np.float32), 50.0, 5.0, 0.5)     # two grid columns: narrower than the blur's five taps
    assert rc == ab.B200_ERR_UNSUPPORTED
