"""GPU parity: diffuse or sharpen through the C ABI against the oracle, bit for bit on all four lanes.  The oracle
is bit-identical to the reference's own process() cut verbatim from iop/diffuse.c (tests/test_cpu_oracle_pin.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from test_cpu_oracle_pin import _diffuse_cases

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def cuda_diffuse(img, data, host=False, iscale=1.0, roi_scale=1.0):
    import ansel_b200 as ab
    ab.init()
    h, w = img.shape[:2]
    piece = ab.make_piece(w, h, filters=0, channels=4, data=data, devid=0, scale=roi_scale)
    piece.iscale = iscale
    src = np.ascontiguousarray(img)
    if host:
        out = np.full_like(src, -7.0)
        ab.check(ab.lib().b200_diffuse_process_host(C.byref(piece), src.ctypes.data, out.ctypes.data))
        return out
    import torch
    d_in = torch.from_numpy(src).cuda()
    d_out = torch.zeros_like(d_in)
    ab.check(ab.lib().b200_diffuse_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("name", list(_diffuse_cases()))
@pytest.mark.parametrize("size", [(640, 427), (97, 61), (33, 200)])
def test_diffuse_bit_exact(built, name, size):
    import ansel_b200 as ab
    w, h = size
    img = util.hdr_rgba(w, h, 3)                      # scene-referred values incl. NaN, inf, negatives, zeros
    d = ab.diffuse_data(**_diffuse_cases()[name])
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()


@pytest.mark.parametrize("zoom", [0.5, 2.0, 4.0])
def test_diffuse_zoom_changes_scale_count(built, zoom):
    """dt_dev_get_module_scale: preview pipes run fewer wavelet scales with rescaled radii (diffuse.c:1173-1183)."""
    import ansel_b200 as ab
    img = util.hdr_rgba(300, 200, 5)
    d = ab.diffuse_data(**_diffuse_cases()["lens_deblur_soft"])
    assert same_bits(cuda_diffuse(img, d, iscale=zoom), util.oracle_diffuse(img, d, iscale=zoom)).all()


def test_diffuse_golden_and_host_entry(built):
    import ansel_b200 as ab
    g = np.load(os.path.join(util.GOLDEN_DIR, "diffuse.npz"))
    for name in ("sharpen_demosaic_aa", "gradient_sharpen", "inpaint_highlights"):
        d = ab.diffuse_data(**_diffuse_cases()[name])
        assert same_bits(cuda_diffuse(g["img"], d, host=True), g[name]).all()


def test_diffuse_full_iteration_count(built):
    """the stock 32-iteration denoise preset, unshortened, on a small frame"""
    import ansel_b200 as ab
    img = util.hdr_rgba(128, 96, 9)
    d = ab.diffuse_data(**ab.DIFFUSE_PRESETS["denoise_medium"])
    assert d.iterations == 32
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()


def test_diffuse_large_radius_many_scales(built):
    """radius 512 -> 10 scales (MAX_NUM_SCALES), dilation up to 512 px, wider than the frame"""
    import ansel_b200 as ab
    img = util.rgba_scene(700, 500, 2)
    d = ab.diffuse_data(iterations=1, radius=512, regularization=2.5, anisotropy_first=2.0, anisotropy_third=2.0, first=-0.2, second=0.1,
                        third=-0.2, fourth=0.1)
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()


def test_diffuse_full_width_strip_bit_exact(built):
    """a strip as wide as the 45 MP bench frame (8256 px)"""
    import ansel_b200 as ab
    img = util.rgba_scene(util.SIZE_45MP[0], 600, 12)
    d = ab.diffuse_data(**_diffuse_cases()["sharpen_demosaic_aa"])
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()


def test_diffuse_12mp_matches_oracle(built):
    import ansel_b200 as ab
    img = util.rgba_scene(4000, 3000, util.SEEDS[0])
    d = ab.diffuse_data(**_diffuse_cases()["sharpen_demosaic_aa"])
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()


def test_diffuse_45mp_properties(built):
    """BASELINE size: deterministic, finite, non-negative (dt_simd_max_zero), and zero speeds with zero
    sharpness reconstruct max(0, HF+LF) = the clipped input up to float rounding of the wavelet split."""
    import torch
    import ansel_b200 as ab
    w, h = util.SIZE_45MP
    img = util.rgba_scene(w, h, util.SEEDS[2])
    d = ab.diffuse_data(**_diffuse_cases()["sharpen_demosaic_aa"])
    a = cuda_diffuse(img, d)
    b = cuda_diffuse(img, d)
    assert same_bits(a, b).all() and np.isfinite(a).all() and (a >= 0).all()
    ident = cuda_diffuse(img, ab.diffuse_data())
    assert np.abs(ident - np.maximum(img, 0)).max() < 1e-5
    torch.cuda.empty_cache()


def test_luminance_mask_inpainting(built):
    """threshold > 0: mask, Box-Muller noise start image (device glibc logf/sinf/cosf), masked PDE; full preset count"""
    import ansel_b200 as ab
    d = ab.diffuse_data(**ab.DIFFUSE_PRESETS["inpaint_highlights"])
    assert d.iterations == 32 and d.threshold > 0
    img = (util.rgba_scene(200, 150, 4) * 2.5).astype(np.float32)
    assert (img[..., :3] > d.threshold).any(-1).mean() > 0.02
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()
    # a frame wide enough that the seed's float index passes 2^24, and one with every pixel masked
    img = (util.rgba_scene(3000, 1500, 6) * 2.5).astype(np.float32)
    d = ab.diffuse_data(**_diffuse_cases()["masked_all_orders"])
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()
    img = np.full((40, 50, 4), 3.0, np.float32)
    assert same_bits(cuda_diffuse(img, d), util.oracle_diffuse(img, d)).all()
