"""GPU parity: RGB <-> Lab glue and the denoise (non-local means) iop through the C ABI against the oracle, bit for
bit; both oracles are bit-identical to the reference functions cut verbatim (tests/test_cpu_oracle_pin.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
WORK = util.profile_pair(util.REC2020_TO_XYZ_D50)


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def cuda_transform(img, cst_from, cst_to, inplace=False, nonlinearlut=0):
    import torch
    import ansel_b200 as ab
    ab.init()
    h, w = img.shape[:2]
    pm = ab.profile_matrices(*WORK)
    d_in = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    d_out = d_in if inplace else torch.zeros_like(d_in)
    f = ab.lib().b200_colorspace_transform_dev
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ab.ProfileMatrices), C.c_int, C.c_void_p]
    rc = f(d_in.data_ptr(), d_out.data_ptr(), w, h, cst_from, cst_to, C.byref(pm), nonlinearlut, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rc, d_out.cpu().numpy()


@pytest.mark.parametrize("inplace", [False, True])
def test_rgb_to_lab_bit_exact(built, inplace):
    import ansel_b200 as ab
    rgb = util.hdr_rgba(1000, 700, 6)
    rc, got = cuda_transform(rgb, ab.CS_RGB, ab.CS_LAB, inplace)
    assert rc == 0 and same_bits(got, util.oracle_rgb_to_lab(rgb, WORK)).all()      # lane 3: the pixel's own alpha


@pytest.mark.parametrize("inplace", [False, True])
def test_lab_to_rgb_bit_exact(built, inplace):
    import ansel_b200 as ab
    lab = util.lab_scene(1000, 700, 6)
    lab[5, 5, :3] = (0.0, 0.0, 0.0)
    lab[6, 6, :3] = (7.9, -300.0, 250.0)
    lab[7, 7, 0] = np.nan
    rc, got = cuda_transform(lab, ab.CS_LAB, ab.CS_RGB, inplace)
    assert rc == 0 and same_bits(got, util.oracle_lab_to_rgb(lab, WORK)).all()


def test_golden_and_dispatch(built):
    import ansel_b200 as ab
    g = np.load(os.path.join(util.GOLDEN_DIR, "labglue.npz"))
    assert same_bits(cuda_transform(g["rgb"], ab.CS_RGB, ab.CS_LAB)[1][..., :3], g["lab_of_rgb"][..., :3]).all()
    assert same_bits(cuda_transform(g["lab"], ab.CS_LAB, ab.CS_RGB)[1], g["rgb_of_lab"]).all()
    rc, same = cuda_transform(g["rgb"], ab.CS_RGB, ab.CS_RGB)                        # iop_profile.c:1305-1309
    assert rc == 0 and same_bits(same, g["rgb"]).all()
    assert cuda_transform(g["rgb"], ab.CS_RGB, ab.CS_LAB, nonlinearlut=1)[0] == ab.B200_ERR_UNSUPPORTED
    assert cuda_transform(g["rgb"], 0, ab.CS_LAB)[0] == ab.B200_ERR_ARG              # "invalid conversion", :594


def cuda_transform_trc(img, cst_from, cst_to, curves, inplace=False):
    import torch
    import ansel_b200 as ab
    ab.init()
    h, w = img.shape[:2]
    pm = ab.profile_matrices(*util.profile_pair(util.SRGB_TO_XYZ_D50))
    d_in = torch.from_numpy(np.ascontiguousarray(img)).cuda()
    d_out = d_in if inplace else d_in.clone()        # a channel without a curve and lane 3 keep the pixel (in-place behaviour)
    f = ab.lib().b200_colorspace_transform_trc_dev
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ab.ProfileMatrices), C.POINTER(ab.ProfileCurves), C.c_void_p]
    rc = f(d_in.data_ptr(), d_out.data_ptr(), w, h, cst_from, cst_to, C.byref(pm), C.byref(curves), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rc, d_out.cpu().numpy()


@pytest.mark.parametrize("partial", [False, True])
@pytest.mark.parametrize("inplace", [False, True])
def test_glue_with_tone_curves_bit_exact(built, partial, inplace):
    """a work profile with tone curves (sRGB): lut_in before the matrix, lut_out after it, per channel"""
    import ansel_b200 as ab
    from test_cpu_oracle_pin import _srgb_curves, SRGB_PROFILE
    d, cd, e, ce = _srgb_curves(partial)
    cv = ab.profile_curves(d, cd, e, ce, identity=0x51 + partial)
    rgb, lab = util.hdr_rgba(1000, 700, 6), util.lab_scene(1000, 700, 6)
    rc, got = cuda_transform_trc(rgb, ab.CS_RGB, ab.CS_LAB, cv, inplace)
    assert rc == 0 and same_bits(got, util.oracle_rgb_to_lab_trc(rgb, SRGB_PROFILE, d, cd)).all()
    rc, got = cuda_transform_trc(lab, ab.CS_LAB, ab.CS_RGB, cv, inplace)
    assert rc == 0 and same_bits(got, util.oracle_lab_to_rgb_trc(lab, SRGB_PROFILE, e, ce)).all()


def test_glue_tone_curves_golden_and_flag(built):
    import ansel_b200 as ab
    from test_cpu_oracle_pin import _srgb_curves, SRGB_PROFILE
    g = np.load(os.path.join(util.GOLDEN_DIR, "labglue.npz"))
    d, cd, e, ce = _srgb_curves(False)
    cv = ab.profile_curves(d, cd, e, ce)
    assert same_bits(cuda_transform_trc(g["rgb"], ab.CS_RGB, ab.CS_LAB, cv)[1][..., :3], g["lab_of_rgb_trc"][..., :3]).all()
    assert same_bits(cuda_transform_trc(g["lab"], ab.CS_LAB, ab.CS_RGB, cv)[1], g["rgb_of_lab_trc"]).all()
    # three linear input curves: the profile counts as linear and lut_out is ignored (iop_profile.c:303-329)
    d[:, 0] = -1.0
    cv = ab.profile_curves(d, util.fit_unbounded_coeffs(d), e, ce)
    rc, got = cuda_transform_trc(g["lab"], ab.CS_LAB, ab.CS_RGB, cv)
    assert rc == 0 and same_bits(got, util.oracle_lab_to_rgb(g["lab"], SRGB_PROFILE)).all()


def test_round_trip_45mp(built):
    """full size: RGB -> Lab -> RGB returns the in-gamut input to ~1e-5 (Halley cube root, one step)."""
    import ansel_b200 as ab
    w, h = util.SIZE_45MP
    rgb = np.abs(util.rgba_scene(w, h, util.SEEDS[1])) + 1e-3
    rc, lab = cuda_transform(rgb, ab.CS_RGB, ab.CS_LAB)
    assert rc == 0
    rc, back = cuda_transform(lab, ab.CS_LAB, ab.CS_RGB, inplace=True)
    assert rc == 0 and np.abs(back[..., :3] - rgb[..., :3]).max() < 2e-4
    assert same_bits(back[..., 3], rgb[..., 3]).all()


def cuda_nlmeans(img, data, roi_scale=1.0, pipe_type=1, mask_display=0, host=False):
    import torch
    import ansel_b200 as ab
    ab.init()
    h, w = img.shape[:2]
    piece = ab.make_piece(w, h, filters=0, channels=4, data=data, devid=0, scale=roi_scale, pipe_type=pipe_type)
    piece.mask_display = mask_display
    src = np.ascontiguousarray(img)
    if host:
        out = np.zeros_like(src)
        ab.check(ab.lib().b200_nlmeans_process_host(C.byref(piece), src.ctypes.data, out.ctypes.data))
        return out
    d_in = torch.from_numpy(src).cuda()
    d_out = torch.zeros_like(d_in)
    ab.check(ab.lib().b200_nlmeans_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


@pytest.mark.parametrize("scale,pipe", [(1.0, 1), (0.5, 1), (2.5, 2), (0.3, 4), (0.4, 3)])
def test_nlmeans_iop_bit_exact(built, scale, pipe):
    import ansel_b200 as ab
    img = util.lab_scene(400, 300, 3)
    for d in (ab.nlmeans_data(), ab.nlmeans_data(radius=1.0, strength=120.0, luma=1.0, chroma=1.0)):
        for mask in (0, 1):
            got = cuda_nlmeans(img, d, scale, pipe, mask, host=(mask == 1))
            want = util.oracle_nlmeans_iop(img, d, scale, 1 if pipe in (3, 4) else 0, mask)
            assert same_bits(got, want).all()


def test_lab_module_inside_an_rgb_pipe(built):
    """the glue as the pipe uses it: work RGB -> Lab, local contrast (a Lab module), Lab -> work RGB, in place."""
    import ansel_b200 as ab
    rgb = np.abs(util.rgba_scene(640, 480, 8)) + 1e-3
    rc, lab = cuda_transform(rgb, ab.CS_RGB, ab.CS_LAB)
    from test_bilat_gpu import cuda_ll
    ll = cuda_ll(lab, host=False)
    rc, back = cuda_transform(ll, ab.CS_LAB, ab.CS_RGB, inplace=True)
    o_lab = util.oracle_rgb_to_lab(rgb, WORK)
    o_ll = util.oracle_local_laplacian(o_lab)
    o_ll[..., 3] = o_lab[..., 3]
    want = util.oracle_lab_to_rgb(o_ll, WORK)
    assert same_bits(back, want).all()
