"""GPU parity of the BASELINE.json / SURVEY.md 8d chains C3 and C4, module after module through the C ABI on device-
resident buffers, against the same chain evaluated by the oracle.

The reference leaves ~0.03 % of RCD's output undefined (uninitialised scratch) and every later module
spreads those pixels (NLM by K+P, diffuse by 2^scales, the local Laplacian over the whole frame), so the oracle
chain is fed the CUDA demosaic output -- itself compared with the oracle on every defined pixel right here.
Wavelet denoise is the one module with a tolerance (<= 2 ULP on < 0.1 % of floats); C3/C4 use the NLM mode."""
import ctypes as C
import os

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
WORK = util.profile_pair(util.REC2020_TO_XYZ_D50)
EXPORT = util.profile_pair(util.SRGB_TO_XYZ_D50)


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


class Dev:
    """device-resident module calls"""

    def __init__(self, w, h):
        import torch
        import ansel_b200 as ab
        ab.init()
        self.torch, self.ab, self.w, self.h = torch, ab, w, h
        self.s = torch.cuda.current_stream().cuda_stream

    def run(self, op, data, src, channels_in=4, filters=0):
        ab, torch = self.ab, self.torch
        piece = ab.make_piece(self.w, self.h, filters=filters, channels=channels_in, devid=0)
        piece.data, piece.data_size = C.addressof(data), C.sizeof(data)
        dst = torch.zeros((self.h, self.w, 4), device="cuda")
        ab.check(getattr(ab.lib(), f"b200_{op}_process_dev")(C.byref(piece), src.data_ptr(), dst.data_ptr(), self.s))
        return dst

    def glue(self, src, cst_from, cst_to):
        ab = self.ab
        pm = ab.profile_matrices(*WORK)
        f = ab.lib().b200_colorspace_transform_dev
        f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ab.ProfileMatrices), C.c_int, C.c_void_p]
        ab.check(f(src.data_ptr(), src.data_ptr(), self.w, self.h, cst_from, cst_to, C.byref(pm), 0, self.s))
        return src


def _oracle_nlm_module(img, data):
    f = util.oracle().orc_denoiseprofile_nlmeans
    f.restype = C.c_int
    out = np.zeros_like(img)
    h, w = img.shape[:2]
    assert f(util.fptr(np.ascontiguousarray(img)), util.fptr(out), w, h, C.byref(data), C.c_float(1.0), 1, (C.c_float * 4)(2.0, 1.0, 1.5, 0.0),
             (C.c_float * 4)(1.0, 1.0, 1.0, 1.0)) == 0
    return out


def _front(w, h, seed, K):
    """demosaic -> denoiseprofile(NLM) -> colorin, CUDA and oracle side by side"""
    import torch
    import ansel_b200 as ab
    dev = Dev(w, h)
    mosaic = util.frame_natural(w, h, seed)
    conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    dn = ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=K)   # P=1, scattering 0, strength 1, cpw 0.1
    g_dem = dev.run("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), torch.from_numpy(mosaic).cuda(), channels_in=1, filters=util.BAYER["RGGB"])
    dem = g_dem.cpu().numpy()
    o_dem = util.oracle_rcd(mosaic, util.BAYER["RGGB"])
    defined = (util.oracle_rcd_mask(mosaic, util.BAYER["RGGB"]) & 1) == 0
    assert (same_bits(dem[..., :3], o_dem[..., :3]).all(axis=2) | ~defined).all()
    g = dev.run("colorin", ab.colorin_data(conv_in), dev.run("denoiseprofile", dn, g_dem))
    o = util.oracle_convert(_oracle_nlm_module(dem, dn), util.MATRIX_CAM_TO_REC2020, fp=util.FP_CONTRACT)
    assert same_bits(g.cpu().numpy(), o).all(), "demosaic -> denoise -> colorin"
    return dev, g, o, (conv_in,)


def _filmic(dev, g, o):
    import ansel_b200 as ab
    blob = np.load(os.path.join(util.GOLDEN_DIR, "filmic_data.npz"))["default_v8"]       # filmicrgb.c:246-273 defaults
    fp = ab.filmic_piece(blob, WORK, EXPORT)
    return dev.run("filmicrgb", fp, g), util.oracle_filmic_agx(o, blob, WORK, EXPORT)


def test_c3_chain_bit_exact(built):
    """C3: RCD -> denoiseprofile (NLM, P=1, K=7) -> colorin -> filmic rgb (defaults)."""
    dev, g, o, _k = _front(1000, 700, util.SEEDS[0], K=7)
    g, o = _filmic(dev, g, o)
    assert same_bits(g.cpu().numpy(), o).all()


def test_c3_chain_full_width_strip_bit_exact(built):
    """C3 on a strip as wide as the 45 MP bench frame (8256 x 600): RCD's 88 tile columns, the 115 chunk columns of the
    non-local means and the pointwise kernels' row geometry are the frame's own"""
    dev, g, o, _k = _front(util.SIZE_45MP[0], 600, util.SEEDS[2], K=7)
    g, o = _filmic(dev, g, o)
    assert same_bits(g.cpu().numpy(), o).all()


def test_c4_chain_bit_exact(built):
    """C4: RCD -> denoiseprofile -> colorin -> diffuse (stock sharpen preset) -> filmic -> [RGB->Lab] local
    contrast [Lab->RGB] -> colorout (matrix + sRGB curve)."""
    import ansel_b200 as ab
    dev, g, o, _k = _front(900, 620, util.SEEDS[1], K=4)
    dd = ab.diffuse_data(**ab.DIFFUSE_PRESETS["sharpen_demosaic_aa"])
    g, o = dev.run("diffuse", dd, g), util.oracle_diffuse(o, dd)
    assert same_bits(g.cpu().numpy(), o).all(), "diffuse"
    g, o = _filmic(dev, g, o)
    assert same_bits(g.cpu().numpy(), o).all(), "filmic"
    g = dev.glue(g, ab.CS_RGB, ab.CS_LAB)
    o = util.oracle_rgb_to_lab(o, WORK)
    assert same_bits(g.cpu().numpy(), o).all(), "rgb -> lab"
    g = dev.run("bilat", ab.bilat_data(), g)
    ll = util.oracle_local_laplacian(o)
    ll[..., 3] = o[..., 3]
    o = ll
    assert same_bits(g.cpu().numpy(), o).all(), "local laplacian"
    g = dev.glue(g, ab.CS_LAB, ab.CS_RGB)
    o = util.oracle_lab_to_rgb(o, WORK)
    enc = util.srgb_encode_lut()
    co_t = util.fit_unbounded_coeffs(enc)
    conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=co_t)
    g = dev.run("colorout", ab.colorout_data(conv_out), g)
    o = util.oracle_convert(o, util.MATRIX_REC2020_TO_SRGB, lut_t=enc, co_t=co_t, fp=util.FP_CONTRACT)
    out = g.cpu().numpy()
    assert same_bits(out, o).all(), "colorout"
    assert np.isfinite(out[..., :3]).all()
