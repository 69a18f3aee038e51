"""GPU parity: colorin / colorout (matrix + tone curves) against the oracle, bit for bit, in both
rounding flavours (reference release-build contraction, and strict C semantics)."""
import ctypes as C

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


def cuda_convert(rgba, conv, op="colorout", mask_display=0, host=False, inplace=False):
    import torch
    import ansel_b200 as ab
    ab.init()
    h, w = rgba.shape[:2]
    data = ab.colorin_data(conv) if op == "colorin" else ab.colorout_data(conv)
    piece = ab.make_piece(w, h, filters=0, channels=4, data=data, devid=0)
    piece.mask_display = mask_display
    L = ab.lib()
    if host:
        out = np.zeros_like(rgba)
        ab.check(getattr(L, f"b200_{op}_process_host")(piece, rgba.ctypes.data, out.ctypes.data))
        return out
    d_in = torch.from_numpy(np.ascontiguousarray(rgba)).cuda()
    d_out = d_in if inplace else torch.zeros_like(d_in)
    ab.check(getattr(L, f"b200_{op}_process_dev")(piece, d_in.data_ptr(), d_out.data_ptr(),
                                                 torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def cases():
    enc, dec = util.srgb_encode_lut(), util.srgb_decode_lut()
    co_t, co_s = util.fit_unbounded_coeffs(enc), util.fit_unbounded_coeffs(dec)
    lin = enc.copy()
    lin[1, 0] = -1.0  # a channel marked linear passes through (conversion.c:652-663)
    return {
        "colorin_matrix": dict(matrix=util.MATRIX_CAM_TO_REC2020),
        "colorin_clip": dict(matrix=util.MATRIX_CAM_TO_REC2020, clip=util.MATRIX_CLIP_IN),
        "colorin_source_curves": dict(matrix=util.MATRIX_CAM_TO_REC2020, lut_s=dec, co_s=co_s),
        "colorout_trc": dict(matrix=util.MATRIX_REC2020_TO_SRGB, lut_t=enc, co_t=co_t),
        "colorout_trc_one_linear": dict(matrix=util.MATRIX_REC2020_TO_SRGB, lut_t=lin, co_t=util.fit_unbounded_coeffs(lin)),
        "both_curves_clip": dict(matrix=util.MATRIX_REC2020_TO_SRGB, clip=util.MATRIX_CLIP_IN, lut_s=dec, co_s=co_s,
                                 lut_t=enc, co_t=co_t),
    }


@pytest.mark.parametrize("fp", [util.FP_CONTRACT, util.FP_STRICT])
@pytest.mark.parametrize("name", list(cases()))
def test_conversion_bit_exact(built, name, fp):
    import ansel_b200 as ab
    kw = cases()[name]
    img = util.rgba_test_image(777, 431, 3)
    img[5, 5, :3] = np.nan
    img[6, 6, :3] = (np.inf, -np.inf, 0.0)
    conv = ab.make_conversion(kw["matrix"], clip_matrix=kw.get("clip"), lut_source=kw.get("lut_s"),
                              coeffs_source=kw.get("co_s"), lut_target=kw.get("lut_t"), coeffs_target=kw.get("co_t"),
                              fp_mode=ab.FP_CONTRACT if fp == util.FP_CONTRACT else ab.FP_STRICT)
    got = cuda_convert(img, conv, op="colorin" if name.startswith("colorin") else "colorout")
    want = util.oracle_convert(img, fp=fp, **kw)
    bad = ~same_bits(got, want)
    assert not bad.any(), f"{int(bad.sum())} floats differ, first {np.argwhere(bad)[:4].tolist()}"


def test_conversion_inplace_host_and_alpha(built):
    import ansel_b200 as ab
    kw = cases()["colorout_trc"]
    img = util.rgba_test_image(640, 333, 5)
    conv = ab.make_conversion(kw["matrix"], lut_target=kw["lut_t"], coeffs_target=kw["co_t"], identity=77)
    a = cuda_convert(img, conv)
    b = cuda_convert(img, conv, inplace=True)
    c = cuda_convert(img, conv, host=True)
    assert same_bits(a, b).all() and same_bits(a, c).all()
    # pipe->mask_display & DISPLAY_MASK copies alpha through (colorout.c:386-387)
    d = cuda_convert(img, conv, mask_display=1)
    assert (d[..., 3] == img[..., 3]).all() and same_bits(d[..., :3], a[..., :3]).all()


def test_lab_and_null_conversion_are_copies(built):
    import ansel_b200 as ab
    img = util.rgba_test_image(100, 50, 1)
    conv = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    import torch
    ab.init()
    for data in (ab.colorin_data(conv, type_=ab.COLORSPACE_LAB), ab.colorin_data(None)):
        piece = ab.make_piece(100, 50, filters=0, channels=4, data=data, devid=0)
        d_in = torch.from_numpy(img).cuda()
        d_out = torch.zeros_like(d_in)
        ab.check(ab.lib().b200_colorin_process_dev(piece, d_in.data_ptr(), d_out.data_ptr(), 0))
        torch.cuda.synchronize()
        assert same_bits(d_out.cpu().numpy(), img).all()


def test_unsupported_conversions_are_loud(built):
    import ansel_b200 as ab
    ab.init()
    img = util.rgba_test_image(32, 32, 1)
    out = np.zeros_like(img)
    conv = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    conv.is_matrix = 0  # lcms2 transform
    piece = ab.make_piece(32, 32, filters=0, channels=4, data=ab.colorout_data(conv), devid=0)
    assert ab.lib().b200_colorout_process_host(piece, img.ctypes.data, out.ctypes.data) == ab.B200_ERR_UNSUPPORTED
    d = ab.colorin_data(ab.make_conversion(util.MATRIX_CAM_TO_REC2020))
    d.blue_mapping = 1
    piece = ab.make_piece(32, 32, filters=0, channels=4, data=d, devid=0)
    assert ab.lib().b200_colorin_process_host(piece, img.ctypes.data, out.ctypes.data) == ab.B200_ERR_UNSUPPORTED


def test_fit_unbounded_coeffs_matches_reference_formula(built):
    import ansel_b200 as ab
    enc = util.srgb_encode_lut()
    co = np.zeros((3, 3), np.float32)
    ptrs = (C.c_void_p * 3)(*[enc[k].ctypes.data for k in range(3)])
    n = ab.lib().b200_fit_unbounded_coeffs(ptrs, co.ctypes.data)
    assert n == 3
    want = util.fit_unbounded_coeffs(enc)
    assert np.max(util.ulp_distance(co, want)) <= 1  # numpy's logf vs glibc's: 1 ulp


def test_full_size_chain_matches_oracle_chain(built):
    """C2 at 45 MP through the C module adapters and the device-resident pipe glue vs the oracle chain."""
    import torch
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    w, h = util.SIZE_45MP
    m = util.frame_natural(w, h, util.SEEDS[1])
    enc = util.srgb_encode_lut()
    co_t = util.fit_unbounded_coeffs(enc)
    conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=co_t)
    datas = [ab.demosaic_data(ab.DEMOSAIC_RCD), ab.colorin_data(conv_in), ab.colorout_data(conv_out)]
    M = ds.modlib()
    pipe = ds.make_pipe(devid=0)
    pieces = [ds.make_piece_iop("demosaic", w, h, datas[0], channels_in=1, channels_out=4, filters=util.BAYER["RGGB"]),
              ds.make_piece_iop("colorin", w, h, datas[1], channels_in=4, channels_out=4),
              ds.make_piece_iop("colorout", w, h, datas[2], channels_in=4, channels_out=4)]
    nodes = (ds.PipeNode * 3)()
    for k, op in enumerate(("demosaic", "colorin", "colorout")):
        nodes[k].process_cl = C.cast(getattr(M, f"dt_iop_{op}__process_cl"), C.c_void_p)
        nodes[k].module = pieces[k].module
        nodes[k].piece = C.pointer(pieces[k])
    bufs = M.b200_pipe_buffers_new()
    out = np.zeros((h, w, 4), np.float32)
    assert M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, 3, bufs, m.ctypes.data, out.ctypes.data) == 0
    M.b200_pipe_buffers_free(bufs)

    rgb = util.oracle_rcd(m, util.BAYER["RGGB"])
    mask = util.oracle_rcd_mask(m, util.BAYER["RGGB"])
    rgb = util.oracle_convert(rgb, util.MATRIX_CAM_TO_REC2020, fp=util.FP_CONTRACT)
    want = util.oracle_convert(rgb, util.MATRIX_REC2020_TO_SRGB, lut_t=enc, co_t=co_t, fp=util.FP_CONTRACT)
    defined = (mask & 1) == 0
    bad = (~same_bits(out[..., :3], want[..., :3])).any(axis=2) & defined
    assert not bad.any(), f"{int(bad.sum())} defined pixels differ"
