"""GPU parity: rawprepare, temperature, highlights (clip + bypass), the fused raw front, exposure, gamma and the export
conversions through the C ABI against the oracle, bit for bit.  The oracle is pinned to the reference's own lines
(tests/test_cpu_pipe_ends.py) and the kernels already agree with it on the CPU (tests/test_cpu_pipe_ends_emulation.py);
what only this file sees is nvcc's code generation and the launch code.  Sorted last: written after the round's GPU
budget was spent, so these tests have not run on a B200 yet."""
import ctypes as C
import os

import numpy as np
import pytest

import pipe_ends_util as pe
import test_cpu_pipe_ends as cases
import test_cpu_pipe_ends_emulation as fcases
import util

pytestmark = pytest.mark.gpu
same_bits = cases.same_bits


def run_dev(op, piece, src, out_shape, out_dtype=np.float32, fill=-7.0):
    """b200_<op>_process_dev on device copies of src; -> (rc, out)"""
    import torch
    import ansel_b200 as ab
    ab.init()
    piece.devid = 0
    d_in = torch.from_numpy(np.ascontiguousarray(src).view(np.uint8).reshape(-1)).cuda()
    out = np.full(out_shape, fill, out_dtype)
    d_out = torch.from_numpy(out.view(np.uint8).reshape(-1)).cuda()
    rc = getattr(ab.lib(), f"b200_{op}_process_dev")(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rc, d_out.cpu().numpy().view(out_dtype).reshape(out_shape)


def run_host(op, piece, src, out_shape, out_dtype=np.float32, fill=-7.0):
    import ansel_b200 as ab
    ab.init()
    piece.devid = 0
    src = np.ascontiguousarray(src)
    out = np.full(out_shape, fill, out_dtype)
    rc = getattr(ab.lib(), f"b200_{op}_process_host")(C.byref(piece), src.ctypes.data, out.ctypes.data)
    return rc, out


@pytest.mark.parametrize("name", list(cases.RAWPREPARE_CASES))
def test_rawprepare_bit_exact(built, name):
    piece, src, _ = cases.rawprepare_case(name)
    want = pe.oracle_rawprepare(piece, src)
    for run in (run_dev, run_host):
        rc, got = run("rawprepare", piece, src, want.shape)
        assert rc == 0 and same_bits(got, want).all(), run.__name__


def test_rawprepare_45mp_uint16(built):
    """BASELINE's frame size; the oracle takes a second on it"""
    import ansel_b200 as ab
    w, h = util.SIZE_45MP
    d = ab.rawprepare_data(cases.SUB, cases.DIV, 4, 2)
    raw = np.random.default_rng(5).integers(0, 16384, (h + 2, w + 4), dtype=np.uint16)
    piece = pe.rawprepare_piece(w + 4, h + 2, d)
    rc, got = run_dev("rawprepare", piece, raw, (h, w))
    assert rc == 0 and same_bits(got, pe.oracle_rawprepare(piece, raw)).all()


@pytest.mark.parametrize("name", list(cases.TEMPERATURE_CASES))
def test_temperature_bit_exact(built, name):
    piece, img = cases.temperature_case(name)
    want = pe.oracle_temperature(piece, img)
    for run in (run_dev, run_host):
        rc, got = run("temperature", piece, img, want.shape)
        assert rc == 0 and same_bits(got, want).all(), run.__name__


@pytest.mark.parametrize("name", list(cases.HIGHLIGHTS_CASES))
def test_highlights_bit_exact(built, name):
    piece, img = cases.highlights_case(name)
    rc0, want, _ = pe.oracle_highlights(piece, img)
    assert rc0 == 0
    for run in (run_dev, run_host):
        rc, got = run("highlights", piece, img, want.shape)
        assert rc == 0 and same_bits(got, want).all(), run.__name__


@pytest.mark.parametrize("size", [(1037, 613), (2600, 1702)])
@pytest.mark.parametrize("name", ["inpaint_mosaic", "inpaint_mosaic_wb_roi", "inpaint_xtrans_wb"])
def test_highlights_inpaint_larger_frames(built, name, size):
    """frames of several blocks of lines and of ragged 32x32 tiles: the four directions run at once, the row ones on a transposed copy"""
    piece, img = cases.highlights_case(name, size)
    rc0, want, _ = pe.oracle_highlights(piece, img)
    rc, got = run_dev("highlights", piece, img, want.shape)
    assert rc0 == 0 and rc == 0 and same_bits(got, want).all()
    assert (got != img).mean() > 0.01


def test_highlights_refuses_harmonic_transposition_past_the_bypass(built):
    import ansel_b200 as ab
    _, img = cases.highlights_case("clip_mosaic")
    piece = pe.mosaic_piece(img.shape[1], img.shape[0], ab.highlights_data(ab.HIGHLIGHTS_HARMONIC, 1.0))
    rc, got = run_dev("highlights", piece, img, img.shape)
    assert rc == ab.B200_ERR_UNSUPPORTED and b"clipped" in ab.lib().b200_last_error()
    assert (got == -7.0).all()       # nothing written
    piece = pe.mosaic_piece(img.shape[1], img.shape[0], ab.highlights_data(ab.HIGHLIGHTS_LAPLACIAN, 1.0))     # built: tests/test_zz_hl_laplacian_gpu.py
    rc, got = run_dev("highlights", piece, img, img.shape)
    assert rc == 0 and not (got == -7.0).any() and (got != img).any()      # this frame carries a NaN sample: tests/test_zz_hl_laplacian_gpu.py has the parity cases


def test_highlights_count_is_fresh_on_every_call(built):
    """the device counter is reset per call: a clipped frame followed by a clean one must take the bypass"""
    pc, clipped = cases.highlights_case("clip_mosaic")
    pb, clean = cases.highlights_case("clip_mosaic_24_clipped")
    for piece, img in ((pc, clipped), (pb, clean), (pc, clipped), (pb, clean)):
        rc, got = run_dev("highlights", piece, img, img.shape)
        assert rc == 0 and same_bits(got, pe.oracle_highlights(piece, img)[1]).all()


@pytest.mark.parametrize("name", list(fcases.FRONT_CASES))
def test_fused_raw_front_equals_the_three_modules(built, name):
    """b200_rawfront_process_dev against the oracle chain, and against the three entry points one after the other"""
    import torch
    import ansel_b200 as ab
    ab.init()
    pieces, src = fcases.front_case(name)
    for p in pieces:
        if p is not None:
            p.devid = 0
    want, _ = fcases.oracle_front(pieces, src)
    d_in = torch.from_numpy(np.ascontiguousarray(src).view(np.uint8).reshape(-1)).cuda()
    d_out = torch.full(want.shape, -7.0, dtype=torch.float32, device="cuda")
    ptr = [C.byref(p) if p is not None else None for p in pieces]
    st = torch.cuda.current_stream().cuda_stream
    assert ab.lib().b200_rawfront_process_dev(ptr[0], ptr[1], ptr[2], d_in.data_ptr(), d_out.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert same_bits(d_out.cpu().numpy(), want).all()
    frame = run_dev("rawprepare", pieces[0], src, want.shape)[1]
    if pieces[1] is not None:
        frame = run_dev("temperature", pieces[1], frame, want.shape)[1]
    if pieces[2] is not None:
        frame = run_dev("highlights", pieces[2], frame, want.shape)[1]
    assert same_bits(frame, want).all()


def test_fused_raw_front_45mp(built):
    import torch
    import ansel_b200 as ab
    ab.init()
    w, h = util.SIZE_45MP
    raw = pe.sensor_frame(w, h, 3, clipped=500)
    d = ab.rawprepare_data(cases.SUB, cases.DIV)
    rp = pe.rawprepare_piece(w, h, d, devid=0)
    tp = pe.mosaic_piece(w, h, ab.temperature_data(cases.COEFFS), devid=0)
    hp = pe.mosaic_piece(w, h, ab.highlights_data(ab.HIGHLIGHTS_CLIP, 1.0), pm=(cases.COEFFS[0], 1.0, cases.COEFFS[2], 0.0), devid=0)
    want, n = fcases.oracle_front([rp, tp, hp], raw)
    assert n >= 25
    d_in = torch.from_numpy(raw).cuda()
    d_out = torch.zeros((h, w), dtype=torch.float32, device="cuda")
    assert ab.lib().b200_rawfront_process_dev(C.byref(rp), C.byref(tp), C.byref(hp), d_in.data_ptr(), d_out.data_ptr(),
                                              torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert same_bits(d_out.cpu().numpy(), want).all()


@pytest.mark.parametrize("name", list(cases.EXPOSURE_CASES))
def test_exposure_bit_exact(built, name):
    piece, img = cases.exposure_case(name)
    want = pe.oracle_exposure(piece, img)
    for run in (run_dev, run_host):
        rc, got = run("exposure", piece, img, want.shape)
        assert rc == 0 and same_bits(got, want).all(), run.__name__


def test_gamma_and_export_conversions_bit_exact(built):
    import torch
    import ansel_b200 as ab
    ab.init()
    for img in (pe.awkward_rgba(141, 67, 12), pe.awkward_rgba(1999, 1201, 13)):
        h, w = img.shape[:2]
        piece = ab.make_piece(w, h, filters=0, channels=4)
        want = pe.oracle_gamma(img)
        for run in (run_dev, run_host):
            rc, got = run("gamma", piece, img, img.shape, np.uint8, 0x5A)
            assert rc == 0 and (got == want).all() and (got[..., 3] == 0x5A).all(), run.__name__
        d_in = torch.from_numpy(img).cuda()
        for fmt in (ab.EXPORT_UINT8, ab.EXPORT_UINT8_SWAP, ab.EXPORT_UINT16):
            want = pe.oracle_export(img, fmt)
            d_out = torch.zeros(want.nbytes, dtype=torch.uint8, device="cuda")
            assert ab.lib().b200_export_convert_dev(d_in.data_ptr(), d_out.data_ptr(), w, h, fmt, torch.cuda.current_stream().cuda_stream) == 0
            torch.cuda.synchronize()
            assert (d_out.cpu().numpy().view(want.dtype).reshape(want.shape) == want).all(), fmt
            host = np.zeros_like(want)
            assert ab.lib().b200_export_convert_host(img.ctypes.data, host.ctypes.data, w, h, fmt) == 0
            assert (host == want).all(), fmt


def test_golden_vectors(built):
    """the committed outputs of the reference builds, straight against the CUDA path"""
    g = np.load(os.path.join(util.GOLDEN_DIR, "pipe_ends.npz"))
    for name in cases.RAWPREPARE_CASES:
        piece, src, _ = cases.rawprepare_case(name)
        assert same_bits(run_dev("rawprepare", piece, src, g["rawprepare_" + name].shape)[1], g["rawprepare_" + name]).all(), name
    for name in cases.TEMPERATURE_CASES:
        piece, img = cases.temperature_case(name)
        assert same_bits(run_dev("temperature", piece, img, img.shape)[1], g["temperature_" + name]).all(), name
    for name in cases.HIGHLIGHTS_CASES:
        piece, img = cases.highlights_case(name)
        assert same_bits(run_dev("highlights", piece, img, img.shape)[1], g["highlights_" + name]).all(), name
    for name in cases.EXPOSURE_CASES:
        piece, img = cases.exposure_case(name)
        assert same_bits(run_dev("exposure", piece, img, img.shape)[1], g["exposure_" + name]).all(), name


def test_adapters_and_device_resident_export_chain(built):
    """uint16 sensor data in, uint8 display pixels out through the C module adapters and b200_pixelpipe_process_on_gpu:
    rawprepare -> temperature -> highlights -> demosaic (RCD) -> exposure -> gamma, against the oracle chain"""
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    M = ds.modlib()
    w, h = 640, 480
    raw = pe.sensor_frame(w, h, 21, clipped=80)
    wb = cases.COEFFS
    datas = dict(rawprepare=ab.rawprepare_data(cases.SUB, cases.DIV), temperature=ab.temperature_data(wb),
                 highlights=ab.highlights_data(ab.HIGHLIGHTS_CLIP, 1.0), demosaic=ab.demosaic_data(ab.DEMOSAIC_RCD),
                 exposure=ab.exposure_data(0.0, 0.5), gamma=None)
    pm_wb = (wb[0], wb[1], wb[2], 0.0)
    spec = [("rawprepare", 1, 1, 2, 1, (1.0, 1.0, 1.0, 1.0)), ("temperature", 1, 1, 1, 1, (1.0, 1.0, 1.0, 1.0)), ("highlights", 1, 1, 1, 1, pm_wb),
            ("demosaic", 1, 4, 1, 1, pm_wb), ("exposure", 4, 4, 1, 1, pm_wb), ("gamma", 4, 4, 1, 3, pm_wb)]
    pipe = ds.make_pipe(devid=0)
    pieces = [ds.make_piece_iop(op, w, h, datas[op], channels_in=ci, channels_out=co, filters=util.BAYER["RGGB"], processed_maximum=pm, wb=wb,
                                type_in=ti, type_out=to) for op, ci, co, ti, to, pm in spec]
    nodes = (ds.PipeNode * len(spec))()
    for k, (op, *_r) in enumerate(spec):
        nodes[k].process_cl = C.cast(getattr(M, f"dt_iop_{op}__process_cl"), C.c_void_p)
        nodes[k].module = pieces[k].module
        nodes[k].piece = C.pointer(pieces[k])
    out = np.full((h, w, 4), 0x5A, np.uint8)
    unfused = np.full((h, w, 4), 0x5A, np.uint8)
    fusion = C.c_int.in_dll(M, "b200_pipe_fusion_enabled")       # the raw front runs as one launch by default
    bufs = M.b200_pipe_buffers_new()
    try:
        assert fusion.value == 1
        assert M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, len(spec), bufs, raw.ctypes.data, out.ctypes.data) == 0
        fusion.value = 0
        assert M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, len(spec), bufs, raw.ctypes.data, unfused.ctypes.data) == 0
    finally:
        fusion.value = 1
        M.b200_pipe_buffers_free(bufs)
    assert (out[..., :3] == unfused[..., :3]).all()              # module by module: the same bytes
    # the oracle chain on the same data
    rp = pe.rawprepare_piece(w, h, datas["rawprepare"])
    tp = pe.mosaic_piece(w, h, datas["temperature"])
    hp = pe.mosaic_piece(w, h, datas["highlights"], pm=pm_wb)
    mosaic, _ = fcases.oracle_front([rp, tp, hp], raw)
    rgb = util.oracle_rcd(mosaic, util.BAYER["RGGB"], pm=pm_wb[:3])
    defined = (util.oracle_rcd_mask(mosaic, util.BAYER["RGGB"], pm=pm_wb[:3]) & 1) == 0
    ep = pe.mosaic_piece(w, h, datas["exposure"], filters=0, channels=4)
    want = pe.oracle_gamma(pe.oracle_exposure(ep, rgb), fill=0x5A)
    assert (got_rows := (out[..., :3] == want[..., :3]).all(axis=2))[defined].all(), int((~got_rows & defined).sum())


# ---- finalscale ---------------------------------------------------------------------------------------------------------------
def finalscale_piece(in_w, in_h, in_scale, out_w, out_h, out_scale, itor):
    import ansel_b200 as ab
    p = ab.make_piece(in_w, in_h, filters=0, channels=4, data=ab.finalscale_data(itor), out_width=out_w, out_height=out_h, devid=0)
    p.roi_in.scale, p.roi_out.scale = in_scale, out_scale
    p.roi_in.x, p.roi_in.y, p.roi_out.x, p.roi_out.y = 11, 5, 3, 2      # process() zeroes the origins: they must not matter
    return p


@pytest.mark.parametrize("name", list(cases.FINALSCALE_CASES))
def test_finalscale_bit_exact(built, name):
    img, ow, oh, si, so, itor = cases.finalscale_case(name)
    want = pe.oracle_finalscale(img, ow, oh, si, so, itor)
    piece = finalscale_piece(img.shape[1], img.shape[0], si, ow, oh, so, itor)
    for run in (run_dev, run_host):
        rc, got = run("finalscale", piece, img, want.shape)
        assert rc == 0 and same_bits(got, want).all(), run.__name__
    g = np.load(os.path.join(util.GOLDEN_DIR, "pipe_ends.npz"))
    assert same_bits(got, g["finalscale_" + name]).all()


def test_finalscale_plan_equals_oracle(built):
    """the host code the product runs (b200_resampling_plan) against the oracle's plan"""
    import ansel_b200 as ab
    for itor in (ab.INTERPOLATION_BILINEAR, ab.INTERPOLATION_BICUBIC, ab.INTERPOLATION_MITCHELL):
        for scale in (0.5, 0.3333, 0.06, 1.7, 3.3):
            n_out = max(int(400 * scale) - 7, 4)
            n, l, k, i = pe._plan(ab.lib(), "b200_resampling_plan", itor, 400, 13, n_out, 7, scale)
            on, ol, ok, oi = pe.oracle_plan(itor, 400, 13, n_out, 7, scale)
            assert n == on and (l == ol).all() and (i == oi).all() and same_bits(k, ok).all()


def test_finalscale_export_sizes(built):
    """a 12 MP frame down to 2048 px wide, and the 45 MP frame to half size through a size-independent property: a constant
    image stays that constant wherever the taps are normalised (every output pixel, both axes)"""
    import torch
    import ansel_b200 as ab
    ab.init()
    img = util.rgba_test_image(4000, 3000, 23, lo=0.0, hi=1.2)
    so = 2048 / 4000
    ow, oh = 2048, int(round(3000 * so))
    piece = finalscale_piece(4000, 3000, 1.0, ow, oh, so, ab.INTERPOLATION_MITCHELL)
    rc, got = run_dev("finalscale", piece, img, (oh, ow, 4))
    assert rc == 0 and same_bits(got, pe.oracle_finalscale(img, ow, oh, 1.0, so, ab.INTERPOLATION_MITCHELL)).all()
    w, h = util.SIZE_45MP
    d_in = torch.full((h, w, 4), 0.5, dtype=torch.float32, device="cuda")
    d_out = torch.zeros((h // 2, w // 2, 4), dtype=torch.float32, device="cuda")
    piece = finalscale_piece(w, h, 1.0, w // 2, h // 2, 0.5, ab.INTERPOLATION_BICUBIC)
    assert ab.lib().b200_finalscale_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream) == 0
    torch.cuda.synchronize()
    assert float((d_out - 0.5).abs().max()) < 1e-6


# ---- colour calibration (channelmixerrgb) ---------------------------------------------------------------------------------------
def channelmixer_gpu(img, cp, host=False):
    import ansel_b200 as ab
    piece = ab.make_piece(img.shape[1], img.shape[0], filters=0, channels=4, devid=0)
    piece.data, piece.data_size = C.addressof(cp), C.sizeof(cp)
    return (run_host if host else run_dev)("channelmixerrgb", piece, img, img.shape)


@pytest.mark.parametrize("name", list(cases.CHANNELMIXER_CASES))
def test_channelmixerrgb_bit_exact(built, name):
    img, cp = cases.channelmixer_case(name)
    want = pe.oracle_channelmixerrgb(img, cp)
    for host in (False, True):
        rc, got = channelmixer_gpu(img, cp, host)
        assert rc == 0 and same_bits(got, want).all(), host
    g = np.load(os.path.join(util.GOLDEN_DIR, "pipe_ends.npz"))
    assert same_bits(got, g["channelmixerrgb_" + name]).all()


def test_channelmixerrgb_every_branch_combination_and_a_large_frame(built):
    import ansel_b200 as ab
    img = cases.channelmixer_case("cat16_v3_default")[0]
    for ad in range(5):
        for ver in range(3):
            for clip in (0, 1):
                for grey in (0, 1):
                    cp = ab.channelmixer_piece(cases.WORK, adaptation=ad, version=ver, clip=clip, apply_grey=grey, illuminant=(0.93, 1.02, 0.71), mix=cases.MIX,
                                               saturation=(0.1, -0.2, 0.05), lightness=(0.05, 0.1, -0.1), grey=(0.3, 0.5, 0.2), p=0.85, gamut=1.5)
                    rc, got = channelmixer_gpu(img, cp)
                    assert rc == 0 and same_bits(got, pe.oracle_channelmixerrgb(img, cp)).all(), (ad, ver, clip, grey)
    big = util.hdr_rgba(4000, 3000, 9)
    cp = ab.channelmixer_piece(cases.WORK, adaptation=ab.ADAPTATION_CAT16, illuminant=(0.93, 1.02, 0.71), mix=cases.MIX, saturation=(0.1, -0.2, 0.05))
    rc, got = channelmixer_gpu(big, cp)
    assert rc == 0 and same_bits(got, pe.oracle_channelmixerrgb(big, cp)).all()


def test_channelmixerrgb_adapter(built):
    """dt_iop_channelmixerrgb__process with the work profile on the pipe; without one the adapter refuses"""
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    M = ds.modlib()
    img, cp = cases.channelmixer_case("cat16_v3_tuned")
    h, w = img.shape[:2]
    piece = ds.make_piece_iop("channelmixerrgb", w, h, cp, channels_in=4, channels_out=4)
    piece.data_size = 192                                   # the reference's own data block is what piece->data holds
    out = np.zeros_like(img)
    pipe = ds.make_pipe(devid=0, work_profile=ds.profile_info(*cases.WORK))
    assert M.dt_iop_channelmixerrgb__process(piece.module, C.byref(pipe), C.byref(piece), img.ctypes.data, out.ctypes.data) == 0
    assert same_bits(out, pe.oracle_channelmixerrgb(img, cp)).all()
    assert M.dt_iop_channelmixerrgb__process(piece.module, C.byref(ds.make_pipe(devid=0)), C.byref(piece), img.ctypes.data, out.ctypes.data) != 0


# ---- initialscale and flip ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(cases.INITIALSCALE_CASES))
def test_initialscale_bit_exact(built, name):
    import ansel_b200 as ab
    img, roi_in, roi_out, itor = cases.initialscale_case(name)
    want = pe.oracle_clip_and_zoom(img, roi_in, roi_out, itor)
    p = ab.make_piece(roi_in[2], roi_in[3], filters=0, channels=4, data=ab.finalscale_data(itor), out_width=roi_out[2], out_height=roi_out[3], devid=0)
    p.roi_in.x, p.roi_in.y, p.roi_in.scale = roi_in[0], roi_in[1], roi_in[4]
    p.roi_out.x, p.roi_out.y, p.roi_out.scale = roi_out[0], roi_out[1], roi_out[4]
    for run in (run_dev, run_host):
        rc, got = run("initialscale", p, img, want.shape)
        assert rc == 0 and same_bits(got, want).all(), run.__name__


@pytest.mark.parametrize("orientation", range(8))
def test_flip_bit_exact(built, orientation):
    import ansel_b200 as ab
    for img, ch in ((util.rgba_test_image(1237, 823, 3), 4), (util.frame_natural(641, 419, 3), 1)):
        want = pe.oracle_flip(img, orientation)
        d = ab.FlipData(orientation)
        p = ab.make_piece(img.shape[1], img.shape[0], filters=0, channels=ch, data=d, devid=0)
        for run in (run_dev, run_host):
            rc, got = run("flip", p, img, want.shape)
            assert rc == 0 and same_bits(got, want).all(), (run.__name__, ch)


def test_basebuffer_upload_is_the_crop(built):
    """basebuffer.c:119-160: the crop of the full sensor buffer, as one strided upload (uint16 and float RGBA)"""
    import torch
    import ansel_b200 as ab
    ab.init()
    for full, bpp in ((pe.sensor_frame(301, 200, 3), 2), (util.rgba_test_image(301, 200, 3), 16)):
        ih, iw = full.shape[:2]
        for (x, y, w, h) in ((0, 0, iw, ih), (17, 9, 120, 80), (250, 150, 100, 90)):      # the last one runs past the buffer
            piece = ab.make_piece(w, h, filters=0, channels=1, devid=0)
            piece.roi_out.x, piece.roi_out.y = x, y
            d_out = torch.full((h, w * bpp), 0x5A, dtype=torch.uint8, device="cuda")
            ab.check(ab.lib().b200_basebuffer_upload_dev(C.byref(piece), full.ctypes.data, iw, ih, bpp, d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
            torch.cuda.synchronize()
            want = np.full((h, w * bpp), 0x5A, np.uint8)
            src = np.ascontiguousarray(full).view(np.uint8).reshape(ih, iw * bpp)
            cw, chh = min(w, iw - x), min(h, ih - y)
            want[:chh, :cw * bpp] = src[y:y + chh, x * bpp:(x + cw) * bpp]
            assert (d_out.cpu().numpy() == want).all(), (bpp, x, y, w, h)
