"""GPU: the C module adapters (dt_iop_<op>__process / process_cl / tiling_callback, ansel_b200/iop/modules.c) and
the device-resident pipe glue give exactly what the C-ABI entry points give -- whose parity with the oracle the
per-module suites establish.  Scene-referred chain in the reference's iop order: demosaic -> denoise (profiled)
-> colorin -> filmic -> colorout (src/common/iop_order.c v30 list)."""
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


def _filmic_blob():
    return np.load(os.path.join(util.GOLDEN_DIR, "filmic_data.npz"))["default_v8"]


def _chain(w, h):
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    enc = util.srgb_encode_lut()
    conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=util.fit_unbounded_coeffs(enc))
    blob = np.ascontiguousarray(_filmic_blob(), np.uint8)
    fdata = (C.c_uint8 * blob.size).from_buffer_copy(blob.tobytes())
    datas = dict(demosaic=ab.demosaic_data(ab.DEMOSAIC_RCD), denoiseprofile=ab.denoiseprofile_data(ab.DENOISE_WAVELETS),
                 colorin=ab.colorin_data(conv_in), filmicrgb=fdata, colorout=ab.colorout_data(conv_out))
    pieces = {op: ds.make_piece_iop(op, w, h, d, channels_in=1 if op == "demosaic" else 4, channels_out=4,
                                    filters=util.BAYER["RGGB"] if op == "demosaic" else 0) for op, d in datas.items()}
    pipe = ds.make_pipe(devid=0, work_profile=ds.profile_info(*WORK), output_profile=ds.profile_info(*EXPORT))
    return datas, pieces, pipe, (conv_in, conv_out)


def test_each_adapter_equals_the_abi_entry(built):
    import torch
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    M, L = ds.modlib(), ab.lib()
    w, h = 1204, 806
    datas, pieces, pipe, _keep = _chain(w, h)
    mosaic = util.frame_natural(w, h, 3)
    rgba = util.hdr_rgba(w, h, 4)
    s = torch.cuda.current_stream().cuda_stream
    for op in datas:
        src = mosaic if op == "demosaic" else rgba
        # direct ABI call
        piece = ab.make_piece(w, h, filters=util.BAYER["RGGB"] if op == "demosaic" else 0, channels=1 if op == "demosaic" else 4, devid=0)
        if op == "filmicrgb":
            fp = ab.filmic_piece(_filmic_blob(), WORK, EXPORT)
            piece.data, piece.data_size = C.addressof(fp), C.sizeof(fp)
        else:
            piece.data, piece.data_size = C.cast(C.pointer(datas[op]), C.c_void_p), C.sizeof(datas[op])
        d_in = torch.from_numpy(src).cuda()
        want = torch.zeros((h, w, 4), device="cuda")
        ab.check(getattr(L, f"b200_{op}_process_dev")(C.byref(piece), d_in.data_ptr(), want.data_ptr(), s))
        # process_cl slot: TRUE on success
        got = torch.zeros((h, w, 4), device="cuda")
        assert getattr(M, f"dt_iop_{op}__process_cl")(pieces[op].module, C.byref(pipe), C.byref(pieces[op]), d_in.data_ptr(), got.data_ptr()) == 1
        torch.cuda.synchronize()
        assert same_bits(got.cpu().numpy(), want.cpu().numpy()).all(), op
        # process slot: host buffers, 0 on success
        out = np.zeros((h, w, 4), np.float32)
        assert getattr(M, f"dt_iop_{op}__process")(pieces[op].module, C.byref(pipe), C.byref(pieces[op]), src.ctypes.data, out.ctypes.data) == 0
        assert same_bits(out, want.cpu().numpy()).all(), op
        # tiling_callback agrees with the ABI's
        t0, t1 = ab.Tiling(), ab.Tiling()
        getattr(M, f"dt_iop_{op}__tiling_callback")(pieces[op].module, C.byref(pipe), C.byref(pieces[op]), C.byref(t0))
        getattr(L, f"b200_{op}_tiling")(C.byref(piece), C.byref(t1))
        assert bytes(t0) == bytes(t1), op


def test_bilat_adapter(built):
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    M = ds.modlib()
    w, h = 640, 427
    img = util.lab_scene(w, h, 2)
    data = ab.bilat_data(detail=0.6)
    piece = ds.make_piece_iop("bilat", w, h, data, channels_in=4, channels_out=4)
    pipe = ds.make_pipe(devid=0)
    out = np.zeros_like(img)
    assert M.dt_iop_bilat__process(piece.module, C.byref(pipe), C.byref(piece), img.ctypes.data, out.ctypes.data) == 0
    want = util.oracle_local_laplacian(img, clarity=0.6)
    assert same_bits(out[..., :3], want[..., :3]).all()


def test_filmic_adapter_without_work_profile_fails(built):
    """filmicrgb.c:2716-2717: no work profile -> process() reports an error instead of guessing."""
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    M = ds.modlib()
    _, pieces, _, _keep = _chain(64, 48)
    pipe = ds.make_pipe(devid=0)
    img = util.hdr_rgba(64, 48, 1)
    out = np.zeros_like(img)
    assert M.dt_iop_filmicrgb__process(pieces["filmicrgb"].module, C.byref(pipe), C.byref(pieces["filmicrgb"]), img.ctypes.data, out.ctypes.data) != 0


def test_scene_referred_chain_device_resident(built):
    """Five modules back to back through b200_pixelpipe_process_on_gpu == the same five called one by one."""
    import torch
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    M = ds.modlib()
    w, h = 2004, 1336
    datas, pieces, pipe, _keep = _chain(w, h)
    order = ("demosaic", "denoiseprofile", "colorin", "filmicrgb", "colorout")
    mosaic = util.frame_natural(w, h, 5)
    nodes = (ds.PipeNode * len(order))()
    for k, op in enumerate(order):
        nodes[k].process_cl = C.cast(getattr(M, f"dt_iop_{op}__process_cl"), C.c_void_p)
        nodes[k].module = pieces[op].module
        nodes[k].piece = C.pointer(pieces[op])
    bufs = M.b200_pipe_buffers_new()
    out = np.zeros((h, w, 4), np.float32)
    assert M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, len(order), bufs, mosaic.ctypes.data, out.ctypes.data) == 0
    M.b200_pipe_buffers_free(bufs)

    cur = torch.from_numpy(mosaic).cuda()
    for op in order:
        nxt = torch.zeros((h, w, 4), device="cuda")
        assert getattr(M, f"dt_iop_{op}__process_cl")(pieces[op].module, C.byref(pipe), C.byref(pieces[op]), cur.data_ptr(), nxt.data_ptr()) == 1
        torch.cuda.synchronize()
        cur = nxt
    want = cur.cpu().numpy()
    assert same_bits(out, want).all()
    assert np.isfinite(out[..., :3]).all() and out[..., :3].min() >= 0.0 and out[..., :3].max() <= 1.001


def test_frames_in_flight_equal_one_at_a_time(built):
    """b200_pixelpipe_submit/_wait with 1, 2 and 3 slots: five different frames, same bits as the synchronous call."""
    import torch
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    ab.init()
    M = ds.modlib()
    w, h = 1204, 806
    datas, pieces, pipe, _keep = _chain(w, h)
    order = ("demosaic", "denoiseprofile", "colorin", "filmicrgb", "colorout")   # denoise uses the shared device scratch
    nodes = (ds.PipeNode * len(order))()
    for k, op in enumerate(order):
        nodes[k].process_cl = C.cast(getattr(M, f"dt_iop_{op}__process_cl"), C.c_void_p)
        nodes[k].module = pieces[op].module
        nodes[k].piece = C.pointer(pieces[op])
    frames = [torch.from_numpy(util.frame_natural(w, h, 20 + i)).pin_memory() for i in range(5)]
    bufs = M.b200_pipe_buffers_new()
    want = []
    for f in frames:
        out = np.zeros((h, w, 4), np.float32)
        assert M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, len(order), bufs, f.data_ptr(), out.ctypes.data) == 0
        want.append(out)
    M.b200_pipe_buffers_free(bufs)
    for depth in (1, 2, 3):
        q = M.b200_pipe_queue_new(depth)
        assert q
        outs = [torch.zeros((h, w, 4), dtype=torch.float32).pin_memory() for _ in frames]
        tickets = [M.b200_pixelpipe_submit(q, C.byref(pipe), nodes, len(order), f.data_ptr(), o.data_ptr()) for f, o in zip(frames, outs)]
        assert tickets == list(range(5))
        for t in tickets:
            assert M.b200_pixelpipe_wait(q, t) == 0
        M.b200_pipe_queue_free(q)
        for o, wnt in zip(outs, want):
            assert same_bits(o.numpy(), wnt).all(), depth
    assert M.b200_pipe_queue_new(0) is None and M.b200_pipe_queue_new(9) is None
