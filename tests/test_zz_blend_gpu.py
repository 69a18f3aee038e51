"""B200: blending in the scene-referred RGB space through the C ABI (device and host entry points) and behind a module in the device-resident
pipe glue, bit for bit against the oracle (itself pinned to develop/blend.c and develop/blends/blendif_rgb_jzczhz.c compiled in place)."""
import ctypes as C

import numpy as np
import pytest

import util
import blend_util as bu

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    import ansel_b200 as ab
    ab.init()
    L = ab.lib()
    L.b200_blend_process_dev.restype = C.c_int
    L.b200_blend_process_dev.argtypes = [C.c_void_p, C.POINTER(bu.BlendParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.b200_blend_process_host.restype = C.c_int
    L.b200_blend_process_host.argtypes = [C.c_void_p, C.POINTER(bu.BlendParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return ab


def cuda(ab, a, b, p, form=None, xoffs=0, yoffs=0, host=False):
    import torch
    ih, iw = a.shape[:2]
    oh, ow = b.shape[:2]
    piece = ab.make_piece(iw, ih, channels=4, devid=0)
    piece.roi_out.x, piece.roi_out.y, piece.roi_out.width, piece.roi_out.height = xoffs, yoffs, ow, oh
    if host:
        out, mask = b.copy(), np.full((oh, ow), -3.0, np.float32)
        rc = ab.lib().b200_blend_process_host(C.byref(piece), C.byref(p), a.ctypes.data, out.ctypes.data, None if form is None else form.ctypes.data, mask.ctypes.data)
        return rc, out, mask
    d_a, d_b = torch.from_numpy(a).cuda(), torch.from_numpy(b.copy()).cuda()
    d_f = None if form is None else torch.from_numpy(form).cuda()
    d_m = torch.full((oh, ow), -3.0, device="cuda")
    rc = ab.lib().b200_blend_process_dev(C.byref(piece), C.byref(p), d_a.data_ptr(), d_b.data_ptr(), None if d_f is None else d_f.data_ptr(), d_m.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return rc, d_b.cpu().numpy(), d_m.cpu().numpy()


@pytest.mark.parametrize("cfg", bu.CONFIGS, ids=[c[0] for c in bu.CONFIGS])
def test_blend_bit_exact(built, cfg):
    name, kw, uses_form = cfg
    a, b, form = bu.frames(301, 177, 3)
    p = bu.params(**kw)
    rc_o, out_o, mask_o = bu.oracle(a, b, p, form if uses_form else None)
    rc, out, mask = cuda(built, a, b, p, form if uses_form else None)
    assert rc == 0 and rc_o == 0
    assert same_bits(out, out_o).all() and same_bits(mask, mask_o).all()
    rc, out_h, mask_h = cuda(built, a, b, p, form if uses_form else None, host=True)
    assert rc == 0 and same_bits(out_h, out_o).all() and same_bits(mask_h, mask_o).all()


@pytest.mark.parametrize("cfg", bu.golden_configs(), ids=[c[0] for c in bu.golden_configs()])
def test_committed_reference_output(built, cfg):
    """tests/golden/blend.npz: what the reference's own lines produced (tests/golden/make_golden_blend.py), both colour spaces"""
    import os
    name, kw, uses_form = cfg
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "blend.npz"))
    a, b, form = bu.golden_frames(kw)
    rc, out, mask = cuda(built, a, b, bu.params(**kw), form if uses_form else None)
    assert rc == 0 and same_bits(out, g[name]).all() and same_bits(mask, g[name + "_mask"]).all()


def test_blend_inside_roi_in_and_what_is_refused(built):
    ab = built
    a, b, form = bu.frames(100, 80, 2, xoffs=7, yoffs=5)
    p = bu.params(mask_mode=bu.MASK_ENABLED | bu.MASK_SHAPE | bu.MASK_PARAMETRIC, drawn=1, channels={0: (0.1, 0.3, 0.7, 0.9), 6: (0.0, 0.0, 0.6, 0.9)})
    want = bu.oracle(a, b, p, form, 7, 5)
    got = cuda(ab, a, b, p, form, 7, 5)
    assert got[0] == 0 and same_bits(got[1], want[1]).all() and same_bits(got[2], want[2]).all()
    a, b, form = bu.frames(64, 48, 4)
    for kw in (dict(feathering_radius=5.0), dict(blur_radius=2.0), dict(details=0.5), dict(blend_cst=0), dict(profile_nonlinear=1)):
        rc, out, _ = cuda(ab, a, b, bu.params(**kw))
        assert rc == ab.B200_ERR_UNSUPPORTED and np.array_equal(out, b), kw


@pytest.mark.parametrize("cfg", bu.LAB_CONFIGS, ids=[c[0] for c in bu.LAB_CONFIGS])
def test_lab_blend_bit_exact(built, cfg):
    """the Lab space (develop/blends/blendif_lab.c): the operators that stay in Lab and the L / a / b channels of the parametric mask are built,
    what goes through LCh is refused with the output untouched"""
    name, kw, uses_form = cfg
    a, b, form = bu.frames_lab(301, 177, 3)
    p = bu.params(**kw)
    rc, out, mask = cuda(built, a, b, p, form if uses_form else None)
    if not bu.lab_on_device(cfg):
        assert rc == built.B200_ERR_UNSUPPORTED and np.array_equal(out, b)
        return
    rc_o, out_o, mask_o = bu.oracle(a, b, p, form if uses_form else None)
    assert rc == 0 and rc_o == 0
    assert same_bits(out, out_o).all() and same_bits(mask, mask_o).all()


@pytest.mark.parametrize("cfg", bu.DISPLAY_CONFIGS, ids=[c[0] for c in bu.DISPLAY_CONFIGS])
def test_display_blend_bit_exact(built, cfg):
    """the display-referred RGB space (develop/blends/blendif_rgb_hsl.c): its 27 operators, the H / S / L channels of the parametric mask"""
    name, kw, uses_form = cfg
    a, b, form = bu.frames_display(301, 177, 3)
    p = bu.params(**kw)
    rc, out, mask = cuda(built, a, b, p, form if uses_form else None)
    rc_o, out_o, mask_o = bu.oracle(a, b, p, form if uses_form else None)
    assert rc == 0 and rc_o == 0
    assert same_bits(out, out_o).all() and same_bits(mask, mask_o).all()


@pytest.mark.parametrize("cfg", bu.RAW_CONFIGS, ids=[c[0] for c in bu.RAW_CONFIGS])
def test_raw_blend_bit_exact(built, cfg):
    """the raw space (develop/blends/blendif_raw.c): buffers of one float per site, device and host entry points"""
    name, kw, uses_form = cfg
    a, b, form = bu.frames_raw(301, 177, 3)
    p = bu.params(**kw)
    rc_o, out_o, mask_o = bu.oracle(a, b, p, form if uses_form else None)
    for host in (False, True):
        rc, out, mask = cuda(built, a, b, p, form if uses_form else None, host=host)
        assert rc == 0 and rc_o == 0 and same_bits(out, out_o).all() and same_bits(mask, mask_o).all(), host


def test_lab_blend_on_a_large_frame(built):
    """45 MP in Lab, the overlay operator under a parametric mask on L and b: a strip against the oracle, the rest deterministic"""
    w, h = util.SIZE_45MP
    rng = np.random.default_rng(9)
    a = rng.random((h, w, 4), dtype=np.float32) * np.array([100.0, 160.0, 160.0, 1.0], np.float32) - np.array([0.0, 80.0, 80.0, 0.0], np.float32)
    b = rng.random((h, w, 4), dtype=np.float32) * np.array([100.0, 160.0, 160.0, 1.0], np.float32) - np.array([0.0, 80.0, 80.0, 0.0], np.float32)
    p = bu.params(cst=bu.CS_LAB, mode="overlay", opacity=80.0, mask_mode=bu.MASK_ENABLED | bu.MASK_PARAMETRIC,
                  channels={0: (0.1, 0.3, 0.7, 0.9), 6: (0.3, 0.4, 0.6, 0.8)})
    rc, out, _ = cuda(built, a, b, p)
    rows = slice(3000, 3048)
    want = bu.oracle(np.ascontiguousarray(a[rows]), np.ascontiguousarray(b[rows]), p)
    assert rc == 0 and same_bits(out[rows], want[1]).all() and not np.array_equal(out[rows, :, :3], b[rows, :, :3])
    assert same_bits(out, cuda(built, a, b, p)[1]).all()


def test_blend_45mp_and_linearity_of_the_normal_operator(built):
    """full frame: equal to the oracle on a strip; with the normal operator and a uniform mask the blend of (a, b) at opacity q is
    a * (1 - q) + b * q exactly as float32 computes it"""
    ab = built
    w, h = util.SIZE_45MP
    rng = np.random.default_rng(5)
    a = rng.random((h, w, 4), dtype=np.float32) * 2.0
    b = rng.random((h, w, 4), dtype=np.float32) * 2.0
    p = bu.params(opacity=35.0)
    rc, out, mask = cuda(ab, a, b, p)
    assert rc == 0
    q = np.float32(np.float32(35.0) / np.float32(100.0))
    want = a[..., :3] * (np.float32(1.0) - q) + b[..., :3] * q
    assert same_bits(out[..., :3], want).all() and (out[..., 3] == q).all() and (mask == q).all()
    rows = slice(2000, 2064)
    p2 = bu.params(mode="harmonic_mean", mask_mode=bu.MASK_ENABLED | bu.MASK_PARAMETRIC, channels={0: (0.1, 0.3, 0.7, 0.9), 5: (0.0, 0.2, 0.9, 1.2)})
    rc, out2, _ = cuda(ab, a, b, p2)
    want2 = bu.oracle(np.ascontiguousarray(a[rows]), np.ascontiguousarray(b[rows]), p2)
    assert rc == 0 and same_bits(out2[rows], want2[1]).all()


def test_blend_behind_a_module_in_the_device_resident_pipe(built):
    """exposure with a drawn + parametric mask at 60 % opacity through b200_pixelpipe_process_on_gpu: the module's process_cl adapter, then the
    blend on the device, against the oracles of both"""
    import torch
    from ansel_b200 import dtsurface as ds
    ab = built
    w, h = 640, 400
    a, _, form = bu.frames(w, h, 6)
    M = ds.modlib()
    d_exp = ab.exposure_data(black=0.002, exposure_ev=0.7)
    piece = ds.make_piece_iop("exposure", w, h, d_exp, channels_in=4, channels_out=4)
    p = bu.params(opacity=60.0, mask_mode=bu.MASK_ENABLED | bu.MASK_SHAPE | bu.MASK_PARAMETRIC, drawn=1, channels={0: (0.05, 0.2, 0.8, 1.0)})
    d_form = torch.from_numpy(form).cuda()
    nodes = (ds.PipeNode * 1)()
    nodes[0].process_cl = C.cast(M.dt_iop_exposure__process_cl, C.c_void_p)
    nodes[0].module = piece.module
    nodes[0].piece = C.pointer(piece)
    nodes[0].blend = C.cast(C.pointer(p), C.c_void_p)
    nodes[0].d_form_mask = d_form.data_ptr()
    pipe = ds.make_pipe(devid=0, stream=None)
    bufs = M.b200_pipe_buffers_new()
    out = np.empty((h, w, 4), np.float32)
    assert M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, 1, bufs, a.ctypes.data, out.ctypes.data) == 0
    # the module alone, through the same adapter: what the blend sees as the module's output
    d_a = torch.from_numpy(a).cuda()
    d_b = torch.zeros((h, w, 4), device="cuda")
    assert M.dt_iop_exposure__process_cl(piece.module, C.byref(pipe), C.byref(piece), d_a.data_ptr(), d_b.data_ptr()) == 1
    torch.cuda.synchronize()
    M.b200_pipe_buffers_free(bufs)
    b = d_b.cpu().numpy()
    assert not np.array_equal(b, a)
    want = bu.oracle(a, b, p, form)
    assert same_bits(out, want[1]).all()
    # without blend parameters the node is the module alone
    nodes[0].blend = None
    bufs = M.b200_pipe_buffers_new()
    assert M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, 1, bufs, a.ctypes.data, out.ctypes.data) == 0
    M.b200_pipe_buffers_free(bufs)
    assert same_bits(out, b).all()
