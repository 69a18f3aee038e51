"""CPU: the kernels of ansel_b200/csrc/pipe_ends.cu compiled with g++ through tests/emul/cuda_on_cpu.h, run thread by
thread in the order the entry points launch them, and compared with the oracle bit for bit on the cases of
tests/test_cpu_pipe_ends.py.  This checks arithmetic, indexing and the host-side set-up those kernels share with the
product (fill_prepare, make_thresholds); nvcc's code generation and the launch code are what the `-m gpu` tests see."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import ansel_b200 as ab
import pipe_ends_util as pe
import test_cpu_pipe_ends as cases
import util

EMUL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")
same_bits = cases.same_bits


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(EMUL, "libemul_pipe_ends.so")
    srcs = [os.path.join(EMUL, "emul_pipe_ends.cpp"), os.path.join(EMUL, "cuda_on_cpu.h"), os.path.join(util.ROOT, "ansel_b200", "csrc", "pipe_ends.cu"),
            os.path.join(util.ROOT, "include", "b200iop.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fno-fast-math", "-ffp-contract=off", "-I", EMUL, "-shared", "-fPIC", "-o", so, srcs[0]], check=True)
    subprocess.run(["make", "-s", "-C", util.ORACLE_DIR, "oracle"], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("name", list(cases.RAWPREPARE_CASES))
def test_rawprepare_kernels_equal_oracle(emul, name):
    piece, src, _ = cases.rawprepare_case(name)
    want = pe.oracle_rawprepare(piece, src)
    got = np.full_like(want, -7.0)
    assert emul.emul_rawprepare(C.byref(piece), pe.vp(src), pe.vp(got)) == 0
    assert same_bits(got, want).all()


@pytest.mark.parametrize("name", list(cases.TEMPERATURE_CASES))
def test_temperature_kernels_equal_oracle(emul, name):
    piece, img = cases.temperature_case(name)
    want = pe.oracle_temperature(piece, img)
    got = np.full_like(want, -7.0)
    assert emul.emul_temperature(C.byref(piece), pe.vp(img), pe.vp(got)) == 0
    assert same_bits(got, want).all()


@pytest.mark.parametrize("name", list(cases.HIGHLIGHTS_CASES))
def test_highlights_kernels_equal_oracle(emul, name):
    piece, img = cases.highlights_case(name)
    rc, want, n_want = pe.oracle_highlights(piece, img)
    got, n = np.full_like(want, -7.0), C.c_ulonglong(0)
    shifted = ab.lib().b200_roi_filters(C.c_uint32(piece.filters), piece.roi_in.x, piece.roi_in.y)
    assert emul.emul_highlights(C.byref(piece), pe.vp(img), pe.vp(got), C.byref(n), C.c_uint32(shifted)) == 0 and rc == 0
    assert n.value == n_want and same_bits(got, want).all()


@pytest.mark.parametrize("name", ["inpaint_mosaic_wb_roi", "inpaint_xtrans_wb"])
def test_inpaint_kernels_on_a_frame_of_several_blocks(emul, name):
    """more lines than one block of threads holds, ragged 32x32 tiles of the transposition, lines that end inside a group of eight steps"""
    piece, img = cases.highlights_case(name, (301, 267))
    rc, want, n_want = pe.oracle_highlights(piece, img)
    got, n = np.full_like(want, -7.0), C.c_ulonglong(0)
    shifted = ab.lib().b200_roi_filters(C.c_uint32(piece.filters), piece.roi_in.x, piece.roi_in.y)
    assert emul.emul_highlights(C.byref(piece), pe.vp(img), pe.vp(got), C.byref(n), C.c_uint32(shifted)) == 0 and rc == 0
    assert n.value == n_want and same_bits(got, want).all() and (got != img).mean() > 0.01


def test_highlights_kernels_refuse_reconstruction_past_the_bypass(emul):
    _, img = cases.highlights_case("clip_mosaic")
    piece = pe.mosaic_piece(img.shape[1], img.shape[0], ab.highlights_data(ab.HIGHLIGHTS_HARMONIC, 1.0))
    got, n = np.zeros_like(img), C.c_ulonglong(0)
    assert emul.emul_highlights(C.byref(piece), pe.vp(img), pe.vp(got), C.byref(n), C.c_uint32(piece.filters)) == 3 and n.value >= 25


@pytest.mark.parametrize("name", list(cases.EXPOSURE_CASES))
def test_exposure_kernel_equals_oracle(emul, name):
    piece, img = cases.exposure_case(name)
    want = pe.oracle_exposure(piece, img)
    got = np.full_like(want, -7.0)
    assert emul.emul_exposure(C.byref(piece), pe.vp(img), pe.vp(got)) == 0
    assert same_bits(got, want).all()


def test_float_to_integer_kernels_equal_oracle(emul):
    img = pe.awkward_rgba(141, 67, 12)
    npx = C.c_size_t(img.shape[0] * img.shape[1])
    got = np.full(img.shape, 0x5A, np.uint8)
    emul.emul_gamma(pe.vp(img), pe.vp(got), npx)
    assert (got == pe.oracle_gamma(img)).all()
    for fmt in (ab.EXPORT_UINT8, ab.EXPORT_UINT8_SWAP, ab.EXPORT_UINT16):
        got = np.zeros(img.shape, pe.EXPORT_DTYPE[fmt])
        emul.emul_export(pe.vp(img), pe.vp(got), npx, fmt)
        assert (got == pe.oracle_export(img, fmt)).all()


FRONT_CASES = {
    "uint16_all_three": dict(),
    "uint16_crop_gainmaps": dict(x=3, y=5, gain=True),
    "float_all_three": dict(datatype=ab.TYPE_FLOAT, x=2, y=2),
    "uint16_bypass": dict(clipped=7, level=0.3),                        # fewer than 25 blown samples: highlights copies through
    "uint16_no_highlights": dict(highlights=False),
    "uint16_rawprepare_only": dict(highlights=False, temperature=False, x=1, y=1),
    "uint16_no_temperature": dict(temperature=False),
    "uint16_odd_width_roi": dict(size=(131, 77), x=1, y=0, out=(3, 2, 101, 60)),
}


def front_case(name, size=(134, 78)):
    """-> (pieces [rawprepare, temperature or None, highlights or None], sensor data)"""
    kw = dict(FRONT_CASES[name])
    w, h = kw.pop("size", size)
    gain = cases.gain_maps() if kw.pop("gain", False) else None
    d = ab.rawprepare_data(cases.SUB, cases.DIV, kw.pop("x", 0), kw.pop("y", 0), gain=gain, spacing=(1.0 / 8, 1.0 / 6), origin=(0.01, -0.02))
    datatype = kw.pop("datatype", ab.TYPE_UINT16)
    raw = pe.sensor_frame(w, h, 11, clipped=kw.pop("clipped", 60), level=kw.pop("level", 0.9))
    src = raw if datatype == ab.TYPE_UINT16 else raw.astype(np.float32)
    with_t, with_h = kw.pop("temperature", True), kw.pop("highlights", True)
    rp = pe.rawprepare_piece(w, h, d, datatype=datatype, **kw)
    ow, oh, ox, oy = rp.roi_out.width, rp.roi_out.height, rp.roi_out.x, rp.roi_out.y
    tp = pe.mosaic_piece(ow, oh, ab.temperature_data(cases.COEFFS), x=ox, y=oy) if with_t else None
    pm = (cases.COEFFS[0], cases.COEFFS[1], cases.COEFFS[2], 0.0) if with_t else (1.0, 1.0, 1.0, 0.0)   # temperature's commit scales the maximum
    hp = pe.mosaic_piece(ow, oh, ab.highlights_data(ab.HIGHLIGHTS_CLIP, 0.98), x=ox, y=oy, pm=pm) if with_h else None
    return [rp, tp, hp], src


def oracle_front(pieces, src):
    """the three modules one after the other through the oracle"""
    rp, tp, hp = pieces
    frame = pe.oracle_rawprepare(rp, src)
    if tp is not None:
        frame = pe.oracle_temperature(tp, frame)
    n = None
    if hp is not None:
        rc, frame, n = pe.oracle_highlights(hp, frame)
        assert rc == 0
    return frame, n


@pytest.mark.parametrize("name", list(FRONT_CASES))
def test_fused_raw_front_kernels_equal_the_oracle_chain(emul, name):
    pieces, src = front_case(name)
    want, n = oracle_front(pieces, src)
    if pieces[2] is not None:
        assert (n < 25) == ("bypass" in name)
    got = np.full_like(want, -7.0)
    ptr = [C.byref(p) if p is not None else None for p in pieces]
    assert emul.emul_rawfront(ptr[0], ptr[1], ptr[2], pe.vp(src), pe.vp(got)) == 0
    assert same_bits(got, want).all()


# ---- finalscale: ansel_b200/csrc/resample.cu -------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul_resample():
    so = os.path.join(EMUL, "libemul_resample.so")
    srcs = [os.path.join(EMUL, "emul_resample.cpp"), os.path.join(EMUL, "cuda_on_cpu.h"), os.path.join(util.ROOT, "ansel_b200", "csrc", "resample.cu"),
            os.path.join(util.ROOT, "include", "b200iop.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fno-fast-math", "-ffp-contract=off", "-I", EMUL, "-shared", "-fPIC", "-o", so, srcs[0]], check=True)
    subprocess.run(["make", "-s", "-C", util.ORACLE_DIR, "oracle"], check=True)
    return C.CDLL(so)


@pytest.mark.parametrize("itor", [ab.INTERPOLATION_BILINEAR, ab.INTERPOLATION_BICUBIC, ab.INTERPOLATION_MITCHELL])
@pytest.mark.parametrize("scale", [0.5, 0.3333, 0.77, 0.06, 0.999, 1.001, 1.7, 2.0, 3.3])
def test_product_plan_builder_equals_oracle(emul_resample, itor, scale):
    """build_axis_plan (the host code the product runs) against the oracle's plan: lengths, taps bit for bit, indexes"""
    n_in = 400
    for x0_in, x0_out in ((0, 0), (13, 7)):
        n_out = max(int(n_in * scale) - x0_out, 4)
        n, l, k, i = pe._plan(emul_resample, "emul_resampling_plan", itor, n_in, x0_in, n_out, x0_out, scale)
        on, ol, ok, oi = pe.oracle_plan(itor, n_in, x0_in, n_out, x0_out, scale)
        assert n == on and n > 0 and (l == ol).all() and (i == oi).all() and same_bits(k, ok).all()
    assert pe._plan(emul_resample, "emul_resampling_plan", itor, n_in, 0, n_in, 0, 1.0)[0] == -1


@pytest.mark.parametrize("name", list(cases.FINALSCALE_CASES))
def test_finalscale_kernels_equal_oracle(emul_resample, name):
    args = cases.finalscale_case(name)
    assert same_bits(pe._finalscale(emul_resample, "emul_finalscale", *args), pe.oracle_finalscale(*args)).all()


# ---- colour calibration: ansel_b200/csrc/channelmixer.cu -------------------------------------------------------------------
@pytest.fixture(scope="module")
def emul_channelmixer():
    so = os.path.join(EMUL, "libemul_channelmixer.so")
    srcs = [os.path.join(EMUL, "emul_channelmixer.cpp"), os.path.join(EMUL, "cuda_on_cpu.h"), os.path.join(util.ROOT, "ansel_b200", "csrc", "channelmixer.cu"),
            os.path.join(util.ROOT, "ansel_b200", "csrc", "flt32_math.cuh"), os.path.join(util.ROOT, "include", "b200iop.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fno-fast-math", "-ffp-contract=off", "-I", EMUL, "-shared", "-fPIC", "-o", so, srcs[0]], check=True)
    subprocess.run(["make", "-s", "-C", util.ORACLE_DIR, "oracle"], check=True)
    return C.CDLL(so)


def _emul_channelmixer(lib, img, cp):
    out = np.full_like(img, -7.0)
    lib.emul_channelmixerrgb(pe.vp(np.ascontiguousarray(img)), pe.vp(out), img.shape[1], img.shape[0], C.byref(cp))
    return out


@pytest.mark.parametrize("name", list(cases.CHANNELMIXER_CASES))
def test_channelmixer_kernel_equals_oracle(emul_channelmixer, name):
    img, cp = cases.channelmixer_case(name)
    assert same_bits(_emul_channelmixer(emul_channelmixer, img, cp), pe.oracle_channelmixerrgb(img, cp)).all()


def test_channelmixer_kernel_every_branch_combination(emul_channelmixer):
    img = cases.channelmixer_case("cat16_v3_default")[0]
    for ad in range(5):
        for ver in range(3):
            for clip in (0, 1):
                for grey in (0, 1):
                    for gamut in (0.0, 1.5):
                        cp = ab.channelmixer_piece(cases.WORK, adaptation=ad, version=ver, clip=clip, apply_grey=grey, illuminant=(0.93, 1.02, 0.71), mix=cases.MIX,
                                                   saturation=(0.1, -0.2, 0.05), lightness=(0.05, 0.1, -0.1), grey=(0.3, 0.5, 0.2), p=0.85, gamut=gamut)
                        assert same_bits(_emul_channelmixer(emul_channelmixer, img, cp), pe.oracle_channelmixerrgb(img, cp)).all(), (ad, ver, clip, grey, gamut)


def test_x87_operations_equal_the_host_long_double():
    """ansel_b200/csrc/x87.cuh (the three 80-bit operations of the LCh highlight reconstruction in integer arithmetic) against the
    host's own long double on 60 million operands: denormals, signed zeros, near-cancellations included"""
    exe = os.path.join(EMUL, "x87_selftest")
    src = os.path.join(EMUL, "x87_selftest.cpp")
    dep = os.path.join(util.ROOT, "ansel_b200", "csrc", "x87.cuh")
    if not os.path.exists(exe) or max(os.path.getmtime(src), os.path.getmtime(dep)) > os.path.getmtime(exe):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fno-fast-math", "-ffp-contract=off", "-I", EMUL, "-o", exe, src], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "0 mismatches" in r.stdout, r.stdout[-500:]


@pytest.mark.parametrize("name", list(cases.INITIALSCALE_CASES))
def test_initialscale_kernels_equal_oracle(emul_resample, name):
    args = cases.initialscale_case(name)
    assert same_bits(pe._clip_and_zoom(emul_resample, "emul_clip_and_zoom", *args), pe.oracle_clip_and_zoom(*args)).all()


@pytest.mark.parametrize("orientation", range(8))
def test_flip_kernel_equals_oracle(emul_resample, orientation):
    for img in (util.rgba_test_image(37, 23, 3), util.frame_natural(41, 19, 3), util.rgba_test_image(300, 5, 4)):
        assert same_bits(pe._flip(emul_resample, "emul_flip", img, orientation, "ch"), pe.oracle_flip(img, orientation)).all()


def test_inpaint_kernels_on_random_frame_sizes(emul):
    """colour inpainting, Bayer and X-Trans: frame sizes around the groups of eight steps, the 32x32 tiles of the transposition and the 128 lines of
    a block, against the oracle (pinned on the reference's lines by tests/test_cpu_pipe_ends.py)"""
    rng = np.random.default_rng(31)
    sizes = [(9, 9), (16, 8), (33, 31), (64, 64), (65, 129), (130, 47)] + [(int(rng.integers(8, 200)), int(rng.integers(8, 160))) for _ in range(10)]
    for k, size in enumerate(sizes):
        name = ("inpaint_mosaic_wb_roi", "inpaint_xtrans_wb", "inpaint_mosaic")[k % 3]
        piece, img = cases.highlights_case(name, size)
        rc, want, n_want = pe.oracle_highlights(piece, img)
        got, n = np.full_like(want, -7.0), C.c_ulonglong(0)
        shifted = ab.lib().b200_roi_filters(C.c_uint32(piece.filters), piece.roi_in.x, piece.roi_in.y)
        assert emul.emul_highlights(C.byref(piece), pe.vp(img), pe.vp(got), C.byref(n), C.c_uint32(shifted)) == 0 and rc == 0, (name, size)
        assert n.value == n_want and same_bits(got, want).all(), (name, size)
