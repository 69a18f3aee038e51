"""CPU: the CUDA kernels of a module compiled with g++ through tests/emul/cuda_on_cpu.h and run thread by thread, compared
with the oracle bit for bit.  This checks a kernel's arithmetic and indexing where no GPU is available; it cannot see
nvcc's code generation (FTZ, division rewrites) or the launch code -- the `-m gpu` tests do that.  Test infrastructure
only: nothing of the product routes through this."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import util

EMUL = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emul")


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


@pytest.fixture(scope="module")
def emul_filmic():
    so = os.path.join(EMUL, "libemul_filmic_reconstruct.so")
    srcs = [os.path.join(EMUL, "emul_filmic_reconstruct.cpp"), os.path.join(EMUL, "cuda_on_cpu.h"),
            os.path.join(util.ROOT, "ansel_b200", "csrc", "filmic_reconstruct.cu"), os.path.join(util.ROOT, "ansel_b200", "csrc", "flt32_math.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fno-fast-math", "-ffp-contract=off", "-I", EMUL, "-shared", "-fPIC", "-o", so, srcs[0]], check=True)
    return C.CDLL(so)


def _emulated_reconstruct(lib, img, blob, iscale=1.0, roi_scale=1.0, buf=None):
    h, w = img.shape[:2]
    bw, bh = buf if buf else (w, h)
    src, out, mask = np.ascontiguousarray(img), np.zeros_like(img), np.zeros((h, w), np.float32)
    data = util.aligned_empty((832,), np.uint8)
    data[:] = blob
    sc = util.oracle().orc_filmic_reconstruct_scales
    sc.restype = C.c_int
    scales = sc(C.c_float(iscale), C.c_double(roi_scale), int(bw), int(bh))
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    lib.emul_filmic_reconstruct(vp(src), vp(out), vp(mask), w, h, vp(data), C.c_float(iscale), C.c_double(roi_scale), scales)
    return out, mask


@pytest.mark.parametrize("name", ["rgb_only_gaussian", "default_poisson", "two_passes_uniform_v3"])
def test_filmic_reconstruct_kernels_equal_oracle(emul_filmic, name):
    """mask, noise inpainting, both B-spline blurs, detail split, RGB and ratio reconstruction, ratios and their restore:
    the kernels of ansel_b200/csrc/filmic_reconstruct.cu in the order filmic_reconstruct_dev() launches them"""
    g = np.load(os.path.join(util.GOLDEN_DIR, "filmic_reconstruct.npz"))
    blob = g["data_" + name]
    for img, kw in ((g["img"], {}), (util.hdr_rgba(200, 150, 3), {}),
                    ((util.rgba_scene(160, 120, 9) * 2.0).astype(np.float32), dict(iscale=2.0, roi_scale=0.3, buf=(4000, 3000)))):
        rc, frame, mask = util.oracle_filmic_reconstruct(img, blob, **kw)
        assert rc == 1
        got, got_mask = _emulated_reconstruct(emul_filmic, img, blob, **kw)
        assert same_bits(got_mask, mask).all() and same_bits(got, frame).all()
