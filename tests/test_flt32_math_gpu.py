"""GPU: the device libm equals the host's glibc (what the reference's CPU path calls) bit for bit.
Tolerance stated for the record: 0 ulp observed; the contract allows <= 1 ulp."""
import ctypes as C

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu
N = 2_000_000
FP = C.POINTER(C.c_float)


def _bits(rng, n):
    return rng.integers(0, 2 ** 32, n, dtype=np.uint32).view(np.float32)


def _device(fn, x, y=None):
    import torch
    import ansel_b200 as ab
    ab.init()
    dx = torch.from_numpy(x).cuda()
    dy = torch.from_numpy(y).cuda() if y is not None else None
    out = torch.empty_like(dx)
    ab.check(ab.lib().b200_flt32_eval_dev(fn, dx.data_ptr(), dy.data_ptr() if dy is not None else None,
                                          out.data_ptr(), x.size, 0))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _host(name, *xs):
    L = util.oracle()
    b = np.empty_like(xs[0])
    getattr(L, f"sys_{name}_array")(*[x.ctypes.data_as(FP) for x in xs], b.ctypes.data_as(FP), C.c_size_t(b.size))
    return b


def _assert_same(a, b, x):
    same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
    # subnormal inputs/outputs: the device flushes (FTZ/DAZ like the reference's pipe threads); the
    # host libm called from this test thread may not have FTZ set
    tiny = (np.abs(x) < 1.2e-38) | (np.abs(b) < 1.2e-38) | (np.abs(a) < 1.2e-38)
    bad = ~same & ~tiny
    assert not bad.any(), f"{int(bad.sum())} mismatches, e.g. x={x[bad][:3]} dev={a[bad][:3]} host={b[bad][:3]}"


@pytest.mark.parametrize("fn,name,lo,hi", [(0, "expf", -104, 89), (1, "exp2f", -151, 129), (2, "logf", 0, 8), (3, "log2f", 0, 8)])
def test_device_unary(built, fn, name, lo, hi):
    rng = np.random.default_rng(21)
    x = np.ascontiguousarray(np.concatenate([_bits(rng, N), rng.uniform(lo, hi, N).astype(np.float32),
                                             np.array([0.0, -0.0, 1.0, np.inf, -np.inf, np.nan, -1.0], np.float32)]))
    _assert_same(_device(fn, x), _host(name, x), x)


def test_device_powf(built):
    rng = np.random.default_rng(22)
    for x, y in ((_bits(rng, N), _bits(rng, N)),
                 (rng.uniform(0, 20, N).astype(np.float32), rng.uniform(-3, 6, N).astype(np.float32)),
                 (rng.uniform(0.9, 1.1, N).astype(np.float32), rng.uniform(-300, 300, N).astype(np.float32)),
                 (-rng.uniform(0, 20, N).astype(np.float32), rng.integers(-5, 6, N).astype(np.float32)),
                 (rng.uniform(1.0, 64.0, N).astype(np.float32), np.full(N, 1 / 2.4, np.float32))):
        x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
        a, b = _device(4, x, y), _host("powf", x, y)
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        tiny = (np.abs(x) < 1.2e-38) | (np.abs(y) < 1.2e-38) | (np.abs(b) < 1.2e-38) | (np.abs(a) < 1.2e-38)
        bad = ~same & ~tiny
        assert not bad.any(), f"{int(bad.sum())} mismatches, e.g. {x[bad][:3]} ^ {y[bad][:3]}: dev {a[bad][:3]} host {b[bad][:3]}"


@pytest.mark.parametrize("fn,name", [(5, "sinf"), (6, "cosf")])
def test_device_sinf_cosf(built, fn, name):
    """every argument the Box-Muller call of iop/noise_generator.h can produce, plus random |x| < 120"""
    rng = np.random.default_rng(23)
    k = np.arange(1 << 24, dtype=np.float64)
    x = np.ascontiguousarray(np.concatenate([((2.0 * np.pi) * (k * 2.0 ** -24)).astype(np.float32), rng.uniform(-119.9, 119.9, N).astype(np.float32),
                                             np.array([0.0, -0.0, 1e-5, -1e-5, 0.7853981, 0.7853982], np.float32)]))
    _assert_same(_device(fn, x), _host(name, x), x)


def test_device_atanf_atan2f_hypotf(built):
    """what dt_Lab_2_LCH and dt_JzAzBz_2_JzCzhz call: the fdlibm float routines and the one-square-root hypotf of glibc 2.39"""
    from test_cpu_flt32_math import _pairs
    rng = np.random.default_rng(24)
    x = np.ascontiguousarray(np.concatenate([_bits(rng, 2 * N), rng.uniform(-3, 3, N).astype(np.float32),
                                             np.array([0.0, -0.0, 0.4375, 0.6875, 1.1875, 2.4375, np.inf, -np.inf, np.nan], np.float32)]))
    _assert_same(_device(7, x), _host("atanf", x), x)
    for y, x in _pairs(rng):
        y, x = np.ascontiguousarray(y), np.ascontiguousarray(x)
        for fn, name in ((8, "atan2f"), (9, "hypotf")):
            a, b = _device(fn, y, x), _host(name, y, x)
            same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
            # subnormal arguments, quotients and results: the device flushes like the reference's pipe threads, this test's host thread does not
            q = np.abs(y.astype(np.float64)) / np.maximum(np.abs(x.astype(np.float64)), 1e-300)
            tiny = (np.abs(x) < 1.2e-38) | (np.abs(y) < 1.2e-38) | (np.abs(b) < 1.2e-38) | (np.abs(a) < 1.2e-38) | ((q < 1.2e-38) & (q > 0))
            bad = ~same & ~tiny
            assert not bad.any(), f"{name}: {int(bad.sum())} mismatches, e.g. {y[bad][:3]}, {x[bad][:3]}: dev {a[bad][:3]} host {b[bad][:3]}"
