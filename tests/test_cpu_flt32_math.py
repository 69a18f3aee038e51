"""CPU tests: the libm restatement (oracle/restate/flt32_math.h) equals the system libm -- the one
the reference's CPU path calls -- bit for bit, on random bit patterns and on the ranges the modules use."""
import ctypes as C

import numpy as np
import pytest

import util

FP = C.POINTER(C.c_float)
N = 1_000_000


@pytest.fixture(scope="module", autouse=True)
def _build():
    util.build_oracle()


def _bits(rng, n):
    return rng.integers(0, 2 ** 32, n, dtype=np.uint32).view(np.float32)


def _same(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


def _run(name, *xs):
    L = util.oracle()
    a, b = np.empty_like(xs[0]), np.empty_like(xs[0])
    args = [x.ctypes.data_as(FP) for x in xs]
    getattr(L, f"orc_{name}_array")(*args, a.ctypes.data_as(FP), C.c_size_t(a.size))
    getattr(L, f"sys_{name}_array")(*args, b.ctypes.data_as(FP), C.c_size_t(a.size))
    return a, b


@pytest.mark.parametrize("name,lo,hi", [("expf", -104, 89), ("exp2f", -151, 129), ("logf", 0, 8), ("log2f", 0, 8)])
def test_unary_equals_system_libm(name, lo, hi):
    rng = np.random.default_rng(11)
    x = np.concatenate([_bits(rng, N), rng.uniform(lo, hi, N).astype(np.float32),
                        np.array([0.0, -0.0, 1.0, np.inf, -np.inf, np.nan, 1e-40, -1.0], np.float32)])
    a, b = _run(name, x)
    assert _same(a, b).all()


def test_powf_equals_system_libm():
    rng = np.random.default_rng(12)
    for x, y in ((_bits(rng, N), _bits(rng, N)),
                 (rng.uniform(0, 20, N).astype(np.float32), rng.uniform(-3, 6, N).astype(np.float32)),
                 (rng.uniform(0.9, 1.1, N).astype(np.float32), rng.uniform(-300, 300, N).astype(np.float32)),
                 (-rng.uniform(0, 20, N).astype(np.float32), rng.integers(-5, 6, N).astype(np.float32)),
                 (rng.uniform(1.0, 64.0, N).astype(np.float32), np.full(N, 1 / 2.4, np.float32))):
        a, b = _run("powf", np.ascontiguousarray(x), np.ascontiguousarray(y))
        assert _same(a, b).all()


def _pairs(rng):
    """argument pairs of atan2f / hypotf: random bit patterns, the Lab plane (|a|, |b| <= 130 on a 0.05 grid, grey pixels included), moderate
    exponents, and the special values"""
    grid = lambda: ((rng.integers(0, 5201, N) - 2600) * 0.05).astype(np.float32)   # noqa: E731
    near = lambda: (rng.standard_normal(N) * np.exp2(rng.integers(-8, 8, N))).astype(np.float32)   # noqa: E731
    sp = np.array([0.0, -0.0, 1.0, -1.0, np.inf, -np.inf, np.nan, 1e-30, -1e30, 3.0e38], np.float32)
    yy, xx = np.meshgrid(sp, sp)
    return [(_bits(rng, N), _bits(rng, N)), (grid(), grid()), (near(), near()), (np.ascontiguousarray(yy.ravel()), np.ascontiguousarray(xx.ravel()))]


def test_atanf_equals_system_libm():
    rng = np.random.default_rng(14)
    x = np.concatenate([_bits(rng, 4 * N), rng.uniform(-3, 3, N).astype(np.float32), np.array([0.0, -0.0, 0.4375, 0.6875, 1.1875, 2.4375, np.inf, -np.inf, np.nan], np.float32)])
    a, b = _run("atanf", np.ascontiguousarray(x))
    assert _same(a, b).all()


@pytest.mark.parametrize("name", ["atan2f", "hypotf"])
def test_atan2f_and_hypotf_equal_system_libm(name):
    """what dt_Lab_2_LCH and dt_JzAzBz_2_JzCzhz call (common/colorspaces_inline_conversions.h:594-606, :775-781)"""
    rng = np.random.default_rng(15)
    for y, x in _pairs(rng):
        a, b = _run(name, np.ascontiguousarray(y), np.ascontiguousarray(x))
        assert _same(a, b).all(), name
