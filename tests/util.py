"""Shared helpers for the parity tests: oracle/_ref loaders, ULP distance, synthetic frames.

The oracle (oracle/liboracle.so) and the compiled reference (oracle/_ref/*.so) are CHECKERS; only
tests/, __graft_entry__.smoke() and bench.py's CPU legs may load them.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
FP = C.POINTER(C.c_float)

BAYER = {"RGGB": 0x94949494, "BGGR": 0x16161616, "GRBG": 0x61616161, "GBRG": 0x49494949}
SEEDS = (20260922, 1, 2)

# synthetic frame sizes of SURVEY.md 8
SIZE_24MP = (6000, 4000)
SIZE_45MP = (8256, 5504)
SIZE_100MP = (11648, 8736)


def build_oracle() -> None:
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    if os.path.isdir("/root/reference/src"):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)


_cache = {}


def oracle() -> C.CDLL:
    if "oracle" not in _cache:
        path = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(path):
            build_oracle()
        _cache["oracle"] = C.CDLL(path)
    return _cache["oracle"]


def ref(kind: str = "strict"):
    """oracle/_ref/libref_{strict,fast}.so or None when it was never built (no /root/reference)."""
    key = "ref_" + kind
    if key not in _cache:
        path = os.path.join(ORACLE_DIR, "_ref", f"libref_{kind}.so")
        _cache[key] = C.CDLL(path) if os.path.exists(path) else None
    return _cache[key]


def fptr(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(FP)


def ulp_distance(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units of float32 representable values (sign-magnitude ordered)."""
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


# ---- RCD through the checkers ------------------------------------------------------------
def oracle_rcd(mosaic: np.ndarray, filters: int, pm=(1.0, 1.0, 1.0), fill: float = 0.0) -> np.ndarray:
    h, w = mosaic.shape
    out = np.zeros((h, w, 4), np.float32)
    f = oracle().orc_rcd_demosaic
    f.restype = C.c_int
    rc = f(fptr(out), fptr(mosaic), w, h, C.c_uint32(filters), (C.c_float * 3)(*pm), C.c_float(fill))
    assert rc in (0, 1)
    return out


def oracle_rcd_mask(mosaic: np.ndarray, filters: int, pm=(1.0, 1.0, 1.0)) -> np.ndarray:
    """bit 0: colour depends on memory the reference never initialised; bit 1: alpha never written."""
    h, w = mosaic.shape
    mask = np.zeros((h, w), np.uint8)
    f = oracle().orc_rcd_undefined_mask
    f.restype = C.c_int
    rc = f(mask.ctypes.data_as(C.POINTER(C.c_uint8)), fptr(mosaic), w, h, C.c_uint32(filters), (C.c_float * 3)(*pm))
    assert rc in (0, 1)
    return mask


def ref_rcd(mosaic: np.ndarray, filters: int, pm=(1.0, 1.0, 1.0), kind: str = "strict", poison: float = 0.0):
    lib = ref(kind)
    if lib is None:
        return None
    h, w = mosaic.shape
    out = np.zeros((h, w, 4), np.float32)
    f = lib.ref_rcd_demosaic
    f.restype = C.c_int
    rc = f(fptr(out), fptr(mosaic), w, h, C.c_uint32(filters), (C.c_float * 3)(*pm), C.c_float(poison))
    assert rc in (0, 1)
    return out


# ---- synthetic Bayer frames ----------------------------------------------
def frame_uniform(w: int, h: int, seed: int) -> np.ndarray:
    """D-uniform: i.i.d. U[0,1) mosaic."""
    return np.random.Generator(np.random.PCG64(seed)).random((h, w), dtype=np.float32)


def _scene(w: int, h: int, rng) -> np.ndarray:
    xn = (np.arange(w, dtype=np.float32) / max(w, 1))[None, :]
    yn = (np.arange(h, dtype=np.float32) / max(h, 1))[:, None]
    s = (0.35 + 0.25 * xn + 0.1 * yn).astype(np.float32)
    for _ in range(6):
        fx, fy = rng.uniform(0.5, 6.0, 2)
        ph = rng.uniform(0, 2 * np.pi)
        ax = (2 * np.pi * fx * xn + ph).astype(np.float32)
        ay = (2 * np.pi * fy * yn).astype(np.float32)
        s += np.float32(0.06) * (np.cos(ax) * np.cos(ay) - np.sin(ax) * np.sin(ay))
    for _ in range(3):
        cx, cy, rad = rng.uniform(0.2, 0.8), rng.uniform(0.2, 0.8), rng.uniform(0.05, 0.2)
        d = np.sqrt((xn - np.float32(cx)) ** 2 + (yn - np.float32(cy)) ** 2)
        t = np.clip((d - np.float32(rad)) * np.float32(120.0), -60.0, 60.0)
        s += np.float32(0.25) / (np.float32(1.0) + np.exp(t))
    return s.astype(np.float32)


def cfa_colours(w: int, h: int, filters: int) -> np.ndarray:
    """FC() of every site (develop/imageop_math.h:190-193) as an (h, w) uint8 array."""
    rows = np.arange(h, dtype=np.uint32)[:, None]
    cols = np.arange(w, dtype=np.uint32)[None, :]
    sh = (((rows << 1) & 14) + (cols & 1)) << 1
    return ((np.uint32(filters) >> sh) & 3).astype(np.uint8)


def frame_natural(w: int, h: int, seed: int, filters: int = BAYER["RGGB"], iso: float = 100.0) -> np.ndarray:
    """D-natural: smooth scene through the CFA with Poisson-Gaussian noise, exact 0s and 1s sprinkled in."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = _scene(w, h, rng)
    gains = np.array([0.5, 1.0, 0.65, 1.0], np.float32)
    m = s * gains[cfa_colours(w, h, filters)]
    a = np.float32(1e-4 * (iso / 100.0))
    m += rng.standard_normal((h, w), dtype=np.float32) * np.sqrt(np.maximum(a * m, 0))
    np.clip(m, 0.0, 1.0, out=m)
    r = rng.random((h, w), dtype=np.float32)
    m[r < 0.005] = 0.0
    m[r > 0.995] = 1.0
    return np.ascontiguousarray(m, dtype=np.float32)


def frame_edge(w: int, h: int, kind: str) -> np.ndarray:
    """D-edge cases."""
    if kind == "zeros":
        return np.zeros((h, w), np.float32)
    if kind == "ones":
        return np.ones((h, w), np.float32)
    if kind == "impulses":
        m = np.full((h, w), 0.25, np.float32)
        m[::17, ::13] = 1.0
        m[5::23, 7::19] = 0.0
        return m
    if kind == "negative":
        m = frame_uniform(w, h, 7)
        m[::3, ::5] = -0.01
        return m
    if kind == "tiny":
        m = frame_uniform(w, h, 8) * 1e-3
        m[::2, ::3] = 1e-30
        m[1::4, 1::5] = 1e-41  # subnormal: flushed by FTZ/DAZ on both sides
        return m.astype(np.float32)
    raise ValueError(kind)
