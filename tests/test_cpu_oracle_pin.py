"""CPU tests: the oracle against (a) the reference's own sources compiled here (oracle/_ref, when
present) and (b) the committed golden vectors those sources produced (tests/golden, always)."""
import os

import numpy as np
import pytest

import util


def same_bits(a, b):
    return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))


@pytest.fixture(scope="module", autouse=True)
def _build():
    util.build_oracle()


need_ref = pytest.mark.skipif(util.ref("strict") is None and not os.path.isdir("/root/reference/src"),
                              reason="oracle/_ref not built (no /root/reference)")


@need_ref
@pytest.mark.parametrize("name", list(util.BAYER))
@pytest.mark.parametrize("size", [(16, 16), (117, 131), (206, 206), (207, 113), (1024, 768)])
def test_rcd_oracle_equals_reference_strict(name, size):
    w, h = size
    m = util.frame_uniform(w, h, 1)
    m[::7, ::5] = 0.0
    m[3::11, 2::9] = -0.02
    got = util.oracle_rcd(m, util.BAYER[name])
    want = util.ref_rcd(m, util.BAYER[name], kind="strict")
    mask = util.oracle_rcd_mask(m, util.BAYER[name])
    bad = (~same_bits(got[..., :3], want[..., :3])).any(axis=2) & ((mask & 1) == 0)
    assert not bad.any()
    assert (mask & 1).sum() < 4 * (w + h)


@need_ref
def test_rcd_reference_undefined_pixels_are_real():
    """The masked set is where the reference itself is not a function of its input: poisoning its
    uninitialised scratch moves (some of) those pixels and nothing else."""
    m = util.frame_uniform(1024, 768, 2)
    f = util.BAYER["RGGB"]
    a = util.ref_rcd(m, f, kind="strict", poison=0.0)
    b = util.ref_rcd(m, f, kind="strict", poison=0.37)
    moved = (~same_bits(a[..., :3], b[..., :3])).any(axis=2)
    mask = util.oracle_rcd_mask(m, f)
    assert moved.any()
    assert not (moved & ((mask & 1) == 0)).any()


@need_ref
def test_rcd_fast_build_distance_is_what_design_md_says():
    """The release (-ffast-math) build of the SAME reference source is not within 1 ulp of its own
    strict build; record the distribution the parity statement in DESIGN.md quotes."""
    m = util.frame_natural(1024, 768, util.SEEDS[0])
    f = util.BAYER["RGGB"]
    s = util.ref_rcd(m, f, kind="strict")
    q = util.ref_rcd(m, f, kind="fast")
    mask = util.oracle_rcd_mask(m, f)
    u = util.ulp_distance(s[..., :3], q[..., :3])[(mask & 1) == 0]
    assert (u > 1).mean() > 0.01          # the release build is NOT within 1 ulp of the source semantics
    assert np.abs(s[..., :3] - q[..., :3])[(mask & 1) == 0].max() < 1e-3


@need_ref
@pytest.mark.parametrize("kind,fp", [("strict", util.FP_STRICT), ("fast", util.FP_CONTRACT)])
def test_colour_oracle_equals_reference(kind, fp):
    enc, dec = util.srgb_encode_lut(), util.srgb_decode_lut()
    co_t, co_s = util.fit_unbounded_coeffs(enc), util.fit_unbounded_coeffs(dec)
    img = util.rgba_test_image(512, 384, 4)
    for kw in (dict(matrix=util.MATRIX_CAM_TO_REC2020),
               dict(matrix=util.MATRIX_CAM_TO_REC2020, clip=util.MATRIX_CLIP_IN),
               dict(matrix=util.MATRIX_CAM_TO_REC2020, lut_s=dec, co_s=co_s),
               dict(matrix=util.MATRIX_REC2020_TO_SRGB, lut_t=enc, co_t=co_t),
               dict(matrix=util.MATRIX_REC2020_TO_SRGB, clip=util.MATRIX_CLIP_IN, lut_s=dec, co_s=co_s, lut_t=enc, co_t=co_t)):
        r = util.ref_convert(img, kind=kind, **kw)
        o = util.oracle_convert(img, fp=fp, **kw)
        assert same_bits(r, o).all()


@need_ref
@pytest.mark.parametrize("scale", [0, 2, 5])
def test_eaw_oracle_equals_reference_strict(scale):
    rng = np.random.default_rng(scale)
    img = rng.normal(10, 1, (203, 301, 4)).astype(np.float32)
    oc, od, osum = util.oracle_eaw_decompose(img, scale, 1.7)
    rc, rd, rsum = util.ref_eaw_decompose(img, scale, 1.7)
    assert same_bits(oc, rc).all() and same_bits(od, rd).all()
    assert np.allclose(osum, rsum, rtol=2e-6)     # the reference sums in float, the oracle in double
    thr = (0.3, 0.2, 0.1, 0.0)
    assert same_bits(util.oracle_eaw_synthesize(img, od, thr), util.ref_eaw_synthesize(img, od, thr)).all()


# ---- golden vectors: produced by tests/golden/make_golden.py from oracle/_ref, committed ----------
def _golden(name):
    return np.load(os.path.join(util.GOLDEN_DIR, name))


@pytest.mark.parametrize("name", list(util.BAYER))
def test_rcd_oracle_equals_golden(name):
    g = _golden(f"rcd_{name}.npz")
    got = util.oracle_rcd(g["mosaic"], util.BAYER[name], tuple(g["pm"]))
    mask = util.oracle_rcd_mask(g["mosaic"], util.BAYER[name], tuple(g["pm"]))
    bad = (~same_bits(got[..., :3], g["rgb_strict"][..., :3])).any(axis=2) & ((mask & 1) == 0)
    assert not bad.any()


@pytest.mark.parametrize("case", ["colorin_matrix", "colorin_clip", "colorout_trc"])
def test_colour_oracle_equals_golden(case):
    g = _golden(f"color_{case}.npz")
    kw = dict(matrix=g["matrix"])
    if "clip" in g.files:
        kw["clip"] = g["clip"]
    if "lut_t_row" in g.files:
        kw["lut_t"] = np.ascontiguousarray(np.tile(g["lut_t_row"], (3, 1)))
        kw["co_t"] = g["co_t"]
    assert same_bits(util.oracle_convert(g["rgba"], fp=util.FP_STRICT, **kw), g["out_strict"]).all()
    assert same_bits(util.oracle_convert(g["rgba"], fp=util.FP_CONTRACT, **kw), g["out_fast"]).all()


def test_eaw_oracle_equals_golden():
    g = _golden("eaw.npz")
    for scale in (0, 3):
        oc, od, _ = util.oracle_eaw_decompose(g["img"], scale, float(g["inv_sigma2"]))
        assert same_bits(oc, g[f"coarse_{scale}"]).all() and same_bits(od, g[f"detail_{scale}"]).all()
    assert same_bits(util.oracle_eaw_synthesize(g["img"], g["detail_0"], tuple(g["thr"])), g["synth"]).all()
