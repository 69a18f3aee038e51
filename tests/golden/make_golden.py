"""Generate tests/golden/*.npz from the reference's own sources (oracle/_ref, built by
oracle/Makefile from /root/reference).  Run in the authoring container only:

    python tests/golden/make_golden.py

Inputs are seeded; the outputs are what libref_strict.so / libref_fast.so returned here
(gcc 13.3, x86-64 with FMA).  /root/reference does not exist on the GPU box, these files do."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import util  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"

for k, (name, filters) in enumerate(util.BAYER.items()):
    w, h = 214 + 2 * k, 135 + k          # three tile columns, two tile rows, ragged edges
    m = util.frame_natural(w, h, 100 + k, filters=filters)
    m[::9, ::7] = -0.01
    pm = np.array([1.0 + 0.1 * k, 1.0, 1.05], np.float32)
    np.savez_compressed(os.path.join(OUT, f"rcd_{name}.npz"), mosaic=m, pm=pm,
                        rgb_strict=util.ref_rcd(m, filters, tuple(pm), kind="strict"))

enc = util.srgb_encode_lut()
co_t = util.fit_unbounded_coeffs(enc)
img = util.rgba_test_image(96, 64, 9)
for case, kw in (("colorin_matrix", dict(matrix=util.MATRIX_CAM_TO_REC2020)),
                 ("colorin_clip", dict(matrix=util.MATRIX_CAM_TO_REC2020, clip=util.MATRIX_CLIP_IN)),
                 ("colorout_trc", dict(matrix=util.MATRIX_REC2020_TO_SRGB, lut_t=enc, co_t=co_t))):
    save = dict(rgba=img, matrix=kw["matrix"], out_strict=util.ref_convert(img, kind="strict", **kw),
                out_fast=util.ref_convert(img, kind="fast", **kw))
    if "clip" in kw:
        save["clip"] = kw["clip"]
    if "lut_t" in kw:
        save["lut_t_row"] = enc[0]
        save["co_t"] = co_t
    np.savez_compressed(os.path.join(OUT, f"color_{case}.npz"), **save)
rng = np.random.default_rng(77)
img = rng.normal(10, 1, (96, 128, 4)).astype(np.float32)
save = dict(img=img, inv_sigma2=np.float32(1.3), thr=np.array([0.3, 0.2, 0.1, 0.0], np.float32))
for scale in (0, 3):
    c, d, _ = util.ref_eaw_decompose(img, scale, 1.3)
    save[f"coarse_{scale}"], save[f"detail_{scale}"] = c, d
save["synth"] = util.ref_eaw_synthesize(img, save["detail_0"], (0.3, 0.2, 0.1, 0.0))
np.savez_compressed(os.path.join(OUT, "eaw.npz"), **save)
CASES = {"default_v8": {}, "no_bleach": dict(version=5), "high_bleach_hue": dict(version=8, saturation=60.0),
         "poly_curves": dict(shadows=0, highlights=1), "rational_curves": dict(shadows=2, highlights=2, contrast=1.5),
         "wide_dr_gamma22": dict(white_point_source=6.0, black_point_source=-10.0, output_power=2.2)}
np.savez_compressed(os.path.join(OUT, "filmic_data.npz"),
                    **{k: util.ref_filmic_commit(util.filmic_default_params(**v)) for k, v in CASES.items()})
work, export = util.profile_pair(util.REC2020_TO_XYZ_D50), util.profile_pair(util.SRGB_TO_XYZ_D50)
img = util.hdr_rgba(96, 64, 11)
blob = util.ref_filmic_commit(util.filmic_default_params())
np.savez_compressed(os.path.join(OUT, "filmic_agx.npz"), img=img, out_export=util.ref_filmic_agx(img, blob, work, export),
                    out_work=util.ref_filmic_agx(img, blob, work, None),
                    prepare=util.filmic_prepare(util.ref("strict"), "ref_filmic_prepare", 7, work, export))
img = (util.rgba_scene(150, 130, 21, noise=0.02) * 60).astype(np.float32)
np.savez_compressed(os.path.join(OUT, "nlmeans.npz"), img=img, out_profiled=util.ref_nlmeans(img),
                    out_lab=util.ref_nlmeans(img, center_weight=-1.0, sharpness=0.01, luma=0.8, chroma=0.6, K=3, P=2))
print("golden vectors written to", OUT)
img = util.lab_scene(180, 131, 31)
np.savez_compressed(os.path.join(OUT, "ll.npz"), img=img, out_default=util.ref_local_laplacian(img),
                    out_strong=util.ref_local_laplacian(img, sigma=0.2, shadows=1.5, highlights=0.1, clarity=1.0))
import ansel_b200 as ab  # noqa: E402
import test_cpu_oracle_pin as pin  # noqa: E402
img = util.hdr_rgba(140, 101, 41)
np.savez_compressed(os.path.join(OUT, "diffuse.npz"), img=img,
                    **{k: util.ref_diffuse(img, ab.diffuse_data(**pin._diffuse_cases()[k])) for k in ("sharpen_demosaic_aa", "gradient_sharpen", "inpaint_highlights")})
wp = util.profile_pair(util.REC2020_TO_XYZ_D50)
rgb, lab = util.hdr_rgba(120, 80, 51), util.lab_scene(120, 80, 52)
sp = util.profile_pair(util.SRGB_TO_XYZ_D50)
dec, cdec, enc, cenc = pin._srgb_curves(False)
np.savez_compressed(os.path.join(OUT, "labglue.npz"), rgb=rgb, lab=lab, lab_of_rgb=util.ref_rgb_to_lab(rgb, wp), rgb_of_lab=util.ref_lab_to_rgb(lab, wp),
                    lab_of_rgb_trc=util.ref_rgb_to_lab_trc(rgb, sp, dec, cdec, enc, cenc), rgb_of_lab_trc=util.ref_lab_to_rgb_trc(lab, sp, dec, cdec, enc, cenc))
rgba = util.hdr_rgba(90, 70, 61)
mos = util.frame_natural(128, 96, 62, iso=400.0)
np.savez_compressed(os.path.join(OUT, "demosaic_extra.npz"), rgba=rgba, mosaic=mos, smoothed2=util.ref_color_smoothing(rgba, 2),
                    geq_local=util.ref_green_eq(mos, util.BAYER["RGGB"], 1, iso=400.0), geq_both=util.ref_green_eq(mos, util.BAYER["RGGB"], 3, iso=400.0))
img = util.hdr_rgba(96, 64, 71)
save = dict(img=img)
for version, pc in ((0, 0), (0, 3), (1, 0), (2, 5), (3, 0), (3, 2), (4, 1)):
    blob = util.ref_filmic_commit(util.filmic_default_params(version=version, preserve_color=pc, saturation=10.0))
    save[f"data_v{version}_n{pc}"] = blob
    save[f"out_v{version}_n{pc}"] = util.ref_filmic_legacy(img, blob, work, export)
np.savez_compressed(os.path.join(OUT, "filmic_legacy.npz"), **save)
mos = util.frame_natural(300, 200, 81)
np.savez_compressed(os.path.join(OUT, "amaze.npz"), mosaic=mos, rgb_carried=util.ref_amaze(mos, util.BAYER["RGGB"]))

img = (util.rgba_scene(150, 110, 71) * 2.0).astype(np.float32)
rec = {"img": img}
for name, over in pin.RECONSTRUCT_CASES.items():
    blob = util.ref_filmic_commit(util.filmic_default_params(**over))
    rc, frame, mask = util.ref_filmic_reconstruct(img, blob)
    assert rc == 1
    rec["data_" + name], rec["frame_" + name], rec["mask_" + name] = blob, frame, mask
np.savez_compressed(os.path.join(OUT, "filmic_reconstruct.npz"), **rec)
