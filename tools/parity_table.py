"""The distance of every hot-path module to the SHIPPED build of the reference, at the bench frame size (run on the GPU box):
per module the CUDA output (through the C ABI, device buffers) against oracle/_ref/libref_fast.so (the reference's own sources with
its release flags: -O3 -ffast-math -ffp-contract=fast) and against libref_strict.so (the same sources, C float semantics: what the
kernels are pinned to bit for bit), next to the distance between the two builds of the reference itself.

    python tools/parity_table.py [--width 8256 --height 5504] > gpurun_out/parity_vs_release_build.md

Test infrastructure (it loads oracle/_ref); nothing of the product routes through it."""
import argparse
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import util  # noqa: E402
import ansel_b200 as ab  # noqa: E402


def ordered(a):
    """floats -> integers whose difference is the distance in ULPs (both zeros at 0)"""
    i = a.view(np.int32).astype(np.int64)
    return np.where(i < 0, -(i & 0x7fffffff), i)


def distance(a, b, mask=None):
    """(fraction of floats differing, fraction > 1 ULP, max ULP, max abs) over the three colour lanes"""
    a, b = a[..., :3], b[..., :3]
    ok = ~(np.isnan(a) & np.isnan(b))
    if mask is not None:
        ok &= mask[..., None]
    ulp = np.abs(ordered(np.ascontiguousarray(a)) - ordered(np.ascontiguousarray(b)))[ok]
    n = max(ulp.size, 1)
    return (float((ulp > 0).sum()) / n, float((ulp > 1).sum()) / n, int(ulp.max()) if ulp.size else 0,
            float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64))[ok])) if ulp.size else 0.0)


def fmt(d):
    return f"{100 * d[0]:.3f} % / {100 * d[1]:.3f} % / {d[2]} / {d[3]:.3g}"


def dev_module(op, data, src, w, h, channels_in=4, filters=0):
    piece = ab.make_piece(w, h, filters=filters, channels=channels_in, devid=0)
    piece.data, piece.data_size = C.addressof(data), C.sizeof(data)
    d_in = torch.from_numpy(np.ascontiguousarray(src)).cuda()
    d_out = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    ab.check(getattr(ab.lib(), f"b200_{op}_process_dev")(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return d_out.cpu().numpy()


def ref_nlm_module(img, dn, kind):
    R, O = util.ref(kind), util.oracle()
    h, w = img.shape[:2]
    f4 = lambda v: (C.c_float * 4)(*v)  # noqa: E731
    plan = np.zeros(51, np.float32)
    O.orc_dn_plan_export_nlm(C.byref(dn), C.c_float(1.0), w, h, f4((2.0, 1.0, 1.5, 0.0)), f4((1.0,) * 4), util.fptr(plan))
    wb, p, a, b, bias = plan[1:5], plan[5:9], plan[9], plan[10], plan[11]
    src, pre = util.aligned_empty(img.shape), util.aligned_empty(img.shape)
    src[...] = img
    R.ref_dn_precondition_v2(util.fptr(src), util.fptr(pre), w, h, C.c_float(a), f4(p), C.c_float(b), f4(wb))
    out = util.ref_nlmeans(pre, kind=kind, sharpness=float(np.float32(0.045) / np.float32(9)), center_weight=float(np.float32(dn.central_pixel_weight)), P=1, K=7)
    buf = util.aligned_empty(img.shape)
    buf[...] = out
    R.ref_dn_backtransform_v2(util.fptr(buf), w, h, C.c_float(a), f4(p), C.c_float(b), C.c_float(bias), f4(wb))
    return np.array(buf)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=8256)
    ap.add_argument("--height", type=int, default=5504)
    ap.add_argument("--skip", default="")
    a = ap.parse_args()
    w, h = a.width, a.height
    assert util.ref("fast") is not None and util.ref("strict") is not None, "oracle/_ref is not built"
    ab.init()
    rows = []
    filters = util.BAYER["RGGB"]
    mosaic = util.frame_natural(w, h, 20260922)
    work, export = util.profile_pair(util.REC2020_TO_XYZ_D50), util.profile_pair(util.SRGB_TO_XYZ_D50)
    enc = util.srgb_encode_lut()
    co_t = util.fit_unbounded_coeffs(enc)

    def add(name, cuda, strict, fast, mask=None, note=""):
        rows.append((name, fmt(distance(cuda, fast, mask)), fmt(distance(cuda, strict, mask)), fmt(distance(strict, fast, mask)), note))
        print(f"[{time.strftime('%H:%M:%S')}] {name} done", file=sys.stderr, flush=True)

    skip = set(a.skip.split(","))
    # demosaic (RCD): pixels the reference leaves undefined (uninitialised scratch) are masked
    dem = dev_module("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), mosaic, w, h, channels_in=1, filters=filters)
    defined = (util.oracle_rcd_mask(mosaic, filters) & 1) == 0
    add("demosaic RCD", dem, util.ref_rcd(mosaic, filters, kind="strict"), util.ref_rcd(mosaic, filters, kind="fast"), defined,
        f"{int((~defined).sum())} px the reference reads uninitialised scratch for are excluded")
    if "amaze" not in skip:
        amz = dev_module("demosaic", ab.demosaic_data(ab.DEMOSAIC_AMAZE), mosaic, w, h, channels_in=1, filters=filters)
        add("demosaic AMaZE", amz, util.ref_amaze(mosaic, filters, kind="strict", threads=os.cpu_count()), util.ref_amaze(mosaic, filters, kind="fast", threads=os.cpu_count()),
            note="the reference carries scratch from tile to tile per thread: a handful of pixels depend on its thread count")
    rgba = np.ascontiguousarray(dem)
    rgba[~defined] = 0.2
    # denoise (profiled), non-local means, the C3 parameters
    dn = ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=7)
    den = dev_module("denoiseprofile", dn, rgba, w, h)
    add("denoiseprofile NLM (P=1, K=7)", den, ref_nlm_module(rgba, dn, "strict"), ref_nlm_module(rgba, dn, "fast"))
    # colorin, colorout
    conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
    cin = dev_module("colorin", ab.colorin_data(conv_in), den, w, h)
    add("colorin (matrix)", cin, util.ref_convert(den, util.MATRIX_CAM_TO_REC2020, kind="strict"), util.ref_convert(den, util.MATRIX_CAM_TO_REC2020, kind="fast"),
        note="default flavour B200_FP_CONTRACT is pinned to the release build")
    blob = np.load(os.path.join(util.GOLDEN_DIR, "filmic_data.npz"))["default_v8"]
    fil = dev_module("filmicrgb", ab.filmic_piece(blob, work, export), cin, w, h)
    add("filmicrgb (v8 defaults)", fil, util.ref_filmic_agx(cin, blob, work, export, kind="strict"), util.ref_filmic_agx(cin, blob, work, export, kind="fast"))
    conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=co_t)
    cout = dev_module("colorout", ab.colorout_data(conv_out), fil, w, h)
    add("colorout (matrix + sRGB curve)", cout, util.ref_convert(fil, util.MATRIX_REC2020_TO_SRGB, lut_t=enc, co_t=co_t, kind="strict"),
        util.ref_convert(fil, util.MATRIX_REC2020_TO_SRGB, lut_t=enc, co_t=co_t, kind="fast"))
    if "diffuse" not in skip:
        dd = ab.diffuse_data(**ab.DIFFUSE_PRESETS["sharpen_demosaic_aa"])
        dif = dev_module("diffuse", dd, cin, w, h)
        add("diffuse (sharpen demosaicing preset)", dif, util.ref_diffuse(cin, dd, kind="strict"), util.ref_diffuse(cin, dd, kind="fast"))
    if "bilat" not in skip:
        lab = util.lab_scene(w, h, 3)
        ll = dev_module("bilat", ab.bilat_data(), lab, w, h)
        add("bilat (local Laplacian defaults)", ll, util.ref_local_laplacian(lab, kind="strict"), util.ref_local_laplacian(lab, kind="fast"))

    print(f"# Distance to the shipped (release-flag) build of the reference, {w}x{h}\n")
    print("Each cell: floats differing / differing by more than 1 ULP / largest ULP distance / largest absolute difference, over the")
    print("three colour lanes of every pixel.  `fast` = `oracle/_ref/libref_fast.so` (the reference's sources, `-O3 -ffast-math")
    print("-ffp-contract=fast`, OpenMP), `strict` = `libref_strict.so` (same sources, C float semantics).  The kernels are pinned to")
    print("`strict`; their distance to `fast` is, column by column, the release build's own distance to its source.\n")
    print("| module | CUDA vs fast | CUDA vs strict | strict vs fast (the reference against itself) | note |")
    print("|---|---|---|---|---|")
    for r in rows:
        print("| " + " | ".join(r) + " |")


if __name__ == "__main__":
    main()
