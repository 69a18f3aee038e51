"""The reference against itself for the modules added late in round 2: the strict build of its sources (what the kernels are pinned to, bit for
bit, by the -m gpu tests) against the release-flag build, on 2400x1600 frames.  CPU only (oracle/_ref):  python tools/parity_table_cpu.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", "tests"), os.path.join(HERE, "..")]
import util  # noqa: E402
import markesteijn_util as mu  # noqa: E402
import lmmse_util as lu  # noqa: E402
import blend_util as bu  # noqa: E402
import hl_laplacian_util as hu  # noqa: E402
from parity_table import distance, fmt  # noqa: E402

W, H = 2400, 1600
rows = []
m = util.frame_natural(W, H, 3)
rows.append(("demosaic Markesteijn, 1 pass (X-Trans)", distance(mu.ref(m, passes=1, kind="strict"), mu.ref(m, passes=1, kind="fast")), "tests/test_zz_markesteijn_gpu.py"))
rows.append(("demosaic Markesteijn, 3 passes (X-Trans)", distance(mu.ref(m, passes=3, kind="strict"), mu.ref(m, passes=3, kind="fast")), "tests/test_zz_markesteijn_gpu.py"))
rows.append(("demosaic LMMSE, median refinement", distance(lu.ref(m, util.BAYER["RGGB"], 1, kind="strict"), lu.ref(m, util.BAYER["RGGB"], 1, kind="fast")),
             "tests/test_zz_lmmse_gpu.py (tiles from zeroed planes: DESIGN.md row f12)"))
a, b, form = bu.frames(W, H, 3)
p = bu.params(mode="normal", opacity=70.0, mask_mode=bu.MASK_ENABLED | bu.MASK_SHAPE | bu.MASK_PARAMETRIC, drawn=1,
              channels={0: (0.05, 0.2, 0.8, 1.0), 5: (0.0, 0.0, 0.7, 0.9)})
rows.append(("blending, drawn + parametric mask, normal operator", distance(bu.ref(a, b, p, form, kind="strict")[1], bu.ref(a, b, p, form, kind="fast")[1]),
             "tests/test_zz_blend_gpu.py"))
img = hu.clipped_mosaic(W, H, 45, blobs=9)
s, norm = hu.ref(img, util.BAYER["RGGB"], hu.clips_of(), iterations=4, lib=util.ref("strict"))
f, _ = hu.ref(img, util.BAYER["RGGB"], hu.clips_of(), iterations=4, norm=norm, lib=util.ref("fast"))
rows.append(("highlights, guided laplacians, 4 iterations (same normalization vector)", distance(s[..., None].repeat(3, 2), f[..., None].repeat(3, 2)),
             "tests/test_zz_hl_laplacian_gpu.py"))
print(f"| module | strict vs fast (the reference against itself), {W}x{H} | the kernel is bit-identical to strict in |")
print("|---|---|---|")
for name, d, where in rows:
    print(f"| {name} | {fmt(d)} | {where} |")
