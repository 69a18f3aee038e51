"""Generate tests/golden/hl_laplacian.npz from the reference's own lines (oracle/_ref, strict build): process_laplacian() on two of the
cases of tests/test_cpu_hl_laplacian.py and one RGBA frame, each with the normalization vector the run used (an OpenMP reduction: the
vector is part of the fixture).  Run in the authoring container only:  python tests/golden/make_golden_hl_laplacian.py"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import util  # noqa: E402
import hl_laplacian_util as hu  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
save = {}
for name, (w, h, f, kw) in hu.GOLDEN.items():
    img = hu.clipped_mosaic(w, h, len(name)) if f else hu.clipped_rgba(w, h, len(name))
    out, norm = hu.ref(img, f, hu.clips_of(), **kw)
    save[name], save[name + "_norm"] = out, norm
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hl_laplacian.npz")
np.savez_compressed(out, **save)
print("written", out, os.path.getsize(out) // 1024, "KiB")
