"""Generate tests/golden/lmmse.npz from the reference's own lines (oracle/_ref, strict build, compiled without OpenMP: the serial walk of the
tiles) on the cases of tests/lmmse_util.py, modes 0..4.  Run in the authoring container only:  python tests/golden/make_golden_lmmse.py"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import util  # noqa: E402
import lmmse_util as lu  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
save = {}
for name, modes in (("rggb", (1, 4)), ("bggr_odd", (0,)), ("gbrg_small", (0, 1, 4))):
    m, f = lu.case(name)
    for mode in modes:
        save[f"m{mode}_{name}"] = lu.ref(m, f, mode)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lmmse.npz")
np.savez_compressed(out, **save)
print("written", out, os.path.getsize(out) // 1024, "KiB")
