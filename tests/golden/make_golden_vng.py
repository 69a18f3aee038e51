"""Generate tests/golden/vng.npz from the reference's own lines (oracle/_ref, strict serial build): VNG4 and the dual blend
on the cases of tests/vng_util.py.  Run in the authoring container only:  python tests/golden/make_golden_vng.py"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import util  # noqa: E402
import vng_util as vu  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
save = {}
for name in vu.CASES:
    m, filters, x, y = vu.case(name)
    save["vng_" + name] = vu.ref_vng(m, filters, x, y)
    save["dual_" + name] = vu.ref_dual(vu.sharp_frame(m, filters, x, y), m, filters, x, y, 0.2)
for name in vu.XTRANS_CASES:
    m, x, y = vu.xtrans_case(name)
    save["xtrans_" + name] = vu.ref_vng_xtrans(m, x, y)
    save["xtrans_" + name][..., 3] = 0.0        # lane 3 is uninitialised memory in the reference: not a golden value
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "vng.npz")
np.savez_compressed(out, **save)
print("written", out, os.path.getsize(out) // 1024, "KiB")
