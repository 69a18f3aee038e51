"""Generate tests/golden/bilateral.npz from the reference's pixel/bilateral.c (oracle/_ref, strict build, one splat slice) on
the cases of tests/bilateral_util.py.  Run in the authoring container only:  python tests/golden/make_golden_bilateral.py"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import bilateral_util as bu  # noqa: E402
import util  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "bilateral.npz")
np.savez_compressed(out, **{name: bu.ref_bilateral(*bu.case(name), threads=1) for name in bu.CASES})
print("written", out, os.path.getsize(out) // 1024, "KiB")
