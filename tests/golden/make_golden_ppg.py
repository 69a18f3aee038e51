"""Generate tests/golden/ppg.npz from the reference's own lines (oracle/_ref, strict build): the PPG demosaicer on the
cases of tests/ppg_util.py.  Run in the authoring container only:  python tests/golden/make_golden_ppg.py"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import ppg_util as pu  # noqa: E402
import util  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ppg.npz")
np.savez_compressed(out, **{name: pu.ref_ppg(*pu.case(name)) for name in pu.CASES})
print("written", out, os.path.getsize(out) // 1024, "KiB")
