"""Generate tests/golden/markesteijn.npz from the reference's own lines (oracle/_ref, strict build): Markesteijn with one and
three passes on the cases of tests/markesteijn_util.py.  Run in the authoring container only:
python tests/golden/make_golden_markesteijn.py"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import util  # noqa: E402
import markesteijn_util as mu  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
save = {}
for name in mu.CASES:
    m, x, y = mu.case(name)
    for passes in (1, 3):
        r = mu.ref(m, x, y, passes)
        r[..., 3] = 0.0          # lane 3 is not a result of the reference
        save[f"p{passes}_{name}"] = r.astype(np.float32)
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "markesteijn.npz")
np.savez_compressed(out, **save)
print("written", out, os.path.getsize(out) // 1024, "KiB")
