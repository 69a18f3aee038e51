"""Generate tests/golden/blend.npz from the reference's own lines (oracle/_ref, strict build): dt_develop_blend_process on a 96x64 frame for a
spread of the configurations of tests/blend_util.py in both colour spaces (every third one, plus all that go through powf / atan2f / hypotf).
Inputs are regenerated from the seed.  Run in the authoring container only:  python tests/golden/make_golden_blend.py"""
import os
import sys

import numpy as np

TESTS = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [TESTS, os.path.dirname(TESTS)]
import util  # noqa: E402
import blend_util as bu  # noqa: E402

assert util.ref("strict") is not None, "build oracle/_ref first (make -C oracle ref)"
save = {}
for name, kw, uses_form in bu.golden_configs():
    a, b, form = bu.golden_frames(kw)
    rc, out, mask = bu.ref(a, b, bu.params(**kw), form if uses_form else None)
    assert rc == 0, name
    save[name], save[name + "_mask"] = out, mask
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "blend.npz")
np.savez_compressed(out, **save)
print("written", out, os.path.getsize(out) // 1024, "KiB,", len(save) // 2, "configurations")
