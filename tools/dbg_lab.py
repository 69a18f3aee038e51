import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, util, ansel_b200 as ab
import test_labglue_gpu as t
rgb = util.hdr_rgba(1000, 700, 6)
rc, got = t.cuda_transform(rgb, ab.CS_RGB, ab.CS_LAB, False)
print("rc", rc, ab.lib().b200_last_error())
want = util.oracle_rgb_to_lab(rgb, t.WORK)
bad = ~t.same_bits(got, want)
print("bad per lane", [int(bad[..., c].sum()) for c in range(4)], "of", bad[..., 0].size)
idx = np.argwhere(bad)[:6]
for y, x, c in idx:
    print((y, x, c), "in", rgb[y, x], "got", got[y, x], "want", want[y, x])
