import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, util
import test_amaze_gpu as t
for name in ("BGGR", "GRBG", "GBRG", "RGGB"):
    f = util.BAYER[name]
    for (w, h) in ((129, 161), (33, 34), (200, 161), (129, 200), (161, 129)):
        m = util.frame_natural(w, h, 5, filters=f)
        got = t.cuda_amaze(m, f)[..., :3]; want = util.oracle_amaze(m, f)[..., :3]
        bad = ~t.same_bits(got, want)
        if bad.any():
            idx = np.argwhere(bad)
            print(name, (w, h), "bad", int(bad.sum()), "rows", idx[:, 0].min(), idx[:, 0].max(), "cols", idx[:, 1].min(), idx[:, 1].max(), "lanes", sorted(set(idx[:, 2].tolist())), idx[:4].tolist(), got[tuple(idx[0])], want[tuple(idx[0])])
        else:
            print(name, (w, h), "ok")
