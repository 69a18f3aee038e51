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
# spill geometries: right border running over (0 < width+16-(left+160) < 16) and lower border running past row 159
for (w, h) in ((129 + 128, 161), (140, 100), (300, 140), (137, 137), (1300, 900)):
    f = util.BAYER["RGGB"]
    m = util.frame_natural(w, h, 7, filters=f)
    got = t.cuda_amaze(m, f)[..., :3]; want = util.oracle_amaze(m, f)[..., :3]
    bad = ~t.same_bits(got, want)
    print("spill", (w, h), "bad", int(bad.sum()), np.argwhere(bad)[:4].tolist())
w, h = util.SIZE_45MP
f = util.BAYER["RGGB"]
m = util.frame_natural(w, h, util.SEEDS[0])
a = t.cuda_amaze(m, f); b = t.cuda_amaze(m, f)
print("45mp deterministic", bool(t.same_bits(a, b).all()), "finite", bool(np.isfinite(a).all()), "min", float(a[..., :3].min()), "max", float(a[..., :3].max()))
flat = np.full((512, 768), 0.25, np.float32)
print("flat dev", float(np.abs(t.cuda_amaze(flat, f)[..., :3] - 0.25).max()))
