"""debug: where does the group kernel differ from the oracle on the extreme-pixel frame"""
import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
ab.init()
def cuda_nlm(img, *, scattering=0.0, scale=1.0, luma=1.0, chroma=1.0, center_weight=0.1, sharpness=0.005, P=1, K=7, decimate=0, norm=(1.0, 1.0, 1.0, 1.0)):
    h, w = img.shape[:2]
    d_in = torch.from_numpy(np.ascontiguousarray(img)).cuda(); d_out = torch.zeros_like(d_in)
    ab.check(ab.lib().b200_nlmeans_denoise_dev(d_in.data_ptr(), d_out.data_ptr(), w, h, C.c_float(scattering), C.c_float(scale), C.c_float(luma), C.c_float(chroma), C.c_float(center_weight),
                                               C.c_float(sharpness), P, K, decimate, (C.c_float * 4)(*norm), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize(); return d_out.cpu().numpy()
def same(a, b): return (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
base = (util.rgba_scene(160, 130, 4, noise=0.02) * 60).astype(np.float32)
def frame(which):
    img = base.copy()
    if "zeros" in which: img[20:60, 30:90, :3] = 0.0
    if "tiny" in which: img[70:90, 10:50, :3] = 1e-30
    if "inf" in which: img[100, 100, :3] = (np.inf, 1.0, 2.0)
    if "nan" in which: img[101, 120, :3] = (np.nan, 1.0, 2.0)
    if "huge" in which: img[110, 20, :3] = (3e38, -3e38, 1e19)
    if "sub" in which: img[5, 5, :3] = (1e-40, 1e-44, 0.0)
    return img
kw = dict(K=3, center_weight=-1.0, sharpness=0.01)
for which in (["zeros"], ["tiny"], ["inf"], ["nan"], ["huge"], ["sub"], []):
    img = frame(which)
    want = util.oracle_nlmeans(img, **kw)
    for env in ("group", "chunks"):
        os.environ.pop("B200_NLM_CHUNKS", None); os.environ.pop("B200_NLM_NO_PIPE", None)
        if env == "chunks": os.environ["B200_NLM_CHUNKS"] = "1"
        if env == "group": os.environ["B200_NLM_NO_PIPE"] = "1"
        got = cuda_nlm(img, **kw)
        bad = ~same(got, want)
        rows = np.unique(np.argwhere(bad)[:, 0]) if bad.any() else []
        cols = np.unique(np.argwhere(bad)[:, 1]) if bad.any() else []
        print(which, env, int(bad.sum()), "rows", (rows[:3], rows[-3:]) if len(rows) else "", "cols", (cols[:3], cols[-3:]) if len(cols) else "")
        if bad.any():
            i = tuple(np.argwhere(bad)[0]); print("   first", i, got[i], want[i], got[i[0], i[1]], want[i[0], i[1]])
