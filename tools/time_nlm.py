"""Development timing of the non-local-means core at 45 MP (not the bench contract).  B200_NLM_G / B200_NLM_CHUNKS /
B200_NLM_IEEE_DIV select the variants; argv: [P K]."""
import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
ab.init()
w, h = util.SIZE_45MP
P, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1, 7)
n = int(os.environ.get("NLM_REPS", "3"))
img = torch.rand((h, w, 4), device="cuda") * 20
out = torch.empty_like(img)
L = ab.lib(); s = torch.cuda.current_stream().cuda_stream
def run(): ab.check(L.b200_nlmeans_denoise_dev(img.data_ptr(), out.data_ptr(), w, h, 0.0, 1.0, 1.0, 1.0, 0.1, 0.005, P, K, 0, (C.c_float*4)(1,1,1,1), s))
run(); torch.cuda.synchronize(); ts=[]
for _ in range(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("B200_NLM"))
print(f"NLM K={K} P={P} 45MP [{tag}] median ms {np.median(ts):.2f}  MP/s {w*h/np.median(ts)/1e3:.0f}")
