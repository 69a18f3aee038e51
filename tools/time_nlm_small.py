import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
ab.init()
w, h = 2048, 1536   # 444 chunks = 3 per SM
img = torch.rand((h, w, 4), device="cuda") * 20
out = torch.empty_like(img)
L = ab.lib(); s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    ab.check(L.b200_nlmeans_denoise_dev(img.data_ptr(), out.data_ptr(), w, h, 0.0, 1.0, 1.0, 1.0, 0.1, 0.005, 1, 7, 0, (C.c_float*4)(1,1,1,1), s))
torch.cuda.synchronize()
