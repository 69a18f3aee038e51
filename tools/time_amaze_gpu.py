import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
ab.init()
w, h = util.SIZE_45MP
m = torch.from_numpy(util.frame_natural(w, h, 1)).cuda()
out = torch.empty((h, w, 4), device="cuda")
L = ab.lib(); s = torch.cuda.current_stream().cuda_stream
piece = ab.make_piece(w, h, data=ab.demosaic_data(ab.DEMOSAIC_AMAZE), devid=0)
def run(): ab.check(L.b200_demosaic_process_dev(C.byref(piece), m.data_ptr(), out.data_ptr(), s))
run(); torch.cuda.synchronize(); ts=[]
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print("AMaZE 45MP median ms", np.median(ts), "MP/s", w*h/np.median(ts)/1e3)
