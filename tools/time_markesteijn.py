"""Development timing of the Markesteijn demosaicer at 45 MP (not the bench contract)."""
import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
from vng_util import XTRANS
ab.init()
w, h = util.SIZE_45MP
m = torch.from_numpy(util.frame_natural(w, h, 3)).cuda()
out = torch.empty((h, w, 4), device="cuda")
passes = int(sys.argv[1]) if len(sys.argv) > 1 else 1
d = ab.demosaic_data(1025 if passes == 1 else 1026)
piece = ab.make_piece(w, h, filters=9, data=d, devid=0)
for i in range(6):
    for j in range(6):
        piece.xtrans[i][j] = int(XTRANS[i][j])
s = torch.cuda.current_stream().cuda_stream
def run(): ab.check(ab.lib().b200_demosaic_process_dev(C.byref(piece), m.data_ptr(), out.data_ptr(), s))
run(); torch.cuda.synchronize(); ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"Markesteijn {passes}-pass 45MP median ms {np.median(ts):.2f}  MP/s {w*h/np.median(ts)/1e3:.0f}  ({20*w*h/np.median(ts)/1e6:.0f} GB/s algorithmic)")
