"""Development timing of highlights' colour inpainting on a 45 MP mosaic (not the bench contract): time_inpaint.py [xtrans]"""
import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
import hl_laplacian_util as hu
ab.init()
w, h = util.SIZE_45MP
img = hu.clipped_mosaic(w, h, 45, blobs=9)
d = ab.highlights_data(ab.HIGHLIGHTS_INPAINT, 1.0)
xtrans = len(sys.argv) > 1 and sys.argv[1] == "xtrans"
piece = ab.make_piece(w, h, filters=9 if xtrans else util.BAYER["RGGB"], data=d, devid=0)
if xtrans:
    for k, v in enumerate(hu.XTRANS):
        piece.xtrans[k // 6][k % 6] = v
m = torch.from_numpy(img).cuda(); out = torch.empty_like(m)
s = torch.cuda.current_stream().cuda_stream
def run(): ab.check(ab.lib().b200_highlights_process_dev(piece, m.data_ptr(), out.data_ptr(), s))
run(); torch.cuda.synchronize(); ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
print(f"highlights colour inpainting 45MP {'X-Trans' if xtrans else 'Bayer'}: median ms {np.median(ts):.2f}  MP/s {w*h/np.median(ts)/1e3:.0f}")
