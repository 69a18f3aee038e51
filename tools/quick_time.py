"""Scratch timing helper for development runs on the GPU box (not the bench contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import ansel_b200 as ab, util

ab.init()
w, h = util.SIZE_45MP if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
m = util.frame_uniform(w, h, 1)
d_in = torch.from_numpy(m).cuda()
d_out = torch.empty((h, w, 4), dtype=torch.float32, device="cuda")
data = ab.demosaic_data(ab.DEMOSAIC_RCD)
piece = ab.make_piece(w, h, data=data, devid=0)
s = torch.cuda.current_stream().cuda_stream
L = ab.lib()
for _ in range(3):
    ab.check(L.b200_demosaic_process_dev(piece, d_in.data_ptr(), d_out.data_ptr(), s))
torch.cuda.synchronize()
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ab.check(L.b200_demosaic_process_dev(piece, d_in.data_ptr(), d_out.data_ptr(), s))
    e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = np.array(ts)
mp = w * h / 1e6
print(f"RCD {w}x{h}: median {np.median(ts):.3f} ms  min {ts.min():.3f} ms  -> {mp/np.median(ts)*1e3/1e3:.2f} GP/s, "
      f"{20*w*h/np.median(ts)/1e6:.1f} GB/s algorithmic")
