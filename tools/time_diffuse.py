import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
ab.init()
w, h = util.SIZE_45MP
img = torch.rand((h, w, 4), device="cuda")
out = torch.empty_like(img)
L = ab.lib(); s = torch.cuda.current_stream().cuda_stream
for name in ("sharpen_demosaic_aa", "lens_deblur_soft"):
    data = ab.diffuse_data(**ab.DIFFUSE_PRESETS[name])
    piece = ab.make_piece(w, h, filters=0, channels=4, data=data)
    def run(): ab.check(L.b200_diffuse_process_dev(C.byref(piece), img.data_ptr(), out.data_ptr(), s))
    run(); torch.cuda.synchronize(); ts=[]
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"diffuse {name} iterations={data.iterations} 45MP median ms", np.median(ts), "MP/s", w*h/np.median(ts)/1e3)
