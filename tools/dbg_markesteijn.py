"""Where the CUDA Markesteijn differs from the oracle (development aid)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab, markesteijn_util as mu
from vng_util import XTRANS
ab.init()
names = sys.argv[1:] or ["small", "one_tile", "roi"]
for name in names:
    m, x, y = mu.case(name)
    h, w = m.shape
    d = ab.demosaic_data(1025)
    piece = ab.make_piece(w, h, filters=9, data=d, devid=0, roi_x=x, roi_y=y)
    for i in range(6):
        for j in range(6):
            piece.xtrans[i][j] = int(XTRANS[i][j])
    d_in = torch.from_numpy(np.ascontiguousarray(m)).cuda()
    d_out = torch.full((h, w, 4), -7.0, device="cuda")
    ab.check(ab.lib().b200_demosaic_process_dev(C.byref(piece), d_in.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    got = d_out.cpu().numpy(); want = mu.oracle(m, x, y, 1)
    bad = got[..., :3].view(np.uint32) != want[..., :3].view(np.uint32)
    print(name, (w, h), "lane3 kept", bool((got[..., 3] == -7.0).all()), "differing floats", int(bad.sum()), "of", bad.size,
          "max abs", float(np.nanmax(np.abs(got[..., :3] - want[..., :3]))), "nan got", int(np.isnan(got[..., :3]).sum()))
    idx = np.argwhere(bad)
    if len(idx):
        print("  first", idx[:6].tolist(), "rows", idx[:, 0].min(), idx[:, 0].max(), "cols", idx[:, 1].min(), idx[:, 1].max())
        for r, c, k in idx[:4]:
            print("   ", (r, c, k), got[r, c, k], want[r, c, k])
