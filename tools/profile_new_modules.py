"""One launch at 45 MP of the modules round 2 added after the first ncu table (Markesteijn 1 pass, LMMSE + median, blending, highlights'
guided laplacians with 2 iterations), for
   ncu --set full --clock-control none -k regex:'markesteijn_tiles|lmmse_tiles|blend_kernel|hl_|bspline' -c 80 ..."""
import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
import hl_laplacian_util as hu, blend_util as bu
ab.init()
L = ab.lib(); s = torch.cuda.current_stream().cuda_stream
w, h = util.SIZE_45MP
mosaic = torch.from_numpy(util.frame_natural(w, h, 3)).cuda()
rgba = torch.empty((h, w, 4), device="cuda")
# Markesteijn, one pass, X-Trans
d = ab.demosaic_data(1025)
piece = ab.make_piece(w, h, filters=9, data=d, devid=0)
for k, v in enumerate(hu.XTRANS):
    piece.xtrans[k // 6][k % 6] = v
ab.check(L.b200_demosaic_process_dev(C.byref(piece), mosaic.data_ptr(), rgba.data_ptr(), s)); torch.cuda.synchronize()
# LMMSE with one median pass
d = ab.demosaic_data(6); d.lmmse_refine = 1
piece = ab.make_piece(w, h, filters=util.BAYER["RGGB"], data=d, devid=0)
ab.check(L.b200_demosaic_process_dev(C.byref(piece), mosaic.data_ptr(), rgba.data_ptr(), s)); torch.cuda.synchronize()
# blending: drawn and parametric mask, normal operator
other = torch.rand((h, w, 4), device="cuda")
form = torch.rand((h, w), device="cuda")
mask = torch.empty((h, w), device="cuda")
p = bu.params(mode="normal", opacity=70.0, mask_mode=bu.MASK_ENABLED | bu.MASK_SHAPE | bu.MASK_PARAMETRIC, drawn=1,
              channels={0: (0.05, 0.2, 0.8, 1.0), 5: (0.0, 0.0, 0.7, 0.9)})
L.b200_blend_process_dev.argtypes = [C.c_void_p] * 7
piece = ab.make_piece(w, h, channels=4, devid=0)
ab.check(L.b200_blend_process_dev(C.byref(piece), C.byref(p), other.data_ptr(), rgba.data_ptr(), form.data_ptr(), mask.data_ptr(), s)); torch.cuda.synchronize()
# highlights, guided laplacians, 2 iterations
img = hu.clipped_mosaic(w, h, 45, blobs=9)
piece, d = hu.piece_of(ab, img, util.BAYER["RGGB"], iterations=2, scales=8)
m = torch.from_numpy(img).cuda(); out = torch.empty_like(m)
ab.check(L.b200_highlights_process_dev(piece, m.data_ptr(), out.data_ptr(), s)); torch.cuda.synchronize()
print("done")
