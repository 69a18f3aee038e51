"""One launch of every module at 45 MP (NLM on a 3 MP frame: 225 patches per chunk are slow under replay), for
   ncu --set full -k regex:'rcd_tiles|convert_kernel|heat_pde|bspline|filmic_agx|eaw_|vst_|ll_|nlm_chunks|lab_kernel' ..."""
import sys, ctypes as C, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests")); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, util, ansel_b200 as ab
ab.init()
L = ab.lib(); s = torch.cuda.current_stream().cuda_stream
w, h = util.SIZE_45MP
mosaic = torch.from_numpy(util.frame_natural(w, h, util.SEEDS[0])).cuda()
a = torch.empty((h, w, 4), device="cuda"); b = torch.empty_like(a)
def run(op, data, src, dst, ww=w, hh=h, ch=4, filters=0):
    pc = ab.make_piece(ww, hh, filters=filters, channels=ch, devid=0)
    pc.data, pc.data_size = C.addressof(data), C.sizeof(data)
    ab.check(getattr(L, f"b200_{op}_process_dev")(C.byref(pc), src.data_ptr(), dst.data_ptr(), s))
    torch.cuda.synchronize()
enc = util.srgb_encode_lut()
conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020)
conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=util.fit_unbounded_coeffs(enc))
run("demosaic", ab.demosaic_data(ab.DEMOSAIC_AMAZE), mosaic, a, ch=1, filters=util.BAYER["RGGB"])
run("demosaic", ab.demosaic_data(ab.DEMOSAIC_RCD), mosaic, a, ch=1, filters=util.BAYER["RGGB"])
run("colorin", ab.colorin_data(conv_in), a, b)
run("denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_WAVELETS), b, a)
run("diffuse", ab.diffuse_data(**ab.DIFFUSE_PRESETS["sharpen_demosaic_aa"]), a, b)
work, export = util.profile_pair(util.REC2020_TO_XYZ_D50), util.profile_pair(util.SRGB_TO_XYZ_D50)
fp = ab.filmic_piece(np.load(os.path.join(util.GOLDEN_DIR, "filmic_data.npz"))["default_v8"], work, export)
run("filmicrgb", fp, b, a)
pm = ab.profile_matrices(*work)
f = L.b200_colorspace_transform_dev
f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ab.ProfileMatrices), C.c_int, C.c_void_p]
ab.check(f(a.data_ptr(), a.data_ptr(), w, h, ab.CS_RGB, ab.CS_LAB, C.byref(pm), 0, s))
run("bilat", ab.bilat_data(), a, b)
ab.check(f(b.data_ptr(), b.data_ptr(), w, h, ab.CS_LAB, ab.CS_RGB, C.byref(pm), 0, s))
run("colorout", ab.colorout_data(conv_out), b, a)
if "--nlm" in sys.argv:
    ws, hs = 2048, 1536
    sm = torch.rand((hs, ws, 4), device="cuda") * 20; so = torch.empty_like(sm)
    run("denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=7), sm, so, ws, hs)
print("done")
