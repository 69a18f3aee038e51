"""Development helper: per-section cycles of the AMaZE tile kernel (CTA 0), from a -DB200_AMAZE_PROF build.
    python tools/prof_amaze.py build      (here)
    python tools/prof_amaze.py run        (on the GPU box)"""
import os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tools", "variants"); os.makedirs(OUT, exist_ok=True)
LIB = os.path.join(OUT, "libb200iop_amaze_prof.so")
NAMES = ["clear scratch", "tile load", "gradients", "colour differences", "hcd rows (serial)", "vcd columns (serial) + cddiffsq", "adaptive weight", "nyquist test + area", "hvwt rows (serial)",
         "green at R/B", "nyquist refinement", "diagonal gradients", "diagonal interpolation", "pmwt rows (serial)", "rbint", "green re-interpolation", "chroma split", "chroma interpolation", "output"]
if sys.argv[1] == "build":
    from ansel_b200 import build as B
    B.build()
    objs = [o for o in glob.glob(os.path.join(ROOT, "ansel_b200", "build", "*.o")) if not o.endswith("/amaze.o")]
    obj = os.path.join(OUT, "amaze_prof.o")
    subprocess.run([B._nvcc()] + [f for f in B.NVCC_FLAGS if f != "-shared"] + ["-DB200_AMAZE_PROF"] + sys.argv[2:] + ["-c", os.path.join(B.CSRC, "amaze.cu"), "-o", obj], check=True)
    subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o", LIB, obj] + objs, check=True)
    print("built", LIB)
else:
    os.environ["B200IOP_LIB"] = LIB
    import ctypes as C, numpy as np, torch, util, ansel_b200 as ab
    ab.init()
    w, h = util.SIZE_45MP
    m = torch.from_numpy(util.frame_natural(w, h, 1)).cuda(); out = torch.empty((h, w, 4), device="cuda")
    L = ab.lib(); s = torch.cuda.current_stream().cuda_stream
    piece = ab.make_piece(w, h, data=ab.demosaic_data(ab.DEMOSAIC_AMAZE), devid=0)
    run = lambda: ab.check(L.b200_demosaic_process_dev(C.byref(piece), m.data_ptr(), out.data_ptr(), s))
    run(); torch.cuda.synchronize()
    buf = (C.c_ulonglong * 40)()
    L.b200_amaze_prof(buf, 40, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    L.b200_amaze_prof(buf, 40, 0)
    v = np.array(list(buf)[:len(NAMES)], dtype=np.float64); tot = v.sum()
    print("kernel ms", e0.elapsed_time(e1), "CTA0 cycles", tot)
    for n, x in zip(NAMES, v): print(f"{n:36s} {x/1e3:10.1f} kcycles {100*x/tot:5.1f}%")
