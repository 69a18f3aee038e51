"""Development helper: build libb200iop.so variants with different RCD tuning macros and time each on the GPU.
    python tools/rcd_variants.py build      (here, cross-compiling)
    python tools/rcd_variants.py time       (on the GPU box)"""
import os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ansel_b200 import build as B
VARIANTS = {"rg8_u1": ["-DRCD_RG=8", "-DRCD_UNROLL=1"], "rg9_u1": ["-DRCD_RG=9", "-DRCD_UNROLL=1"], "rg8_u2": ["-DRCD_RG=8", "-DRCD_UNROLL=2"],
            "rg4_u2": ["-DRCD_RG=4", "-DRCD_UNROLL=2"], "rg6_u2": ["-DRCD_RG=6", "-DRCD_UNROLL=2"], "rg4_u1": ["-DRCD_RG=4", "-DRCD_UNROLL=1"]}
OUT = os.path.join(ROOT, "tools", "variants")
if sys.argv[1] == "build":
    B.build()
    objs = [o for o in glob.glob(os.path.join(ROOT, "ansel_b200", "build", "*.o")) if not o.endswith("/rcd.o")]
    for name, defs in VARIANTS.items():
        obj = os.path.join(OUT, f"rcd_{name}.o")
        cmd = [B._nvcc()] + [f for f in B.NVCC_FLAGS if f != "-shared"] + defs + ["-Xptxas=-v", "-c", os.path.join(B.CSRC, "rcd.cu"), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
        print(name, [l for l in r.stdout.splitlines() if "registers" in l][:2])
        subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o",
                        os.path.join(OUT, f"libb200iop_{name}.so"), obj] + objs, check=True)
else:
    for name in VARIANTS:
        env = dict(os.environ, B200IOP_LIB=os.path.join(OUT, f"libb200iop_{name}.so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "quick_time.py")], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(name, r.stdout.strip().splitlines()[-1])
