"""Development helper: build libb200iop.so variants of ONE kernel file with different tuning macros and time each.
    python tools/variants.py build nlm          (here, cross-compiling)
    python tools/variants.py time nlm           (on the GPU box)"""
import os, subprocess, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ansel_b200 import build as B
SETS = {
    "rcd": ("rcd.cu", "quick_time.py", {"rg8_u1": ["-DRCD_RG=8", "-DRCD_UNROLL=1"], "rg8_u2": ["-DRCD_RG=8", "-DRCD_UNROLL=2"], "rg7_u1": ["-DRCD_RG=7"]}),
    "nlm": ("nlm.cu", "time_nlm.py", {'nt512_2224': ['-DNLM_NT=512', '-DNLM_UB=2', '-DNLM_UE=2', '-DNLM_UA=2', '-DNLM_UC=4'], 'nt384_3448': ['-DNLM_NT=384', '-DNLM_UB=3', '-DNLM_UE=4', '-DNLM_UA=4', '-DNLM_UC=8'], 'nt384_2224': ['-DNLM_NT=384', '-DNLM_UB=2', '-DNLM_UE=2', '-DNLM_UA=2', '-DNLM_UC=4'], 'nt256_4448': ['-DNLM_NT=256', '-DNLM_UB=4', '-DNLM_UE=4', '-DNLM_UA=4', '-DNLM_UC=8'], 'nt512_1112': ['-DNLM_NT=512', '-DNLM_UB=1', '-DNLM_UE=1', '-DNLM_UA=1', '-DNLM_UC=2']}),
    "amaze": ("amaze.cu", "time_amaze_gpu.py", {"nt640_b1": [], "nt960_b1": ["-DAMAZE_NT=960"], "nt640_b2": ["-DAMAZE_MINB=2"], "nt800_b1": ["-DAMAZE_NT=800"]}),
}
OUT = os.path.join(ROOT, "tools", "variants")
os.makedirs(OUT, exist_ok=True)
src, timer, variants = SETS[sys.argv[2]]
stem = src[:-3]
if sys.argv[1] == "build":
    B.build()
    objs = [o for o in glob.glob(os.path.join(ROOT, "ansel_b200", "build", "*.o")) if not o.endswith(f"/{stem}.o")]
    for name, defs in variants.items():
        obj = os.path.join(OUT, f"{stem}_{name}.o")
        cmd = [B._nvcc()] + [f for f in B.NVCC_FLAGS if f != "-shared"] + defs + ["-Xptxas=-v", "-c", os.path.join(B.CSRC, src), "-o", obj]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, check=True)
        print(name, [l.strip() for l in r.stdout.splitlines() if "registers" in l or "spill" in l][-2:])
        subprocess.run([B._nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o",
                        os.path.join(OUT, f"libb200iop_{stem}_{name}.so"), obj] + objs, check=True)
else:
    for name in variants:
        env = dict(os.environ, B200IOP_LIB=os.path.join(OUT, f"libb200iop_{stem}_{name}.so"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", timer)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print(name, r.stdout.strip().splitlines()[-1])
