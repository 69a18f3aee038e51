"""Per-source-line totals of an ncu report captured with --import-source on: samples, instructions, stall reasons.
   python tools/ncu_lines.py rep.ncu-rep [file-substring] [top-n]"""
import csv, subprocess, sys, collections
def I(x):
    try: return int(x)
    except ValueError: return 0
rep = sys.argv[1]; want = sys.argv[2] if len(sys.argv) > 2 else ""; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur = None; hdr = None; out = []
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] in ("Function Name",) or hdr is None: continue
    if r[0] != "" and want in (cur or ""):
        d = dict(zip(hdr[4:], r[4:]))
        out.append((cur, int(r[0]), r[1], d))
tot_s = sum(I(o[3]["# Samples"]) for o in out); tot_i = sum(I(o[3]["Instructions Executed"]) for o in out)
print(f"total samples {tot_s} warp-instructions {tot_i}")
stalls = [k for k in out[0][3] if k.startswith("stall_") and "Not Issued" not in k]
for cur, ln, src, d in sorted(out, key=lambda o: -I(o[3]["# Samples"]))[:topn]:
    s = I(d["# Samples"]); top = sorted(((I(d[k]), k[6:]) for k in stalls), reverse=True)[:3]
    print(f"{cur.split('/')[-1]}:{ln:4d} samp {100*s/tot_s:5.1f}% inst {100*I(d['Instructions Executed'])/tot_i:5.1f}%  {' '.join(f'{k}={v}' for v,k in top)} | {src.strip()[:90]}")
