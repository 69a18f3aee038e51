#!/usr/bin/env python
"""bench.py -- the develop-pipe benchmark contract (see DESIGN.md "Measurement").

  python bench.py --gpus N --steps K --warmup W              # this repo's B200 path
  python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle/_ref)

Workload (BASELINE.json configs[1], SURVEY.md 8d "C2"): one synthetic 45.44 MP RGGB Bayer frame
(8256x5504, D-natural, seed 20260922) per step through  demosaic(RCD) -> colorin(matrix) ->
colorout(matrix + sRGB tone curve).  A "step" is one frame through that chain.
  value : device-resident chain throughput, MP/s, frames already in HBM, CUDA-event timed.
  e2e   : the same chain through the C module adapters and the device-resident pixelpipe glue
          (ansel_b200/iop/), host (pinned) buffers in and out, H2D + D2H inside the timed region.
Multi-GPU (torchrun): frames are independent, each rank develops its own frame per step, no
data-path collective (weak scaling, SURVEY.md 8e "batch (C5) replicas").
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
# The CPU arm uses whatever OpenMP placement the environment asks for (OMP_PROC_BIND / OMP_PLACES /
# OMP_NUM_THREADS); by default libgomp's own default: all host threads, unbound.
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import warnings  # noqa: E402

warnings.filterwarnings("ignore")  # numpy notices the FTZ/DAZ mode the reference sets on its threads

import numpy as np  # noqa: E402

METRIC = "megapixels/sec full develop pipe @45MP"
UNIT = "MP/s"
WORKLOAD = "C2: 45MP RGGB Bayer (8256x5504) -> demosaic(RCD) -> colorin(matrix) -> colorout(matrix+sRGB TRC)"
W45, H45 = 8256, 5504
SEED = 20260922
ALGO_BYTES_PER_PX = {"demosaic": 20, "colorin": 32, "colorout": 32}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=W45)
    ap.add_argument("--height", type=int, default=H45)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-modules", action="store_true", help="skip the untimed per-module table of the non-C2 modules")
    ap.add_argument("--pipe-ends-only", action="store_true",
                    help="internal: run the pipe-end module table in this process and print it as one JSON object (the main run calls this in a "
                         "child process so that code which has not been through a GPU round cannot take the headline down)")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------
# CPU arm: the reference's own sources compiled by oracle/Makefile (oracle/_ref), else the port
# ------------------------------------------------------------------------------------------
class CpuChain:
    def __init__(self, w, h):
        import util
        self.util = util
        self.w, self.h = w, h
        self.kind = "reference" if util.ref("fast") is not None else "port"
        self.enc = util.srgb_encode_lut()
        self.co_t = util.fit_unbounded_coeffs(self.enc)
        self.rgb = [util.aligned_empty((h, w, 4)) for _ in range(3)]
        self.fp = C.POINTER(C.c_float)

    def _conv(self, src, dst, matrix, lut_t=None, co_t=None):
        u = self.util
        m = np.ascontiguousarray(matrix, np.float32).reshape(-1)
        lt = u.fptr(lut_t) if lut_t is not None else None
        ct = u.fptr(np.ascontiguousarray(co_t, np.float32).reshape(-1)) if co_t is not None else None
        if self.kind == "reference":
            f = u.ref("fast").ref_apply_matrix_conversion
            f.restype = C.c_int
            f(u.fptr(src), u.fptr(dst), C.c_size_t(self.w), C.c_size_t(self.h), u.fptr(m), None, C.c_int(0), None, None, lt, ct)
        else:
            f = u.oracle().orc_apply_matrix_conversion
            f.restype = C.c_int
            f(u.fptr(src), u.fptr(dst), C.c_size_t(self.w), C.c_size_t(self.h), u.fptr(m), None, C.c_int(0), None, None, lt, ct,
              C.c_int(u.FP_CONTRACT))

    def step(self, mosaic):
        u = self.util
        pm = (C.c_float * 3)(1.0, 1.0, 1.0)
        lib = u.ref("fast") if self.kind == "reference" else u.oracle()
        name = "ref_rcd_demosaic" if self.kind == "reference" else "orc_rcd_demosaic"
        f = getattr(lib, name)
        f.restype = C.c_int
        f(u.fptr(self.rgb[0]), u.fptr(mosaic), self.w, self.h, C.c_uint32(u.BAYER["RGGB"]), pm, C.c_float(0.0))
        self._conv(self.rgb[0], self.rgb[1], u.MATRIX_CAM_TO_REC2020)
        self._conv(self.rgb[1], self.rgb[2], u.MATRIX_REC2020_TO_SRGB, self.enc, self.co_t)
        return self.rgb[2]


def tune_cpu_threads(chain, mosaic):
    """The reference's CPU path is OpenMP; give it the thread count it runs fastest with on this host (all logical
    CPUs, or one per physical core when SMT hurts this bandwidth-bound chain).  Returns the count chosen."""
    n_all = os.cpu_count() or 1
    if "OMP_NUM_THREADS" in os.environ:
        return int(os.environ["OMP_NUM_THREADS"])
    try:
        omp = C.CDLL("libgomp.so.1")
    except OSError:
        return n_all
    best_n, best_t = n_all, None
    for n in sorted({n_all, max(1, n_all // 2)}, reverse=True):
        omp.omp_set_num_threads(n)
        chain.step(mosaic)                       # first touch / thread start-up
        t0 = time.perf_counter()
        chain.step(mosaic)
        t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best_n, best_t = n, t
    omp.omp_set_num_threads(best_n)
    return best_n


def time_cpu(chain, mosaic, warm, steps):
    for _ in range(warm):
        chain.step(mosaic)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        chain.step(mosaic)
        ts.append(time.perf_counter() - t0)
    return ts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import util
    w, h = args.width, args.height
    mosaic = util.frame_natural(w, h, SEED)
    chain = CpuChain(w, h)
    cores = tune_cpu_threads(chain, mosaic)
    ts = time_cpu(chain, mosaic, max(args.warmup, 2), args.steps)
    total = float(np.sum(ts))
    mps = w * h * len(ts) / total / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frame": f"{w}x{h}", "cpu_threads": cores, "host_logical_cpus": os.cpu_count(),
                   "what": "reference sources compiled in place (oracle/_ref, release flags)" if chain.kind == "reference"
                   else "oracle port (oracle/_ref not built)",
                   "step_ms": [round(1e3 * t, 1) for t in ts]},
        "cpu_baseline": {"value": mps, "unit": UNIT, "cores": cores, "kind": chain.kind,
                         "sample": f"{len(ts)} full {w}x{h} frames through the chain, one per step"},
        "e2e": {"value": mps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# clocks: sample SM clock and throttle reasons during the timed region
# ------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), threading.Event()
        self.max_mhz = None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        nv = self.nv
        names = {"nvmlClocksEventReasonHwSlowdown": "hw_slowdown", "nvmlClocksEventReasonHwThermalSlowdown": "hw_thermal_slowdown",
                 "nvmlClocksEventReasonSwThermalSlowdown": "sw_thermal_slowdown", "nvmlClocksEventReasonSwPowerCap": "sw_power_cap",
                 "nvmlClocksThrottleReasonHwSlowdown": "hw_slowdown", "nvmlClocksThrottleReasonHwThermalSlowdown": "hw_thermal_slowdown",
                 "nvmlClocksThrottleReasonSwThermalSlowdown": "sw_thermal_slowdown", "nvmlClocksThrottleReasonSwPowerCap": "sw_power_cap"}
        bits = {getattr(nv, k): v for k, v in names.items() if hasattr(nv, k)}
        while not self.stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for b, name in bits.items():
                    if r & b:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------
def pipe_ends_extras(ab, ds, util, L, M, torch, w, h, local, dev, stream, conv_in, conv_out, d_cin, d_cout):
    """Per-module device times of rawprepare / temperature / highlights(clip) / their fused pass / exposure / gamma /
    finalscale at the bench frame size, and the sensor-to-display chain end to end: uint16 sensor data up (2 B/px), uint8
    display pixels down (4 B/px), rawprepare -> temperature -> highlights -> demosaic(RCD) -> colorin -> colorout -> gamma
    through the C module adapters with two frames in flight."""
    npx = w * h
    wb = (2.13, 1.0, 1.57, 1.02)
    pm = (wb[0], wb[1], wb[2], 0.0)
    filters = util.BAYER["RGGB"]
    sub, div = (512.0, 520.0, 508.0, 515.0), (15871.0, 15863.0, 15875.0, 15868.0)
    rng = np.random.default_rng(SEED)
    raw = np.clip(util.frame_natural(w, h, SEED) * 15871.0 * 0.45 + 512.0 + rng.normal(0, 3, (h, w)), 0, 16383).astype(np.uint16)
    raw[rng.integers(0, h, 4000), rng.integers(0, w, 4000)] = 16383           # blown samples: highlights takes the clip branch
    d_rp, d_tp, d_hl = ab.rawprepare_data(sub, div), ab.temperature_data(wb), ab.highlights_data(ab.HIGHLIGHTS_CLIP, 1.0)
    d_ex, d_fs, d_dem = ab.exposure_data(0.0, 0.5), ab.finalscale_data(ab.INTERPOLATION_MITCHELL), ab.demosaic_data(ab.DEMOSAIC_RCD)

    def piece(data, ch, datatype=ab.TYPE_FLOAT, pmax=(1.0, 1.0, 1.0, 1.0), out=None):
        p = ab.make_piece(w, h, filters=filters if ch == 1 else 0, channels=ch, data=data, processed_maximum=pmax, devid=local,
                          out_width=out[0] if out else None, out_height=out[1] if out else None)
        p.datatype = datatype
        return p

    p_rp, p_tp, p_hl = piece(d_rp, 1, ab.TYPE_UINT16), piece(d_tp, 1), piece(d_hl, 1, pmax=pm)
    p_ex, p_gm = piece(d_ex, 4), piece(None, 4)
    p_fs = piece(d_fs, 4, out=(w // 2, h // 2))
    p_fs.roi_out.scale = 0.5
    t_raw = torch.from_numpy(raw).to(dev)
    t_m = [torch.empty((h, w), dtype=torch.float32, device=dev) for _ in range(2)]
    t_rgba = torch.rand((h, w, 4), dtype=torch.float32, device=dev)
    t_out = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
    t_u8 = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)

    def timed(call, reps=5):
        call()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            call()
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        return float(np.median(ts))

    res = {}

    def report(label, call, bytes_per_px, reps=5):
        """time one entry point; a failure is recorded under its label and does not stop the table"""
        try:
            ms = timed(call, reps)
            res[label] = {"ms": ms, "MP_per_s": npx / ms / 1e3, "algorithmic_GBps": bytes_per_px * npx / (ms * 1e-3) / 1e9,
                          "algorithmic_bytes_per_px": bytes_per_px}
        except Exception as e:
            res[label] = {"unavailable": f"{type(e).__name__}: {e}"[:200]}

    report("rawprepare_uint16", lambda: ab.check(L.b200_rawprepare_process_dev(p_rp, t_raw.data_ptr(), t_m[0].data_ptr(), stream)), 6)
    report("temperature", lambda: ab.check(L.b200_temperature_process_dev(p_tp, t_m[0].data_ptr(), t_m[1].data_ptr(), stream)), 8)
    report("highlights_clip", lambda: ab.check(L.b200_highlights_process_dev(p_hl, t_m[1].data_ptr(), t_m[0].data_ptr(), stream)), 8)
    report("rawfront_fused_uint16", lambda: ab.check(L.b200_rawfront_process_dev(p_rp, p_tp, p_hl, t_raw.data_ptr(), t_m[0].data_ptr(), stream)), 6)
    report("exposure", lambda: ab.check(L.b200_exposure_process_dev(p_ex, t_rgba.data_ptr(), t_out.data_ptr(), stream)), 32)
    report("gamma_uint8", lambda: ab.check(L.b200_gamma_process_dev(p_gm, t_rgba.data_ptr(), t_u8.data_ptr(), stream)), 20)
    report("export_uint16", lambda: ab.check(L.b200_export_convert_dev(t_rgba.data_ptr(), t_out.data_ptr(), w, h, ab.EXPORT_UINT16, stream)), 24)
    report("finalscale_half_mitchell", lambda: ab.check(L.b200_finalscale_process_dev(p_fs, t_rgba.data_ptr(), t_out.data_ptr(), stream)), 20)

    # the demosaicers, colour calibration, the bilateral grid and highlight inpainting added with them
    t_dem = torch.empty((h, w, 4), dtype=torch.float32, device=dev)
    for label, method in (("demosaic_ppg_median", ab.DEMOSAIC_PPG), ("demosaic_vng4", ab.DEMOSAIC_VNG4), ("demosaic_rcd_dual_vng4", ab.DEMOSAIC_RCD | 2048)):
        dd = ab.demosaic_data(method)
        dd.median_thrs, dd.dual_thrs = 0.02, 0.2
        p_d = piece(dd, 1)
        report(label, lambda: ab.check(L.b200_demosaic_process_dev(p_d, t_m[0].data_ptr(), t_dem.data_ptr(), stream)), 20, reps=3)
    # the half-size method (algorithmic bytes per INPUT pixel: 4 in + 16 / 4 out), its post-filter, and the X-Trans variant
    hw, hh = (w + 1) // 2, (h + 1) // 2
    for label, smoothing, xtrans in (("demosaic_downsample", 0, False), ("demosaic_downsample_postfilter1", 1, False), ("demosaic_downsample_xtrans", 0, True)):
        dd = ab.demosaic_data(7)
        dd.color_smoothing = smoothing
        p_d = piece(dd, 1, out=(hw, hh))
        if xtrans:
            p_d.filters = 9
            for i, rowv in enumerate(((1, 1, 0, 1, 1, 2), (1, 1, 2, 1, 1, 0), (2, 0, 1, 0, 2, 1), (1, 1, 2, 1, 1, 0), (1, 1, 0, 1, 1, 2), (0, 2, 1, 2, 0, 1))):
                for j, v in enumerate(rowv):
                    p_d.xtrans[i][j] = v
        report(label, lambda: ab.check(L.b200_demosaic_process_dev(p_d, t_m[0].data_ptr(), t_dem.data_ptr(), stream)), 8, reps=3)
    cp = ab.channelmixer_piece(util.profile_pair(util.REC2020_TO_XYZ_D50), illuminant=(0.93, 1.02, 0.71))
    p_cm = piece(None, 4)
    p_cm.data, p_cm.data_size = C.addressof(cp), C.sizeof(cp)
    report("channelmixerrgb_cat16_v3", lambda: ab.check(L.b200_channelmixerrgb_process_dev(p_cm, t_rgba.data_ptr(), t_out.data_ptr(), stream)), 32)
    t_lab = t_rgba * torch.tensor([100.0, 60.0, 60.0, 1.0], device=dev)
    p_bl = piece(ab.bilat_data(sigma_r=5.0, sigma_s=50.0, detail=0.5, mode=0), 4)
    report("bilat_bilateral_grid_sigma50", lambda: ab.check(L.b200_bilat_process_dev(p_bl, t_lab.data_ptr(), t_out.data_ptr(), stream)), 32, reps=3)
    p_hi = piece(ab.highlights_data(ab.HIGHLIGHTS_INPAINT, 1.0), 1, pmax=pm)
    report("highlights_inpaint", lambda: ab.check(L.b200_highlights_process_dev(p_hi, t_m[1].data_ptr(), t_m[0].data_ptr(), stream)), 8, reps=3)

    # sensor-to-display chain, end to end
    try:
        order = [("rawprepare", d_rp, 1, 1, 2, 1, (1.0,) * 4), ("temperature", d_tp, 1, 1, 1, 1, (1.0,) * 4), ("highlights", d_hl, 1, 1, 1, 1, pm),
                 ("demosaic", d_dem, 1, 4, 1, 1, pm), ("colorin", d_cin, 4, 4, 1, 1, pm), ("colorout", d_cout, 4, 4, 1, 1, pm), ("gamma", None, 4, 4, 1, 3, pm)]
        pieces = [ds.make_piece_iop(op, w, h, data, channels_in=ci, channels_out=co, filters=filters, processed_maximum=pmx, wb=wb, type_in=ti, type_out=to)
                  for op, data, ci, co, ti, to, pmx in order]
        nodes = (ds.PipeNode * len(order))()
        for k, (op, *_rest) in enumerate(order):
            nodes[k].process_cl = C.cast(getattr(M, f"dt_iop_{op}__process_cl"), C.c_void_p)
            nodes[k].module = pieces[k].module
            nodes[k].piece = C.pointer(pieces[k])
        pipe = ds.make_pipe(devid=local, stream=None)
        DEPTH, steps = 2, 12
        queue = M.b200_pipe_queue_new(DEPTH)
        try:
            h_in = torch.from_numpy(raw).pin_memory()
            h_out = [torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]

            def run(n):
                tickets = []
                for i in range(n):
                    t = M.b200_pixelpipe_submit(queue, C.byref(pipe), nodes, len(order), h_in.data_ptr(), h_out[i % DEPTH].data_ptr())
                    if t < 0:
                        raise RuntimeError("chain failed: " + L.b200_last_error().decode())
                    tickets.append(t)
                    if i >= DEPTH - 1 and M.b200_pixelpipe_wait(queue, tickets[i - DEPTH + 1]) != 0:
                        raise RuntimeError("wait failed: " + L.b200_last_error().decode())
                if M.b200_pixelpipe_wait(queue, tickets[-1]) != 0:
                    raise RuntimeError("wait failed: " + L.b200_last_error().decode())

            run(3)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(steps)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        finally:
            M.b200_pipe_queue_free(queue)
        res["sensor_to_display_e2e"] = {"value": npx * steps / dt / 1e6, "unit": UNIT, "steps": steps, "h2d_bytes_per_step": 2 * npx, "d2h_bytes_per_step": 4 * npx,
                                        "chain": "rawprepare(uint16) -> temperature -> highlights(clip) -> demosaic(RCD) -> colorin -> colorout -> gamma(uint8)",
                                        "path": "dt_iop_<op>__process_cl adapters via b200_pixelpipe_submit/_wait (the raw front fused into one launch by the pipe glue), 2 frames in flight, pinned host buffers",
                                        "mean_display_value": float(h_out[0][..., :3].float().mean())}
    except Exception as e:
        res["sensor_to_display_e2e"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    return res


def pipe_ends_in_child(w, h, local):
    """pipe_ends_extras() in a child process with a time limit: its kernels were written after round 1's GPU budget was spent, so a
    crash or a hang there must not cost the benchmark line.  Reported, never fatal."""
    import subprocess
    cmd = [sys.executable, os.path.abspath(__file__), "--pipe-ends-only", "--width", str(w), "--height", str(h)]
    env = dict(os.environ, LOCAL_RANK=str(local), WORLD_SIZE="1", RANK="0")
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    except subprocess.TimeoutExpired:
        return {"unavailable": "child process exceeded 240 s"}
    for line in reversed(r.stdout.strip().splitlines()):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except Exception:
                break
    return {"unavailable": f"child exit {r.returncode}: {(r.stderr or r.stdout).strip()[-300:]}"}


def run_pipe_ends_only(args):
    import torch
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    import util
    local = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        if not torch.cuda.is_available():
            raise RuntimeError("no CUDA device")
        torch.cuda.set_device(local)
        ab.init()
        L, M = ab.lib(), ds.modlib()
        enc = util.srgb_encode_lut()
        co_t = np.zeros((3, 3), np.float32)
        L.b200_fit_unbounded_coeffs((C.c_void_p * 3)(*[enc[k].ctypes.data for k in range(3)]), co_t.ctypes.data)
        conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020, identity=0x2001)
        conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=co_t, identity=0x2002)
        res = pipe_ends_extras(ab, ds, util, L, M, torch, args.width, args.height, local, torch.device("cuda", local), torch.cuda.current_stream().cuda_stream,
                               conv_in, conv_out, ab.colorin_data(conv_in), ab.colorout_data(conv_out))
    except Exception as e:
        res = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    print(json.dumps(res))


def run_b200(args):
    import torch
    import torch.distributed as dist
    import ansel_b200 as ab
    import ansel_b200.dtsurface as ds
    import util

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ab.init()
    L = ab.lib()
    w, h = args.width, args.height
    npx = w * h

    mosaic = util.frame_natural(w, h, SEED + rank)
    filters = util.BAYER["RGGB"]
    enc = util.srgb_encode_lut()
    co_t = np.zeros((3, 3), np.float32)
    lut_ptrs = (C.c_void_p * 3)(*[enc[k].ctypes.data for k in range(3)])
    L.b200_fit_unbounded_coeffs(lut_ptrs, co_t.ctypes.data)
    conv_in = ab.make_conversion(util.MATRIX_CAM_TO_REC2020, identity=0x1001)
    conv_out = ab.make_conversion(util.MATRIX_REC2020_TO_SRGB, lut_target=enc, coeffs_target=co_t, identity=0x1002)
    d_dem, d_cin, d_cout = ab.demosaic_data(ab.DEMOSAIC_RCD), ab.colorin_data(conv_in), ab.colorout_data(conv_out)
    p_dem = ab.make_piece(w, h, filters=filters, data=d_dem, devid=local)
    p_cin = ab.make_piece(w, h, filters=0, channels=4, data=d_cin, devid=local)
    p_cout = ab.make_piece(w, h, filters=0, channels=4, data=d_cout, devid=local)

    dev = torch.device("cuda", local)
    t_mosaic = torch.from_numpy(mosaic).to(dev)
    t_rgb = [torch.empty((h, w, 4), dtype=torch.float32, device=dev) for _ in range(3)]
    stream = torch.cuda.current_stream().cuda_stream
    mod_ms = {"demosaic": [], "colorin": [], "colorout": []}
    launches = [0]

    def step(record=False):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if record else None
        if record:
            evs[0].record()
        ab.check(L.b200_demosaic_process_dev(p_dem, t_mosaic.data_ptr(), t_rgb[0].data_ptr(), stream))
        if record:
            evs[1].record()
        ab.check(L.b200_colorin_process_dev(p_cin, t_rgb[0].data_ptr(), t_rgb[1].data_ptr(), stream))
        if record:
            evs[2].record()
        ab.check(L.b200_colorout_process_dev(p_cout, t_rgb[1].data_ptr(), t_rgb[2].data_ptr(), stream))
        if record:
            evs[3].record()
        launches[0] += 4  # rcd_ring_kernel, rcd_tiles_kernel, convert_kernel x2
        return evs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: exactly K steps, device-resident ----------------------------------------
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        e0.record()
        all_evs = [step(record=True) for _ in range(args.steps)]
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    for evs in all_evs:
        mod_ms["demosaic"].append(evs[0].elapsed_time(evs[1]))
        mod_ms["colorin"].append(evs[1].elapsed_time(evs[2]))
        mod_ms["colorout"].append(evs[2].elapsed_time(evs[3]))
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * npx * args.steps / (ms_max * 1e-3) / 1e6
    gpu_launches = launches[0]

    # ---- e2e: host buffers through the C module adapters + device-resident pipe glue -----------
    M = ds.modlib()
    pipe = ds.make_pipe(devid=local, stream=None)
    nodes = (ds.PipeNode * 3)()
    pieces = [ds.make_piece_iop("demosaic", w, h, d_dem, channels_in=1, channels_out=4, filters=filters),
              ds.make_piece_iop("colorin", w, h, d_cin, channels_in=4, channels_out=4),
              ds.make_piece_iop("colorout", w, h, d_cout, channels_in=4, channels_out=4)]
    for k, op in enumerate(("demosaic", "colorin", "colorout")):
        nodes[k].process_cl = C.cast(getattr(M, f"dt_iop_{op}__process_cl"), C.c_void_p)
        nodes[k].module = pieces[k].module
        nodes[k].piece = C.pointer(pieces[k])
    # two frames in flight (b200_pixelpipe_submit/_wait): the read-back of frame n overlaps upload + compute of
    # frame n+1.  Every step still uploads its own input and reads its own result back inside the timed region.
    DEPTH = 2
    queue = M.b200_pipe_queue_new(DEPTH)
    h_in = torch.from_numpy(mosaic).pin_memory()
    h_out = [torch.empty((h, w, 4), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]

    def e2e_run(n):
        tickets = []
        for i in range(n):
            t = M.b200_pixelpipe_submit(queue, C.byref(pipe), nodes, 3, h_in.data_ptr(), h_out[i % DEPTH].data_ptr())
            if t < 0:
                raise RuntimeError("e2e chain failed: " + L.b200_last_error().decode())
            tickets.append(t)
            if i >= DEPTH - 1:                      # the consumer takes frame i-DEPTH+1 before its buffer is reused
                if M.b200_pixelpipe_wait(queue, tickets[i - DEPTH + 1]) != 0:
                    raise RuntimeError("e2e wait failed: " + L.b200_last_error().decode())
        for t in tickets[-(DEPTH - 1):] if DEPTH > 1 else []:
            if M.b200_pixelpipe_wait(queue, t) != 0:
                raise RuntimeError("e2e wait failed: " + L.b200_last_error().decode())

    e2e_run(3)
    barrier()
    t0 = time.perf_counter()
    e2e_run(args.e2e_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    # the synchronous single-frame call, for reference (one frame at a time: upload, chain, read back)
    bufs = M.b200_pipe_buffers_new()
    for _ in range(2):
        M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, 3, bufs, h_in.data_ptr(), h_out[0].data_ptr())
    t1 = time.perf_counter()
    for _ in range(4):
        if M.b200_pixelpipe_process_on_gpu(C.byref(pipe), nodes, 3, bufs, h_in.data_ptr(), h_out[0].data_ptr()) != 0:
            raise RuntimeError("e2e chain failed: " + L.b200_last_error().decode())
    e2e_sync_ms = (time.perf_counter() - t1) / 4 * 1e3
    M.b200_pipe_buffers_free(bufs)
    t_e = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * npx * args.e2e_steps / float(t_e.item()) / 1e6
    M.b200_pipe_queue_free(queue)

    # ---- the other modules of SURVEY.md 8a at the same frame size (outside the timed region; N=1 only) ----
    other = None
    if world == 1 and not args.no_other_modules:
        other = {}
        work, export = util.profile_pair(util.REC2020_TO_XYZ_D50), util.profile_pair(util.SRGB_TO_XYZ_D50)
        fblob = np.load(os.path.join(util.GOLDEN_DIR, "filmic_data.npz"))["default_v8"]
        fp = ab.filmic_piece(fblob, work, export)
        cases = [("denoiseprofile_wavelets", "denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_WAVELETS), 2),
                 ("denoiseprofile_nlmeans_K7_P1", "denoiseprofile", ab.denoiseprofile_data(ab.DENOISE_NLMEANS, radius=1, nbhood=7), 1),
                 ("filmicrgb_v8_agx", "filmicrgb", fp, 3),
                 ("diffuse_sharpen_demosaic_1it_5scales", "diffuse", ab.diffuse_data(**ab.DIFFUSE_PRESETS["sharpen_demosaic_aa"]), 2),
                 ("bilat_local_laplacian", "bilat", ab.bilat_data(), 2)]
        t_rgb[0].copy_(torch.rand((h, w, 4), device=dev))
        for label, op, data, reps in cases:
            pc = ab.make_piece(w, h, filters=0, channels=4, devid=local)
            pc.data, pc.data_size = C.addressof(data), C.sizeof(data)
            fn = getattr(L, f"b200_{op}_process_dev")
            ab.check(fn(pc, t_rgb[0].data_ptr(), t_rgb[1].data_ptr(), stream))
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                ab.check(fn(pc, t_rgb[0].data_ptr(), t_rgb[1].data_ptr(), stream))
                a1.record()
                torch.cuda.synchronize()
                ts.append(a0.elapsed_time(a1))
            m = float(np.median(ts))
            other[label] = {"ms": m, "MP_per_s": npx / m / 1e3, "algorithmic_GBps": 32 * npx / (m * 1e-3) / 1e9}
        # the second demosaicer of the north star on the bench frame (20 B/px at the module boundary)
        d_amz = ab.demosaic_data(ab.DEMOSAIC_AMAZE)
        p_amz = ab.make_piece(w, h, filters=filters, data=d_amz, devid=local)
        ts = []
        for k in range(4):
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            ab.check(L.b200_demosaic_process_dev(p_amz, t_mosaic.data_ptr(), t_rgb[1].data_ptr(), stream))
            a1.record()
            torch.cuda.synchronize()
            if k:
                ts.append(a0.elapsed_time(a1))
        m = float(np.median(ts))
        other["demosaic_amaze"] = {"ms": m, "MP_per_s": npx / m / 1e3, "algorithmic_GBps": 20 * npx / (m * 1e-3) / 1e9}

        # the modules either side of the path and the sensor-to-display chain; never part of the headline
        other["pipe_ends"] = pipe_ends_in_child(w, h, local)

    # ---- second mode at N>1: ONE frame cut into row bands + all-gather ---------
    banded = None
    if world > 1:
        from ansel_b200 import bands
        bnodes = [bands.Node("demosaic", d_dem, channels_in=1), bands.Node("colorin", d_cin), bands.Node("colorout", d_cout)]
        frame0 = util.frame_natural(w, h, SEED)          # the same frame on every rank
        ch = bands.BandedChain(bnodes, w, h, rank, world, device=dev)
        t_band = torch.from_numpy(np.ascontiguousarray(ch.band_rows(frame0))).to(dev)
        for _ in range(3):
            ch(t_band, stream=stream)
        barrier()
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        for _ in range(args.steps):
            out_frame = ch(t_band, stream=stream)
        b1.record()
        barrier()
        t_b = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t_b, op=dist.ReduceOp.MAX)
        # every rank must now hold the frame rank 0 computes untiled
        same = torch.ones(1, dtype=torch.int32, device=dev)
        if rank == 0:
            t_full = torch.from_numpy(frame0).to(dev)
            ab.check(L.b200_demosaic_process_dev(p_dem, t_full.data_ptr(), t_rgb[0].data_ptr(), stream))
            ab.check(L.b200_colorin_process_dev(p_cin, t_rgb[0].data_ptr(), t_rgb[1].data_ptr(), stream))
            ab.check(L.b200_colorout_process_dev(p_cout, t_rgb[1].data_ptr(), t_rgb[2].data_ptr(), stream))
            same[0] = int(torch.equal(out_frame.view(torch.int32), t_rgb[2].view(torch.int32)))
        dist.broadcast(same, 0)
        banded = {"value": npx * args.steps / (float(t_b.item()) * 1e-3) / 1e6, "unit": UNIT, "ms_per_frame": float(t_b.item()) / args.steps,
                  "collective": "all-gather of finished RGBA bands (one ncclBroadcast per band, NVLink)",
                  "cuts": "RCD 94-row block grid, 9-row halo", "bit_identical_to_untiled": bool(same.item()),
                  "gathered_bytes_per_frame": 16 * npx}
        # the same, with the gather fused into colorout's stores (peer frames mapped over CUDA IPC)
        try:
            ch2 = bands.BandedChain(bnodes, w, h, rank, world, device=dev, p2p=True)
            for _ in range(3):
                ch2(t_band, stream=stream)
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(args.steps):
                out2 = ch2(t_band, stream=stream)
            c1.record()
            barrier()
            t_c = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t_c, op=dist.ReduceOp.MAX)
            same2 = torch.tensor([int(torch.equal(out2.view(torch.int32), out_frame.view(torch.int32)))], dtype=torch.int32, device=dev)
            dist.all_reduce(same2, op=dist.ReduceOp.MIN)
            ch2.close()
            # gather to the exporting rank only (what an export needs): every band has one destination
            ch3 = bands.BandedChain(bnodes, w, h, rank, world, device=dev, p2p=True, p2p_dst=0)
            for _ in range(3):
                ch3(t_band, stream=stream)
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(args.steps):
                out3 = ch3(t_band, stream=stream)
            g1.record()
            barrier()
            t_g = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(t_g, op=dist.ReduceOp.MAX)
            same3 = torch.ones(1, dtype=torch.int32, device=dev)
            if rank == 0:
                same3[0] = int(torch.equal(out3.view(torch.int32), out_frame.view(torch.int32)))
            dist.broadcast(same3, 0)
            ch3.close()
            banded["fused_p2p_gather_to_rank0"] = {"value": npx * args.steps / (float(t_g.item()) * 1e-3) / 1e6,
                                                   "ms_per_frame": float(t_g.item()) / args.steps, "bit_identical_to_collective": bool(same3.item())}
            banded["fused_p2p"] = {"value": npx * args.steps / (float(t_c.item()) * 1e-3) / 1e6, "ms_per_frame": float(t_c.item()) / args.steps,
                                   "how": "colorout stores each pixel into every rank's frame (CUDA IPC peer mappings), then one barrier",
                                   "bit_identical_to_collective": bool(same2.item())}
        except Exception as e:  # no peer access on this box: keep the NCCL number
            banded["fused_p2p"] = {"unavailable": str(e)[:200]}

    # ---- roofline of the dominant kernel (RCD tiles) --------------------------------------------
    peak, peak_src = peaks()
    dem_ms = float(np.mean(mod_ms["demosaic"]))
    achieved = ALGO_BYTES_PER_PX["demosaic"] * npx / (dem_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("rcd_tiles_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None
    per_module = {k: {"ms": float(np.mean(v)), "algorithmic_GBps": ALGO_BYTES_PER_PX[k] * npx / (float(np.mean(v)) * 1e-3) / 1e9,
                      "frac_of_hbm_peak": ALGO_BYTES_PER_PX[k] * npx / (float(np.mean(v)) * 1e-3) / 1e9 / peak}
                  for k, v in mod_ms.items()}

    # ---- CPU baseline, rank 0 at N=1 only, bounded sample ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        chain = CpuChain(w, h)
        cores = tune_cpu_threads(chain, mosaic)
        ts = time_cpu(chain, mosaic, 1, 4)
        cpu = {"value": w * h / float(np.median(ts)) / 1e6, "unit": UNIT, "cores": cores, "kind": chain.kind,
               "sample": f"4 full {w}x{h} frames through the same chain after 1 warm-up (median), "
                         f"{'reference sources, release flags, ' if chain.kind == 'reference' else 'oracle port, '}OpenMP, thread count tuned (all logical CPUs or one per core, whichever is faster)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frame": f"{w}x{h}", "frames_per_step": world,
                       "parallelism": f"{world} independent frame replicas, no data-path collective",
                       "l2": "inputs larger than L2: 182 MB mosaic + 3 x 727 MB RGBA per step vs 126 MB L2",
                       "fp": "RCD: C-standard float semantics (no contraction); colour: reference release-build contraction",
                       "per_module": per_module},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 4 * npx, "d2h_bytes_per_step": 16 * npx,
                    "steps": args.e2e_steps,
                    "path": "dt_iop_<op>__process_cl adapters via b200_pixelpipe_submit/_wait, 2 frames in flight, pinned host buffers",
                    "single_frame_sync_ms": e2e_sync_ms},
            "gpu_launches": gpu_launches,
            "clocks": clk.summary(),
            "roofline": {"bound": "hbm", "kernel": "rcd_tiles_kernel (+rcd_ring_kernel, <1% of the pair)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX["demosaic"] * npx,
                         "avg_launch_ms": dem_ms},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if other is not None:
            line["config"]["other_modules_45mp"] = other
        if banded is not None:
            line["banded_one_frame"] = banded
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.pipe_ends_only:
        run_pipe_ends_only(a)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
