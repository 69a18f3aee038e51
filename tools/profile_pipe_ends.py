"""One launch of every kernel added after round 1's GPU budget was spent, at 45 MP, for
   ncu --set full --clock-control none --import-source on -k regex:'raw_front|flat_kernel|gamma_kernel|export_kernel|resample_kernel|channelmixer_kernel|ppg_kernel|pre_median|vng_kernel|lin_interpolate|inpaint_|detail_' ...
Not a benchmark: bench.py reports the timings (config.other_modules_45mp.pipe_ends)."""
import ctypes as C
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "..", "tests"), os.path.join(HERE, "..")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
import ansel_b200 as ab  # noqa: E402
import pipe_ends_util as pe  # noqa: E402
import util  # noqa: E402

ab.init()
L = ab.lib()
s = torch.cuda.current_stream().cuda_stream
w, h = util.SIZE_45MP
wb = (2.13, 1.0, 1.57, 1.02)
raw = torch.from_numpy(pe.sensor_frame(w, h, 3, clipped=4000)).cuda()
m0 = torch.empty((h, w), device="cuda")
m1 = torch.empty_like(m0)
a = torch.rand((h, w, 4), device="cuda")
b = torch.empty_like(a)
u8 = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda")


def piece(data, ch, datatype=ab.TYPE_FLOAT, pm=(1.0,) * 4, out=None):
    p = ab.make_piece(w, h, filters=util.BAYER["RGGB"] if ch == 1 else 0, channels=ch, data=data, processed_maximum=pm, devid=0,
                      out_width=out[0] if out else None, out_height=out[1] if out else None, wb_coeffs=wb)
    p.datatype = datatype
    return p


def go(rc):
    ab.check(rc)
    torch.cuda.synchronize()


p_rp = piece(ab.rawprepare_data((512.0,) * 4, (15871.0,) * 4), 1, ab.TYPE_UINT16)
p_tp = piece(ab.temperature_data(wb), 1)
p_hl = piece(ab.highlights_data(ab.HIGHLIGHTS_CLIP, 1.0), 1, pm=(wb[0], wb[1], wb[2], 0.0))
p_hi = piece(ab.highlights_data(ab.HIGHLIGHTS_INPAINT, 1.0), 1, pm=(wb[0], wb[1], wb[2], 0.0))
go(L.b200_rawprepare_process_dev(p_rp, raw.data_ptr(), m0.data_ptr(), s))
go(L.b200_temperature_process_dev(p_tp, m0.data_ptr(), m1.data_ptr(), s))
go(L.b200_highlights_process_dev(p_hl, m1.data_ptr(), m0.data_ptr(), s))
go(L.b200_highlights_process_dev(p_hi, m1.data_ptr(), m0.data_ptr(), s))
go(L.b200_rawfront_process_dev(p_rp, p_tp, p_hl, raw.data_ptr(), m0.data_ptr(), s))
for method in (ab.DEMOSAIC_PPG, ab.DEMOSAIC_VNG4, ab.DEMOSAIC_RCD | 2048):
    d = ab.demosaic_data(method)
    d.median_thrs, d.dual_thrs = 0.02, 0.2
    go(L.b200_demosaic_process_dev(piece(d, 1), m0.data_ptr(), b.data_ptr(), s))
p_lc = piece(ab.highlights_data(ab.HIGHLIGHTS_LCH, 1.0), 1, pm=(wb[0], wb[1], wb[2], 0.0))
go(L.b200_highlights_process_dev(p_lc, m1.data_ptr(), m0.data_ptr(), s))
lab = a * torch.tensor([100.0, 60.0, 60.0, 1.0], device="cuda")
go(L.b200_bilat_process_dev(piece(ab.bilat_data(sigma_r=5.0, sigma_s=50.0, detail=0.5, mode=0), 4), lab.data_ptr(), b.data_ptr(), s))
go(L.b200_exposure_process_dev(piece(ab.exposure_data(0.0, 0.5), 4), a.data_ptr(), b.data_ptr(), s))
cp = ab.channelmixer_piece(util.profile_pair(util.REC2020_TO_XYZ_D50), illuminant=(0.93, 1.02, 0.71))
pc = piece(None, 4)
pc.data, pc.data_size = C.addressof(cp), C.sizeof(cp)
go(L.b200_channelmixerrgb_process_dev(pc, a.data_ptr(), b.data_ptr(), s))
go(L.b200_gamma_process_dev(piece(None, 4), a.data_ptr(), u8.data_ptr(), s))
go(L.b200_export_convert_dev(a.data_ptr(), b.data_ptr(), w, h, ab.EXPORT_UINT16, s))
p_fs = piece(ab.finalscale_data(), 4, out=(w // 2, h // 2))
p_fs.roi_out.scale = 0.5
go(L.b200_finalscale_process_dev(p_fs, a.data_ptr(), b.data_ptr(), s))
print("done")
