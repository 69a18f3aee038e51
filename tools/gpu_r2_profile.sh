# round 2: one ncu --set full capture per kernel that had none in profiles/ (amaze, filmic, vst, ppg, vng, inpaint, bilateral, resample ...)
# the reports are summarised on the box (tools/ncu_summary.py) and removed: gpurun_out/ carries 64 MiB at most
set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none -k regex:'amaze_tiles|filmic_agx|vst_forward|vst_backward|eaw_decompose|eaw_synthesize|heat_pde|bspline|ll_|rcd_tiles|convert_kernel' -c 40 -o /tmp/r02_modules python tools/profile_all.py > gpurun_out/ncu_modules.log 2>&1; tail -2 gpurun_out/ncu_modules.log
python tools/ncu_summary.py /tmp/r02_modules.ncu-rep > gpurun_out/r02_modules_ncu.md
timeout 1200 ncu --set full --clock-control none -k regex:'ppg_kernel|pre_median|vng_kernel|lin_interpolate|inpaint_|bilateral|splat|slice|resample_kernel|channelmixer_kernel|raw_front|gamma_kernel|lch_bayer|detail_|dual_' -c 40 -o /tmp/r02_pipe_ends python tools/profile_pipe_ends.py > gpurun_out/ncu_pipe_ends.log 2>&1; tail -2 gpurun_out/ncu_pipe_ends.log
python tools/ncu_summary.py /tmp/r02_pipe_ends.ncu-rep > gpurun_out/r02_pipe_ends_ncu.md
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
