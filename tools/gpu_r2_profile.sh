# round 2: one ncu --set full capture per kernel that had none in profiles/ (amaze, filmic, vst, ppg, vng, inpaint, bilateral, resample ...)
set -x
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'amaze_tiles|filmic_agx|vst_forward|vst_backward|eaw_decompose|eaw_synthesize|heat_pde|bspline|ll_|rcd_tiles|convert_kernel' -c 40 -o gpurun_out/r02_modules python tools/profile_all.py > gpurun_out/ncu_modules.log 2>&1; tail -2 gpurun_out/ncu_modules.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'ppg_kernel|pre_median|vng_kernel|lin_interpolate|inpaint_|bilateral|splat|slice|resample_kernel|channelmixer_kernel|raw_front|gamma_kernel|lch_bayer|detail_|dual_' -c 40 -o gpurun_out/r02_pipe_ends python tools/profile_pipe_ends.py > gpurun_out/ncu_pipe_ends.log 2>&1; tail -2 gpurun_out/ncu_pipe_ends.log
ls -la gpurun_out/*.ncu-rep
