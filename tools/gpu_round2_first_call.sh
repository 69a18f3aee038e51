# First gpurun call of the next round: everything written after round 1's GPU budget was spent runs on a B200 for the
# first time here.  `gpurun --timeout 1500 -- 'bash tools/gpu_round2_first_call.sh'`
set -x
mkdir -p gpurun_out
# the new tests on their own first (no -x: every failure is wanted), then the whole suite the way the driver runs it
timeout 900 python -m pytest tests/test_zz_bilateral_gpu.py tests/test_zz_filmic_reconstruct_gpu.py tests/test_zz_pipe_ends_gpu.py tests/test_zz_ppg_gpu.py tests/test_zz_vng_gpu.py -m gpu -q \
  > gpurun_out/pytest_zz.log 2>&1; echo "zz rc=$?"; tail -25 gpurun_out/pytest_zz.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_pipe_ends.csv \
  python tools/profile_pipe_ends.py > gpurun_out/ncu_pipe_ends.log 2>&1; wc -l gpurun_out/launches_pipe_ends.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'raw_front|resample_kernel|channelmixer_kernel|ppg_kernel|vng_kernel|gamma_kernel|bilateral_splat|lch_bayer' \
  -c 12 -o gpurun_out/pipe_ends_full python tools/profile_pipe_ends.py > gpurun_out/ncu_full.log 2>&1; ls -la gpurun_out/*.ncu-rep
