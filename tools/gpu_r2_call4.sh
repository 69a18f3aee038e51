# round 2, fourth GPU call: bench (C3 headline), launch list, the new strip / dual / extreme-pixel tests, the parity table
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_c3.err; head -c 5000 gpurun_out/bench_c3.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-modules > gpurun_out/ncu_bench.log 2>&1; wc -l gpurun_out/r02_launches_bench_c3.csv
timeout 1200 python -m pytest tests/test_nlm_gpu.py tests/test_zz_vng_gpu.py tests/test_chains_gpu.py tests/test_denoise_gpu.py tests/test_diffuse_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; echo "new tests rc=$?"; tail -6 gpurun_out/pytest_new.log
timeout 1500 python tools/parity_table.py > gpurun_out/parity_vs_release_build.md 2> gpurun_out/parity.err; echo "parity rc=$?"; tail -3 gpurun_out/parity.err; cat gpurun_out/parity_vs_release_build.md
