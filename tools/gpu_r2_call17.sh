set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_lmmse_gpu.py -m gpu -q > gpurun_out/pytest_lmmse.log 2>&1; echo "lmmse tests rc=$?"; tail -6 gpurun_out/pytest_lmmse.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -6 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err; head -c 400 gpurun_out/bench_c3.json; echo
