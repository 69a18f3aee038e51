set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_final2.json 2> gpurun_out/bench_ref_final2.err; tail -c 600 gpurun_out/bench_ref_final2.json
timeout 400 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 1500 gpurun_out/bench_final2.json
timeout 200 python tools/time_amaze.py 2>&1 | tail -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_bench_c2.csv python bench.py --steps 2 --warmup 1 --no-other-modules > gpurun_out/ncu_bench2.log 2>&1; wc -l gpurun_out/launches_bench_c2.csv
