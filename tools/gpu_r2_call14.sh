set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_blend_gpu.py tests/test_nlm_gpu.py tests/test_chains_gpu.py tests/test_adapters_gpu.py -m gpu -q > gpurun_out/pytest_blend.log 2>&1; echo "blend+nlm+chains+adapters tests rc=$?"; tail -8 gpurun_out/pytest_blend.log
timeout 120 python tools/time_nlm.py
NLM_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlm_pipe -s 1 -c 1 -o /tmp/r02_nlm_pipe python tools/time_nlm.py > gpurun_out/ncu_nlm.log 2>&1
python tools/ncu_summary.py /tmp/r02_nlm_pipe.ncu-rep > gpurun_out/r02_nlm_pipe_ncu.md
python tools/ncu_lines.py /tmp/r02_nlm_pipe.ncu-rep "" 40 > gpurun_out/r02_nlm_pipe_lines.txt
cat gpurun_out/r02_nlm_pipe_ncu.md
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 400 gpurun_out/bench_c3.err; head -c 700 gpurun_out/bench_c3.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-modules --e2e-steps 2 > gpurun_out/bench_under_ncu.log 2>&1; tail -3 gpurun_out/r02_launches_bench_c3.csv | cut -c1-200
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; head -c 600 gpurun_out/bench_ref.json
