set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlm_gpu.py tests/test_denoise_gpu.py tests/test_chains_gpu.py -m gpu -q > gpurun_out/pytest_nlm.log 2>&1; echo "nlm+denoise+chains tests rc=$?"; tail -8 gpurun_out/pytest_nlm.log
timeout 120 python tools/time_nlm.py
B200_NLM_NO_PIPE=1 timeout 120 python tools/time_nlm.py
timeout 120 python tools/time_nlm.py 2 7
NLM_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlm_pipe -s 1 -c 1 -o gpurun_out/r02_nlm_pipe python tools/time_nlm.py > gpurun_out/ncu_nlm.log 2>&1; ls -la gpurun_out/*.ncu-rep
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 600 gpurun_out/bench_c3.err; head -c 1500 gpurun_out/bench_c3.json
