# round 2, first GPU call: the non-local-means group kernel -- parity, timing against the chunk kernel, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlm_gpu.py tests/test_denoise_gpu.py -m gpu -q > gpurun_out/pytest_nlm.log 2>&1; echo "nlm rc=$?"; tail -15 gpurun_out/pytest_nlm.log
for g in 6 4 2; do B200_NLM_G=$g timeout 120 python tools/time_nlm.py; done
B200_NLM_IEEE_DIV=1 timeout 120 python tools/time_nlm.py
B200_NLM_CHUNKS=1 NLM_REPS=2 timeout 120 python tools/time_nlm.py
timeout 120 python tools/time_nlm.py 2 7
timeout 120 python tools/time_nlm.py 1 4
NLM_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlm_group -s 1 -c 1 -o gpurun_out/r02_nlm_group python tools/time_nlm.py > gpurun_out/ncu_nlm.log 2>&1; ls -la gpurun_out/*.ncu-rep
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
