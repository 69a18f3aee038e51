# round 2, second GPU call: group kernel v2 (compile-time pitches, packed B1, interior paths)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlm_gpu.py tests/test_denoise_gpu.py tests/test_chains_gpu.py -m gpu -q > gpurun_out/pytest_nlm.log 2>&1; echo "nlm rc=$?"; tail -8 gpurun_out/pytest_nlm.log
timeout 120 python tools/time_nlm.py
B200_NLM_G=4 timeout 120 python tools/time_nlm.py
B200_NLM_IEEE_DIV=1 timeout 120 python tools/time_nlm.py
timeout 120 python tools/time_nlm.py 2 7
NLM_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlm_group -s 1 -c 1 -o gpurun_out/r02_nlm_group_v2 python tools/time_nlm.py > gpurun_out/ncu_nlm.log 2>&1; ls -la gpurun_out/*.ncu-rep
