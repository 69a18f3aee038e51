set -x
mkdir -p gpurun_out
timeout 300 python tools/dbg_nlm.py 2>&1 | tail -40
timeout 900 python -m pytest tests/test_nlm_gpu.py tests/test_chains_gpu.py -m gpu -q > gpurun_out/pytest_pipe.log 2>&1; echo "pipe tests rc=$?"; tail -8 gpurun_out/pytest_pipe.log
timeout 120 python tools/time_nlm.py
B200_NLM_NO_PIPE=1 timeout 120 python tools/time_nlm.py
timeout 120 python tools/time_nlm.py 2 7
NLM_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlm_pipe -s 1 -c 1 -o gpurun_out/r02_nlm_pipe python tools/time_nlm.py > gpurun_out/ncu_nlm.log 2>&1; ls -la gpurun_out/*.ncu-rep
