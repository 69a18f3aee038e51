set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_markesteijn_gpu.py -m gpu -q > gpurun_out/pytest_mk.log 2>&1; echo "markesteijn tests rc=$?"; tail -6 gpurun_out/pytest_mk.log
timeout 120 python tools/time_markesteijn.py
for cfg in 0 1 2 3; do B200_NLM_PIPE_CFG=$cfg timeout 120 python tools/time_nlm.py; done
for cfg in 0 1 2 3; do B200_NLM_PIPE_CFG=$cfg timeout 300 python -m pytest tests/test_nlm_gpu.py -m gpu -q -x 2>&1 | tail -1; done
B200_NLM_PIPE_CFG=0 timeout 120 python tools/time_nlm.py
B200_NLM_PIPE_CFG=2 timeout 120 python tools/time_nlm.py
