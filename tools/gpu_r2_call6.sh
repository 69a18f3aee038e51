set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlm_gpu.py tests/test_denoise_gpu.py -m gpu -q > gpurun_out/pytest_nlm.log 2>&1; echo "nlm+denoise tests rc=$?"; tail -8 gpurun_out/pytest_nlm.log
timeout 120 python tools/time_nlm.py
B200_NLM_NO_PIPE=1 timeout 120 python tools/time_nlm.py
bash tools/gpu_r2_profile.sh
