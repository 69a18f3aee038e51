set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nlm_gpu.py tests/test_denoise_gpu.py tests/test_chains_gpu.py "tests/test_zz_vng_gpu.py::test_dual_rcd_vng4_bit_exact" -m gpu -q > gpurun_out/pytest_nlm.log 2>&1; echo "cfg0 tests rc=$?"; tail -4 gpurun_out/pytest_nlm.log
B200_NLM_PIPE_CFG=1 timeout 900 python -m pytest tests/test_nlm_gpu.py -m gpu -q > gpurun_out/pytest_nlm_cfg1.log 2>&1; echo "cfg1 tests rc=$?"; tail -4 gpurun_out/pytest_nlm_cfg1.log
timeout 120 python tools/time_nlm.py
B200_NLM_PIPE_CFG=1 timeout 120 python tools/time_nlm.py
timeout 120 python tools/time_nlm.py 2 7
B200_NLM_PIPE_CFG=1 timeout 120 python tools/time_nlm.py 2 7
bash tools/gpu_r2_profile.sh
