set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "all gpu tests rc=$?"; tail -5 gpurun_out/pytest_all.log
bash tools/gpu_r2_profile.sh
