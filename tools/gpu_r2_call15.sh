set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_markesteijn_gpu.py tests/test_zz_vng_gpu.py -m gpu -q > gpurun_out/pytest_mk.log 2>&1; echo "markesteijn+vng tests rc=$?"; tail -6 gpurun_out/pytest_mk.log
timeout 120 python tools/time_markesteijn.py
timeout 120 python tools/time_markesteijn.py 3
