#!/bin/bash
# round 2, call 19: highlights guided laplacians (Bayer, X-Trans, RGBA) on the GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_hl_laplacian_gpu.py -m gpu -q > gpurun_out/pytest_hl.log 2>&1; echo "hl laplacian tests rc=$?"
tail -12 gpurun_out/pytest_hl.log
timeout 600 python -m pytest tests/test_zz_pipe_ends_gpu.py tests/test_adapters_gpu.py -m gpu -q > gpurun_out/pytest_touched.log 2>&1; echo "touched tests rc=$?"
tail -5 gpurun_out/pytest_touched.log
