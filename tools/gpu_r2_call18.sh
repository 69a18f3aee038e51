#!/bin/bash
# round 2, call 18: highlights guided laplacians on the GPU, the tests touched since call 17, timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_hl_laplacian_gpu.py -m gpu -q -x > gpurun_out/pytest_hl.log 2>&1; echo "hl laplacian tests rc=$?"
tail -8 gpurun_out/pytest_hl.log
timeout 900 python -m pytest tests/test_nlm_gpu.py tests/test_rcd_gpu.py tests/test_zz_lmmse_gpu.py tests/test_diffuse_gpu.py tests/test_zz_pipe_ends_gpu.py tests/test_color_gpu.py -m gpu -q > gpurun_out/pytest_touched.log 2>&1; echo "touched tests rc=$?"
tail -6 gpurun_out/pytest_touched.log
timeout 300 python tools/time_hl_laplacian.py 2>&1 | tail -1
timeout 300 python tools/time_hl_laplacian.py 2 2>&1 | tail -1
