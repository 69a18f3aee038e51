#!/bin/bash
# round 2, call 21: the restructured colour inpainting and the column pass of the guided laplacians
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_pipe_ends_gpu.py tests/test_zz_hl_laplacian_gpu.py tests/test_adapters_gpu.py -m gpu -q > gpurun_out/pytest_touched.log 2>&1; echo "touched tests rc=$?"
tail -5 gpurun_out/pytest_touched.log
timeout 300 python tools/time_inpaint.py 2>&1 | tail -1
timeout 300 python tools/time_hl_laplacian.py 2 2>&1 | tail -1
timeout 300 python tools/time_hl_laplacian.py 2>&1 | tail -1
