#!/bin/bash
# round 2, call 23: blending in the Lab space on the GPU
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zz_blend_gpu.py -m gpu -q > gpurun_out/pytest_blend.log 2>&1; echo "blend tests rc=$?"
tail -8 gpurun_out/pytest_blend.log
