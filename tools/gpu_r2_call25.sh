#!/bin/bash
# round 2, call 25 (the last GPU seconds of the round): blending in the display-referred space and the committed reference outputs on the GPU
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_zz_blend_gpu.py -m gpu -q -k "display or committed" > gpurun_out/pytest_blend_display.log 2>&1; echo "display blend tests rc=$?"
tail -4 gpurun_out/pytest_blend_display.log
