#!/bin/bash
# round 2, call 20: ncu summaries of the kernels added after the first table, the whole GPU suite, smoke, the default bench
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'markesteijn_tiles|lmmse_tiles|blend_kernel|hl_|bspline' -c 80 -o /tmp/r02_new python tools/profile_new_modules.py > gpurun_out/ncu_new.log 2>&1; tail -2 gpurun_out/ncu_new.log
python tools/ncu_summary.py /tmp/r02_new.ncu-rep > gpurun_out/r02_new_modules_ncu.md; head -5 gpurun_out/r02_new_modules_ncu.md
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "all gpu tests rc=$?"
tail -6 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err; head -c 400 gpurun_out/bench_c3.json; echo
