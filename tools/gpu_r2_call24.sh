#!/bin/bash
# round 2, call 24: the whole GPU suite with the Lab / JzCzhz blending and the new libm functions, smoke, the default bench
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "all gpu tests rc=$?"
tail -8 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; tail -c 300 gpurun_out/bench_c3.err; head -c 300 gpurun_out/bench_c3.json; echo
