set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nlm_gpu.py tests/test_chains_gpu.py -m gpu -q > gpurun_out/pytest_nlm.log 2>&1; echo "nlm tests rc=$?"; tail -3 gpurun_out/pytest_nlm.log
timeout 120 python tools/time_nlm.py
timeout 120 python tools/time_nlm.py 2 7
timeout 600 python -m pytest tests/test_zz_markesteijn_gpu.py tests/test_zz_vng_gpu.py -m gpu -q > gpurun_out/pytest_mk.log 2>&1; echo "markesteijn tests rc=$?"; tail -15 gpurun_out/pytest_mk.log
timeout 120 python tools/time_markesteijn.py
NLM_REPS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:nlm_pipe -s 1 -c 1 -o /tmp/r02_nlm_pipe python tools/time_nlm.py > gpurun_out/ncu_nlm.log 2>&1
python tools/ncu_summary.py /tmp/r02_nlm_pipe.ncu-rep > gpurun_out/r02_nlm_pipe_ncu.md
python tools/ncu_lines.py /tmp/r02_nlm_pipe.ncu-rep "" 40 > gpurun_out/r02_nlm_pipe_lines.txt
cat gpurun_out/r02_nlm_pipe_ncu.md; head -16 gpurun_out/r02_nlm_pipe_lines.txt
