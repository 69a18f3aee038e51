set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nlm_gpu.py tests/test_chains_gpu.py -m gpu -q -x > gpurun_out/pytest_nlm.log 2>&1; echo "nlm (half slots) tests rc=$?"; tail -3 gpurun_out/pytest_nlm.log
timeout 100 python tools/time_nlm.py
B200_NLM_WHOLE_SLOTS=1 timeout 100 python tools/time_nlm.py
timeout 600 python -m pytest tests/test_zz_markesteijn_gpu.py tests/test_zz_vng_gpu.py -m gpu -q > gpurun_out/pytest_mk.log 2>&1; echo "markesteijn+vng tests rc=$?"; tail -4 gpurun_out/pytest_mk.log
NLM_REPS=1 timeout 400 ncu --set full --clock-control none --import-source on -k regex:nlm_pipe -s 1 -c 1 -o /tmp/r02_nlm_pipe python tools/time_nlm.py > gpurun_out/ncu_nlm.log 2>&1
python tools/ncu_summary.py /tmp/r02_nlm_pipe.ncu-rep > gpurun_out/r02_nlm_pipe_ncu.md
python tools/ncu_lines.py /tmp/r02_nlm_pipe.ncu-rep "" 40 > gpurun_out/r02_nlm_pipe_lines.txt
cat gpurun_out/r02_nlm_pipe_ncu.md; head -12 gpurun_out/r02_nlm_pipe_lines.txt
