set -x
mkdir -p gpurun_out
timeout 120 python tools/dbg_markesteijn.py small one_tile roi
timeout 300 compute-sanitizer --tool memcheck python tools/dbg_markesteijn.py small 2>&1 | tail -30
timeout 300 compute-sanitizer --tool racecheck python tools/dbg_markesteijn.py small 2>&1 | tail -30
