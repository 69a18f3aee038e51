# round 2, third GPU call: the new bench.py (C3 headline) at N=1, both arms; band tests
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_c3.err; head -c 6000 gpurun_out/bench_c3.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -c 800 gpurun_out/bench_ref.err; head -c 2500 gpurun_out/bench_ref.json
timeout 600 python -m pytest tests/test_bands_gpu.py tests/test_adapters_gpu.py -m gpu -q 2>&1 | tail -5
