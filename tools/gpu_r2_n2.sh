set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_bands_gpu.py -m gpu -q > gpurun_out/pytest_bands_n2.log 2>&1; echo "bands tests rc=$?"; tail -5 gpurun_out/pytest_bands_n2.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --c4 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/bench_n2.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}); print('e2e', d.get('e2e'))
    for k,v in d['config'].items():
        if k in ('banded_one_frame','batch_export','c5','batch'): print(k, json.dumps(v)[:3000])
    print([k for k in d['config']])
except Exception as e: print('parse failed', e)
PY
