#!/bin/bash
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout=400 -k "multigpu and (fused_sgd or regrow)" > gpurun_out/r37_pytest_multi_n8.log 2>&1; echo "pytest rc=$?"; grep -E "worst rel|passed|failed" gpurun_out/r37_pytest_multi_n8.log | tail -6
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29751 bench.py --gpus 8 --steps 20 --warmup 5 --fuse-sgd --no-sweep --no-extra --no-nccl > gpurun_out/r37_bench_n8_fused.json 2> gpurun_out/r37_bench_n8_fused.err; echo "bench fused N=8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r37_bench_n8_fused.json').read().strip().splitlines()[-1])
print('value', round(d['value']), round(d['ms_per_step']*1e3,2), 'us; e2e', round(d['e2e']['value']), 'eager', round(d['eager']['value']), d['impl_config'].get('fused_sgd'), 'parity', d['parity']['ok'], d.get('legs_failed'))
PY
