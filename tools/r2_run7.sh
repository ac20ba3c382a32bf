#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}; TAG=${2:-r26}
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench N=$N rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_n${N}.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], d['block_ms'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'legs_failed', d.get('legs_failed'))
print('eager', d['eager']['value'], 'nccl', d['nccl_ddp']['value'], 'parity ok', d['parity']['ok'], 'extra', d['extra']['samples_s'], d['extra']['e2e_samples_s'], d['extra']['nccl_ddp'].get('value'))
print('roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], 'step', d['roofline_step']['avg_launch_ms'])
PY
