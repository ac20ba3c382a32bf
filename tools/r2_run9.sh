#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29741 bench.py --gpus 2 --workload wide --no-sweep --no-extra > gpurun_out/r32_bench_wide_n2.json 2> gpurun_out/r32_bench_wide_n2.err; echo "wide N=2 rc=$?"
timeout 300 $TR --master-port 29743 bench.py --gpus 2 --graph off --no-sweep --no-extra --steps 200 --warmup 20 > gpurun_out/r32_bench_toy_eager_n2.json 2> gpurun_out/r32_bench_toy_eager_n2.err; echo "toy graph-off N=2 rc=$?"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --workload slices --no-sweep --no-extra --no-cpu-baseline > gpurun_out/r32_bench_slices_n1.json 2> gpurun_out/r32_bench_slices_n1.err; echo "slices N=1 rc=$?"
python - <<'PY'
import json
for f in ('r32_bench_wide_n2','r32_bench_toy_eager_n2','r32_bench_slices_n1'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],4), d['value_mode'][:40], 'e2e', round(d['e2e']['value']), 'nccl', d['nccl_ddp'] and round(d['nccl_ddp']['value']), 'step-roof', d['roofline_step'] and round(d['roofline_step']['frac'],3), 'failed', d.get('legs_failed'))
    except Exception as e:
        print(f, 'ERR', e)
PY
