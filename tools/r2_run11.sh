#!/bin/bash
mkdir -p gpurun_out
for mode in plain fused; do
  flag=""; [ $mode = fused ] && flag="--fuse-sgd"
  timeout 300 python bench.py --steps 20 --warmup 5 --no-sweep --no-extra --no-cpu-baseline --no-nccl $flag > gpurun_out/r34_bench_n1_$mode.json 2> gpurun_out/r34_bench_n1_$mode.err; echo "$mode rc=$?"
done
python - <<'PY'
import json
for m in ('plain','fused'):
    d=json.loads(open(f'gpurun_out/r34_bench_n1_{m}.json').read().strip().splitlines()[-1])
    print(m, 'value', round(d['value']), round(d['ms_per_step']*1e3,2), 'us; e2e', round(d['e2e']['value']), 'eager', round(d['eager']['value']), d['impl_config'].get('fused_sgd'), d['gpu_launches'], d.get('legs_failed'))
PY
