#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q -s --timeout=600 -k "host_batches or auto_graph or resident or script_entry or cuda_graph_step" > gpurun_out/r25_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r25_pytest.log
timeout 300 python tools/prof_host_path.py > gpurun_out/r25_host_path.txt 2>&1; echo "host path rc=$?"; grep -v Warning gpurun_out/r25_host_path.txt | tail -12
timeout 500 python bench.py --steps 20 --warmup 5 --no-sweep > gpurun_out/r25_bench_n1.json 2> gpurun_out/r25_bench_n1.err; echo "bench N=1 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r25_bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['mode'][:60], 'legs_failed', d.get('legs_failed'))
print('extra e2e', d['extra']['e2e_samples_s'], d['extra']['samples_s'])
PY
