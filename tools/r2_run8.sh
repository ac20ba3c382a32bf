#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r28_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r28_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r28_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r28_smoke.log
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/r28_bench_n1.json 2> gpurun_out/r28_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r28_ref_n1.json 2> gpurun_out/r28_ref_n1.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r28_bench_n1.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r28_ref_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ref', r['value'], 'same_config', d['config']==r['config'], 'legs_failed', d.get('legs_failed'), 'launches', d['gpu_launches'])
print('roofline', d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic'])
PY
