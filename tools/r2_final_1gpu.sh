#!/bin/bash
# final 1-GPU evidence run of round 2: full GPU suite, smoke, ncu captures, launch list, bench + reference arm
mkdir -p gpurun_out; T=${1:-r31}
timeout 900 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:multi_tensor_copy -s 10 -c 6 -f -o gpurun_out/${T}_prof_copy python tools/prof_kernels.py > gpurun_out/${T}_ncu_copy.log 2>&1; echo "ncu copy rc=$?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:multi_tensor_copy -c 20 -f -o gpurun_out/${T}_profile_script python tools/prof_profile_script.py > gpurun_out/${T}_ncu_profile_script.log 2>&1; echo "ncu profile-script rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 20 --warmup 5 --blocks 2 --no-sweep --no-extra --no-nccl --no-cpu-baseline > gpurun_out/${T}_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke_ncu.log 2>&1; echo "smoke-ncu rc=$?"; tail -1 gpurun_out/${T}_smoke_ncu.log
timeout 300 python tools/prof_host_path.py > gpurun_out/${T}_host_path.txt 2>&1; echo "host path rc=$?"
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/${T}_ref_n1.json 2> gpurun_out/${T}_ref_n1.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/${T}_bench_n1.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/${T}_ref_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'ref', r['value'], 'same_config', d['config']==r['config'], 'legs_failed', d.get('legs_failed'))
print('roofline', d['roofline']['kernel'], d['roofline']['frac'])
PY
