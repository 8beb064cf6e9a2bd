#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/c40_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c40_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -10 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/c40_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c40_smoke.log
timeout 600 python bench.py --layer-report gpurun_out/layers_r2_final2.json > gpurun_out/bench_r2_final2.json 2> gpurun_out/c40_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final2.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e'].get('blocks_ms_per_step'), d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['cpu_baseline']['value'])
PY
