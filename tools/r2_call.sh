#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head_nms.py tests/test_dropin.py tests/test_prepost.py -q -m gpu > gpurun_out/c41_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c41_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -10 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/c41_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c41_smoke.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/c41_bench.json 2> gpurun_out/c41_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/c41_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['e2e'].get('blocks_ms_per_step'), d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'])
print([ (p['N'],p['K'],round(p['ms'],3),p['bit_exact_vs_c_oracle_2_images']) for p in d['nms']['points']])
PY
timeout 900 tools/profile_hbm.sh r2t --skip-train
