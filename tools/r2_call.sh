#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 150 python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_final2_train.json 2> gpurun_out/c43_train.err; echo "train rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final2_train.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'])
PY
