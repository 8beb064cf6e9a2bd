#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head_nms.py tests/test_prepost.py -q -m gpu > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c19_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -20 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/c19_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c19_smoke.log
timeout 600 python tools/ab_nms_apply.py > gpurun_out/c19_ab_nms.log 2>&1; echo "ab rc=$?"; cat gpurun_out/c19_ab_nms.log | tail -6 | cut -c1-300
timeout 300 python bench.py --size 608 --steps 60 --no-cpu-baseline --no-nms-sweep > gpurun_out/c19_bench_608.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/c19_bench_608.json')); print('608:', d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['share_of_step'], d['nms_load'])"
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c19_bench.json 2> gpurun_out/c19_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c19_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c19_bench.json')); print(d['value'], d['e2e']['value'], d['gpu_launches'], d['roofline']['share_of_step'], [(p['N'],p['K'],round(p['ms'],3),p['bit_exact_vs_c_oracle_2_images']) for p in d['nms']['points']])"
