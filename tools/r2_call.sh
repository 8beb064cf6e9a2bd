#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head_nms.py tests/test_prepost.py -q -m gpu > gpurun_out/c14_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c14_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -20 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/c14_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c14_smoke.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c14_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c14_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['share_of_step'], [(p['N'],p['K'],round(p['ms'],3),p['bit_exact_vs_c_oracle_2_images']) for p in d['nms']['points']])"
timeout 900 tools/profile_hbm.sh r2e --skip-train
