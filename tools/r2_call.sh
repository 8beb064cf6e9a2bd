#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c12_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -30 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/c12_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c12_smoke.log
timeout 600 python bench.py --layer-report gpurun_out/c12_layers.json > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/c12_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c12_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['f32_variant']['value'], d['roofline']['frac'], d['roofline']['share_of_step'], d['cpu_baseline']['value'], d['clocks'])"
timeout 900 tools/profile_hbm.sh r2c
