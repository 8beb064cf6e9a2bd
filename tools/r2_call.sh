#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c23_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c23_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -10 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/c23_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c23_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_reference.json 2> gpurun_out/final_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/final_bench_reference.json
timeout 600 python bench.py --layer-report gpurun_out/final_layers.json > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench.err; python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); print(d['value'], d['e2e']['value'], d['e2e']['f32_variant']['value'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['share_of_step'], d['cpu_baseline']['value'], d['clocks'], [(p['N'],p['K'],round(p['ms'],3),p['bit_exact_vs_c_oracle_2_images']) for p in d['nms']['points']])"
timeout 600 python bench.py --workload train > gpurun_out/final_bench_train.json 2> gpurun_out/final_train.err; echo "train rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/final_bench_train.json')); print(d['value'], d['e2e']['value'], d['train']['phases_ms'], d['cpu_baseline']['value'], d['clocks'])"
timeout 600 python bench.py --size 608 --steps 100 --no-cpu-baseline --no-nms-sweep > gpurun_out/final_bench_608.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/final_bench_608.json')); print('608:', d['value'], d['e2e']['value'], d['roofline']['frac'], d['nms_load'])"
timeout 900 tools/profile_hbm.sh r2i
