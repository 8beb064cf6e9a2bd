#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 900 python tools/probe_fmt.py > gpurun_out/c4_probe_fmt.log 2>&1; echo "probe_fmt rc=$?"; cat gpurun_out/c4_probe_fmt.log | cut -c1-1500
timeout 900 python -m pytest tests/test_gpu_head_nms.py tests/test_prepost.py tests/test_gpu_optimizer.py -x -q -m gpu > gpurun_out/c4_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/c4_pytest.log
timeout 900 tools/profile_hbm.sh r2b --skip-train
timeout 600 ncu --profile-from-start off --set full --clock-control none --import-source on -k 'regex:decode_kernel|nms_select|nms_apply|detections' -c 5 -f -o gpurun_out/prof_head_r2b python tools/hbm_kernels.py --skip-train > gpurun_out/ncu_head_r2b.log 2>&1; echo "ncu head rc=$?"
timeout 600 python bench.py --steps 100 --no-cpu-baseline > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c4_bench.json')); print(d['value'], d['e2e']['value'], d['roofline']['share_of_step'], [(p['N'],p['K'],round(p['ms'],3),p['bit_exact_vs_c_oracle_2_images']) for p in d['nms']['points']])"
