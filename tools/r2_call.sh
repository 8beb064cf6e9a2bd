#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_head_nms.py tests/test_gpu_tiny.py tests/test_dropin.py tests/test_abi_and_host.py tests/test_prepost.py -q -m gpu > gpurun_out/c25_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c25_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -10 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/c25_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/c25_smoke.log
timeout 900 tools/profile_hbm.sh r2j --skip-train
