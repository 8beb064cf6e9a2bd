#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_head_nms.py tests/test_dropin.py -q -m gpu > gpurun_out/c45_pytest.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/c45_pytest.log
timeout 60 python __graft_entry__.py smoke > gpurun_out/c45_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/c45_smoke.log
