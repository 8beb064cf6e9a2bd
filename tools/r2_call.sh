#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c20_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -10 | cut -c1-300
timeout 1500 tools/profile_r2.sh r2
timeout 900 tools/profile_hbm.sh r2g
