#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_tiny.py tests/test_gpu_options.py -q -m gpu -s > gpurun_out/c11_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c11_pytest.log | grep "forward\|gradients vs\|passed\|failed\|Error\|error\|assert \|FAILED\|lowest cos" | tail -40 | cut -c1-330
timeout 2400 tools/sanitize.sh r2
