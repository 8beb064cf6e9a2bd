#!/bin/bash
# scratch driver for one gpurun visit (edited per call)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_tiny.py tests/test_gpu_optimizer.py -x -q -m gpu -s > gpurun_out/c8_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/c8_pytest.log | grep "forward\|gradients vs\|passed\|failed\|Error\|error\|assert" | tail -40 | cut -c1-400
timeout 600 python tools/diag_train.py 64 416 20 1 > gpurun_out/c8_diag_b64.json 2> gpurun_out/c8_diag_b64.err; echo "diag b64 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/c8_diag_b64.json')); print('net', d['net'], 'dnet', d['dnet'], 'grads worst', d['worst_grad'], d['worst_grad_fp32_floor'], 'l2', d['worst_grad_l2'], d['worst_grad_l2_fp32_floor'])"
