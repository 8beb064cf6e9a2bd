#!/bin/bash
# One validation visit to a B200 (gpurun -- 'bash tools/r2_call.sh'): the -m gpu suite, smoke(), the default bench line with the
# per-layer report.  During development this file was the scratch driver of each visit (edited per call).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/validate_pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^$" gpurun_out/validate_pytest.log | grep "passed\|failed\|Error\|error\|assert \|FAILED" | tail -10 | cut -c1-300
timeout 300 python __graft_entry__.py smoke > gpurun_out/validate_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/validate_smoke.log
timeout 600 python bench.py --layer-report gpurun_out/validate_layers.json > gpurun_out/validate_bench.json 2> gpurun_out/validate_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/validate_bench.json
