#!/bin/bash
# First gpurun call of the next round: everything the round-1 tail could not measure, in one box visit.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- tools/round2_first.sh
# 1. the verified suite + smoke (the tree must still be green on a fresh box)
# 2. the experimental mixed-kind conv: gated GPU test, then the probe (error / time per term vs the shipped kernel)
# 3. the default bench line (for the same-box reference of any later A/B)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest -m gpu rc=$?"; tail -3 gpurun_out/r2_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
Y2_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_unverified.py tests/test_gpu_head_nms.py tests/test_gpu_backbone.py tests/test_prepost.py -q -m gpu -k "unverified or reference_source_golden or reference_graph_golden or reference_detect_draws or non_square or resize" > gpurun_out/r2_pytest_unverified.log 2>&1; echo "unverified-path tests rc=$?"; tail -15 gpurun_out/r2_pytest_unverified.log
Y2_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_experimental_mix.py -q -m gpu > gpurun_out/r2_pytest_mix.log 2>&1; echo "experimental mix test rc=$?"; tail -15 gpurun_out/r2_pytest_mix.log
timeout 300 python tools/probe_mix.py > gpurun_out/r2_probe_mix.log 2>&1; echo "probe_mix rc=$?"; tail -8 gpurun_out/r2_probe_mix.log
timeout 600 python tools/probe_mix_network.py > gpurun_out/r2_probe_mix_network.log 2>&1; echo "probe_mix_network rc=$?"; tail -14 gpurun_out/r2_probe_mix_network.log
timeout 600 python bench.py --layer-report gpurun_out/r2_layers.json > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; cat gpurun_out/r2_bench.json
