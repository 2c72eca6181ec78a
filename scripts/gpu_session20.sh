#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout -k 5 500 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== microbench"; timeout -k 10 300 python scripts/gpu_microbench.py > gpurun_out/microbench.log 2>&1; echo "exit $?"; grep -E "bwd|S': (1|4)," gpurun_out/microbench.log | head -30
