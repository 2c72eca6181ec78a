#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout -k 5 700 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu.log | head -30 | cut -c1-300
for PF in 0 1 2; do MULTI_ONLY=1 BF_SK_PREFETCH=$PF timeout 200 python scripts/gpu_multi_microbench.py 2>&1 | tail -1; done
timeout 200 python scripts/gpu_multi_microbench.py 2>&1 | tail -4
