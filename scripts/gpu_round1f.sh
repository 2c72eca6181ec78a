#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu -k resln/output/layernorm"; timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 150 -k "resln or output_blocks or layernorm" > gpurun_out/pytest_sel.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_sel.log | head -30 | cut -c1-300
echo "== resln microbench (staged)"; FUSED_ONLY=1 timeout 200 python scripts/gpu_resln_microbench.py 2>&1 | tail -1
echo "== resln microbench (regs)"; FUSED_ONLY=1 BF_RESLN_BWD=regs timeout 200 python scripts/gpu_resln_microbench.py 2>&1 | tail -1
echo "== resln microbench H=1024 (staged / regs)"; FUSED_ONLY=1 H=1024 timeout 200 python scripts/gpu_resln_microbench.py 2>&1 | tail -1;  FUSED_ONLY=1 H=1024 BF_RESLN_BWD=regs timeout 200 python scripts/gpu_resln_microbench.py 2>&1 | tail -1
echo "== ncu full resln bwd (staged)"; FUSED_ONLY=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"resln_bwd" -s 3 -c 1 -f -o gpurun_out/prof_resln_staged python scripts/gpu_resln_microbench.py > gpurun_out/ncu_resln.log 2>&1; echo "exit $?"; tail -1 gpurun_out/ncu_resln.log
