#!/bin/bash
# ncu --set full of the fused dropout + residual + LayerNorm kernels (one fwd, one bwd launch each variant)
mkdir -p gpurun_out
echo "== ncu full resln (staged)"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"resln" -s 6 -c 2 -f -o gpurun_out/prof_resln python scripts/gpu_resln_microbench.py > gpurun_out/ncu_resln.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_resln.log
echo "== ncu full resln (regs)"; BF_RESLN_BWD=regs timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"resln_bwd" -s 3 -c 1 -f -o gpurun_out/prof_resln_regs python scripts/gpu_resln_microbench.py > gpurun_out/ncu_resln_regs.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_resln_regs.log
ls -la gpurun_out/*.ncu-rep
