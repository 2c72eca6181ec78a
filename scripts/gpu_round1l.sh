#!/bin/bash
# final ncu evidence: (1) --set full of one launch of every hand-written kernel, summarised on the box (the .ncu-rep of
# ~30 launches exceeds what gpurun copies back), (2) launch list of one measured eager step of the default bench
mkdir -p gpurun_out
echo "== ncu full (final kernel set, 1 launch each)"; REPS=1 timeout -k 10 600 ncu --set full --clock-control none -k regex:"sample_kl|bayes_gemm|bayes_wgrad|wgrad_reduce|layernorm|bias_grad|resln|clip_adamw|grad_sumsq" -c 40 -f -o /tmp/prof_r1e python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -1 gpurun_out/ncu_full.log
python scripts/ncu_summary.py /tmp/prof_r1e.ncu-rep > gpurun_out/r01e_ncu_full_kernels.md; wc -l gpurun_out/r01e_ncu_full_kernels.md
bash scripts/gpu_launchlist.sh launches_r1e
ls -la gpurun_out
