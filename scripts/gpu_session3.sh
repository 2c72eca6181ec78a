#!/bin/bash
mkdir -p gpurun_out
echo "== gemm debug+timing"; timeout -k 10 300 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "exit $?"; grep -E "rel_err|TF|rc [1-9]|rows" gpurun_out/debug_gemm.log | head -60
echo "== pytest gpu"; timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== bench"; timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench B=8"; timeout -k 10 600 python bench.py --steps 5 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_b8.json; tail -3 gpurun_out/bench_b8.err
echo "== ncu full"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"sample_kl|bayes_gemm|bayes_wgrad" -c 14 -f -o gpurun_out/prof_r1 python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_full.log
echo "== microbench"; timeout -k 10 300 python scripts/gpu_microbench.py > gpurun_out/microbench.log 2>&1; echo "exit $?"; grep -E "bwd|S': (1|4)," gpurun_out/microbench.log | head -30
