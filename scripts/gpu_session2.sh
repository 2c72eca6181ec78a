#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -k 10 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "== microbench"; timeout -k 10 300 python scripts/gpu_microbench.py > gpurun_out/microbench.log 2>&1; echo "exit $?"; grep -E "S': (1|4|16)," gpurun_out/microbench.log | head -30
echo "== bench"; timeout -k 10 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== bench B=8"; timeout -k 10 600 python bench.py --steps 5 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "exit $?"; cut -c1-400 gpurun_out/bench_b8.json
echo "== ncu full"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"sample_kl|bayes_gemm" -s 4 -c 8 -f -o gpurun_out/prof_r1 python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -3 gpurun_out/ncu_full.log
echo "== ncu launches"; timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 1 --warmup 1 --profile > gpurun_out/ncu_launch.log 2>&1; echo "exit $?"; wc -l gpurun_out/launches_r1.csv
