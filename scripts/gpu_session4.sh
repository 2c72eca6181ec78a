#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -k 5 300 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== gemm timing"; timeout -k 10 300 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]" gpurun_out/debug_gemm.log | head -20
for B in 64 128; do
echo "== bench B=$B"; timeout -k 5 400 python bench.py --steps 8 --warmup 3 --batch $B $( [ $B != 64 ] && echo --no-cpu-baseline ) > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b$B.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],3))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/bench_b$B.err; done
echo "== bench eager B=64"; timeout -k 5 300 python bench.py --steps 5 --warmup 3 --batch 64 --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "exit $?"; cut -c1-200 gpurun_out/bench_eager.json
echo "== microbench"; timeout -k 10 300 python scripts/gpu_microbench.py > gpurun_out/microbench.log 2>&1; echo "exit $?"; grep -E "bwd|S': (1|4)," gpurun_out/microbench.log | head -30
echo "== ncu full"; timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"sample_kl|bayes_gemm|bayes_wgrad" -c 14 -f -o gpurun_out/prof_r1 python scripts/profile_target.py > gpurun_out/ncu_full.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_full.log
