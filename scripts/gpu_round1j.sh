#!/bin/bash
mkdir -p gpurun_out
for B in 384 512; do
echo "== bench graph B=$B"; timeout -k 5 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --batch $B > gpurun_out/bench_b$B.json 2> gpurun_out/bench_b$B.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_b$B.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","hbm_peak_gb")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|run_backward" gpurun_out/bench_b$B.err | tail -3 | cut -c1-300
done
