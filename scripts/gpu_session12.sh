#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
for B in 128 256; do
echo "== bench graph B=$B"; timeout -k 5 400 python bench.py --steps 8 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench_g$B.json 2> gpurun_out/bench_g$B.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_g$B.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/bench_g$B.err | cut -c1-300; done
