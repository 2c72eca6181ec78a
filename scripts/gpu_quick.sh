#!/bin/bash
# usage: gpu_quick.sh "<pytest -k expr>" [batch]   -- selected GPU tests, then one graph-mode bench run
mkdir -p gpurun_out
echo "== pytest gpu -k '$1'"; timeout -k 5 500 python -m pytest tests -m gpu -q --timeout 120 -k "$1" > gpurun_out/pytest_sel.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed" gpurun_out/pytest_sel.log | head -20 | cut -c1-300
B=${2:-256}
echo "== bench graph B=$B"; timeout -k 5 600 python bench.py --steps 8 --warmup 3 --batch $B --no-cpu-baseline > gpurun_out/bench_g$B.json 2> gpurun_out/bench_g$B.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_g$B.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|run_backward" gpurun_out/bench_g$B.err | tail -5 | cut -c1-300
