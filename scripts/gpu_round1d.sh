#!/bin/bash
# one GPU call: all GPU tests, smoke, fused-output-block microbench, bench with / without the fused output blocks
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout -k 5 700 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu.log | head -30 | cut -c1-300
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== resln microbench"; timeout 200 python scripts/gpu_resln_microbench.py 2>&1 | tail -4
for FR in 0 1; do
echo "== bench graph fuse_residual=$FR"; timeout -k 5 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --fuse-residual $FR > gpurun_out/bench_fr$FR.json 2> gpurun_out/bench_fr$FR.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_fr$FR.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|run_backward" gpurun_out/bench_fr$FR.err | tail -5 | cut -c1-300
done
