#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu -k resln/output/layernorm"; timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 150 -k "resln or output_blocks or layernorm" > gpurun_out/pytest_sel.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_sel.log | head -30 | cut -c1-300
echo "== resln microbench (staged)"; FUSED_ONLY=1 timeout 200 python scripts/gpu_resln_microbench.py 2>&1 | tail -1
echo "== resln microbench H=1024 (staged)"; FUSED_ONLY=1 H=1024 timeout 200 python scripts/gpu_resln_microbench.py 2>&1 | tail -1
echo "== bench graph (defaults)"; timeout -k 5 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_g.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|run_backward" gpurun_out/bench_g.err | tail -5 | cut -c1-300
