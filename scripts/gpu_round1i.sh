#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu -k accumulate/output_blocks/sample_kl_forward/graph"; timeout -k 5 500 python -m pytest tests -m gpu -q --timeout 150 -k "accumulate or output_blocks or sample_kl_forward or graph or presample" > gpurun_out/pytest_sel.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_sel.log | head -30 | cut -c1-300
for GS in 0 1; do
echo "== bench graph grad_sinks=$GS"; timeout -k 5 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --grad-sinks $GS > gpurun_out/bench_gs$GS.json 2> gpurun_out/bench_gs$GS.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_gs$GS.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","hbm_peak_gb")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|run_backward" gpurun_out/bench_gs$GS.err | tail -5 | cut -c1-300
done
