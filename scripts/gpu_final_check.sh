#!/bin/bash
# what the driver runs at round end, in one go: gpu tests, smoke, default bench, reference arm
mkdir -p gpurun_out
echo "== pytest gpu (all)"; timeout -k 5 700 python -m pytest tests -m gpu -q --timeout 150 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; grep -E "^E  |passed|failed|^FAILED" gpurun_out/pytest_gpu.log | head -20 | cut -c1-300
echo "== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (defaults)"; timeout -k 5 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_default.json"))
    print({k:d[k] for k in ("value","ms_per_step","steps","warmup","execution","step_frac_of_gemm_roofline","gpu_launches","clocks","hbm_peak_gb")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step","traffic")}, "e2e", d["e2e"], "cpu", d["cpu_baseline"])
    print("sample_kl", d["roofline_sample_kl"])
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|run_backward" gpurun_out/bench_default.err | tail -5 | cut -c1-300
echo "== reference arm"; timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_ref.json
