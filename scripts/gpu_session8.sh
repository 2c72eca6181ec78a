#!/bin/bash
mkdir -p gpurun_out
echo "== graph debug nodes"; timeout -k 5 300 python scripts/gpu_graph_debug.py > gpurun_out/graph_nodes.log 2>&1; echo "exit $?"; grep -v Warn gpurun_out/graph_nodes.log | tail -8
echo "== pytest gpu (layernorm)"; timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120 -x -k "layernorm" > gpurun_out/pytest_ln.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_ln.log | cut -c1-300
echo "== pytest gpu (all)"; timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -8 gpurun_out/pytest_gpu.log | cut -c1-300
for B in 64 128; do
echo "== bench eager B=$B"; timeout -k 5 400 python bench.py --steps 5 --warmup 3 --batch $B --graph 0 --no-cpu-baseline > gpurun_out/bench_e$B.json 2> gpurun_out/bench_e$B.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_e$B.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/bench_e$B.err; done
