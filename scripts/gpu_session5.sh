#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== gemm timing"; timeout -k 10 300 python scripts/gpu_debug_gemm.py > gpurun_out/debug_gemm.log 2>&1; echo "exit $?"; grep -E "TF|rc [1-9]" gpurun_out/debug_gemm.log | head -20
echo "== bench eager B=64"; timeout -k 5 300 python bench.py --steps 5 --warmup 3 --batch 64 --graph 0 --no-cpu-baseline > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_eager.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1))
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/bench_eager.err
echo "== graph debug"; TORCH_SHOW_CPP_STACKTRACES=1 timeout -k 5 300 python bench.py --steps 3 --warmup 3 --batch 8 --layers 2 --graph 1 --no-cpu-baseline > gpurun_out/bench_graphdbg.json 2> gpurun_out/bench_graphdbg.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_graphdbg.json; grep -n "cublas\|at::\|torch::autograd" gpurun_out/bench_graphdbg.err | head -40
