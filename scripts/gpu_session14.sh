#!/bin/bash
mkdir -p gpurun_out
echo "== tc tests, CTA pairs forced"; BF_GEMM_2CTA=2 timeout -k 5 300 python -m pytest tests -m gpu -q --timeout 60 -x -k "tc_contractions or linear_bf16 or tiny_bert_bf16 or presample" > gpurun_out/pytest_2cta.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_2cta.log | cut -c1-300
echo "== pytest gpu (all)"; timeout -k 5 400 python -m pytest tests -m gpu -q --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "exit $?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
echo "== smoke"; timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
for B in 64 256; do
echo "== bench graph B=$B"; timeout -k 5 600 python bench.py --steps 8 --warmup 3 --batch $B $( [ $B != 256 ] && echo --no-cpu-baseline ) > gpurun_out/bench_g$B.json 2> gpurun_out/bench_g$B.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_g$B.json"))
    print({k:d[k] for k in ("value","ms_per_step","execution","step_frac_of_gemm_roofline","gpu_launches","clocks")})
    print("roofline", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["roofline"].items() if k in ("achieved","frac","share_of_step")}, "e2e", round(d["e2e"]["value"],1), "cpu", d["cpu_baseline"])
    print({k:round(v["ms_per_step"],2) for k,v in d["kernels"].items()})
except Exception as e: print("ERR", e)
PY
tail -3 gpurun_out/bench_g$B.err | cut -c1-300; done
echo "== reference arm"; timeout -k 5 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cut -c1-400 gpurun_out/bench_ref.json
