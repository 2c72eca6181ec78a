#!/bin/bash
mkdir -p gpurun_out
echo "== sdpa capture"; timeout -k 5 300 python scripts/gpu_sdpa_capture.py > gpurun_out/sdpa_capture.log 2>&1; echo "exit $?"; grep -v Warning gpurun_out/sdpa_capture.log | tail -12
for B in 128 256; do
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
echo "== launch list B=128 (ncu, 1 step)"; timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python bench.py --profile --steps 1 --warmup 1 --batch 128 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "exit $?"; tail -2 gpurun_out/ncu_launch.log | cut -c1-200
