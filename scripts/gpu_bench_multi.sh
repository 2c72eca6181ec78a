#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
echo "== $N-GPU bench graph B=256"; timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_${N}gpu.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","execution","n_gpus","gpu_launches","grad_allreduce_bytes_per_step","clocks")})
    print("e2e", round(d["e2e"]["value"],1))
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|warn\|run_backward\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_${N}gpu.err | tail -8 | cut -c1-300
