#!/bin/bash
mkdir -p gpurun_out
for G in 1 0; do
echo "== 2-GPU bench graph=$G B=64"; timeout -k 5 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 --batch 64 --graph $G > gpurun_out/bench_2gpu_g$G.json 2> gpurun_out/bench_2gpu_g$G.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_2gpu_g$G.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","execution","n_gpus","gpu_launches","grad_allreduce_bytes_per_step")})
    print("e2e", round(d["e2e"]["value"],1))
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|warn" gpurun_out/bench_2gpu_g$G.err | tail -8 | cut -c1-300; done
