#!/bin/bash
mkdir -p gpurun_out
echo "== 2-GPU bench graph B=256"; timeout -k 5 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "exit $?"; python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/bench_2gpu.json") if l.startswith("{")][-1])
    print({k:d[k] for k in ("value","ms_per_step","execution","n_gpus","gpu_launches","grad_allreduce_bytes_per_step")})
    print("e2e", round(d["e2e"]["value"],1))
except Exception as e: print("ERR", e)
PY
grep -v "Warning\|warn" gpurun_out/bench_2gpu.err | tail -5 | cut -c1-300
echo "== 2-GPU reference arm"; timeout -k 5 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err; echo "exit $?"; cut -c1-200 gpurun_out/bench_2gpu_ref.json
