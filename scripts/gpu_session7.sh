#!/bin/bash
mkdir -p gpurun_out
echo "== graph debug (anomaly)"; BF_ANOMALY=1 timeout -k 5 300 python bench.py --steps 3 --warmup 3 --batch 8 --layers 1 --graph 1 --no-cpu-baseline > gpurun_out/bench_graphdbg.json 2> gpurun_out/bench_graphdbg.err; echo "exit $?"; grep -n "File \"\|Error\|error" gpurun_out/bench_graphdbg.err | head -60
