#!/bin/bash
# One entry point for the gpurun jobs of a round: `scripts/gpu_run.sh stage [stage ...]`, outputs under gpurun_out/.
#   tests        pytest -m gpu (whole suite)                      tests2     the two-rank NCCL test only (needs --gpus 2)
#   bench        default bench.py run (headline config)           configs    the other BASELINE.json configs, short
#   launches     ncu launch list of one eager step                ncu        ncu --set full of every hand-written kernel
#   micro        kernel microbenchmarks (scripts/gpu_microbench.py and friends)
mkdir -p gpurun_out
for stage in "$@"; do
  echo "=================== $stage"
  case "$stage" in
    tests)    timeout -k 10 1500 python -m pytest tests -m gpu -q --maxfail=20 -x -p no:cacheprovider 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log ;;
    tests_all) timeout -k 10 1800 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider -s > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log ;;
    tests_models) timeout -k 10 1800 python -m pytest tests/test_gpu_models.py -m gpu -q --maxfail=30 -p no:cacheprovider -s > gpurun_out/pytest_gpu_models.log 2>&1; grep -E "^\[|passed|failed|^E  |Error" gpurun_out/pytest_gpu_models.log | head -60 ;;
    tests2)   timeout -k 10 900 python -m pytest tests/test_gpu_models.py -m gpu -q -k two_rank -s -p no:cacheprovider > gpurun_out/pytest_gpu2.log 2>&1; grep -E "MULTIRANK|Error|assert|passed|failed" gpurun_out/pytest_gpu2.log | head -20 ;;
    smoke)    timeout -k 10 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 ;;
    bench)    timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "rc $?"; cut -c1-1500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err ;;
    bench_quick) timeout -k 10 900 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --extras 0 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "rc $?"; cut -c1-400 gpurun_out/bench_quick.json; python scripts/bench_kernels.py gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err ;;
    configs)
      for c in mlp linear bert_qa bert_large; do
        timeout -k 10 900 python bench.py --config $c --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.json 2> gpurun_out/bench_$c.err; echo "$c rc $?"; cut -c1-600 gpurun_out/bench_$c.json; tail -2 gpurun_out/bench_$c.err
      done ;;
    launches) timeout -k 10 1500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 1 --warmup 1 --graph 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "rc $?"; tail -2 gpurun_out/ncu_launch.log | cut -c1-200; python scripts/launch_summary.py gpurun_out/launches.csv > gpurun_out/launches.md; head -30 gpurun_out/launches.md ;;
    attn)
      timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "attention" 2>&1 | tail -8
      timeout -k 10 300 python scripts/bench_attention.py 2>&1 | tail -3 | tee gpurun_out/bench_attention.json ;;
    gelu)
      timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "gelu" 2>&1 | tail -4
      timeout -k 10 300 python scripts/bench_gelu.py 2>&1 | tail -2 | tee gpurun_out/bench_gelu.json ;;
    ncu_attn)
      timeout -k 10 900 ncu --set full --clock-control none --import-source on -k regex:"fwd_kernel|bwd_kernel|attention_fwd|attention_bwd|sdpa|cudnn" -c 40 -f -o gpurun_out/prof_attn python scripts/profile_attention.py > gpurun_out/ncu_attn.log 2>&1; echo "rc $?"; tail -2 gpurun_out/ncu_attn.log
      rm -f gpurun_out/r02_ncu_attention.md; python scripts/ncu_summary.py gpurun_out/prof_attn.ncu-rep gpurun_out/r02_ncu_attention.md; cut -c1-260 gpurun_out/r02_ncu_attention.md ;;
    sanitize_colsum)
      timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "dgrad_gelu_kernel or tcgen05_and_mma_sync" > gpurun_out/sanitizer_colsum_memcheck.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/sanitizer_colsum_memcheck.log
      timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "dgrad_gelu_kernel and 2560 or tcgen05_and_mma_sync" > gpurun_out/sanitizer_colsum_racecheck.log 2>&1; echo "racecheck rc $?"; tail -4 gpurun_out/sanitizer_colsum_racecheck.log ;;
    sanitize_attn)
      timeout -k 10 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "attention_kernels and 3-12-128 or attention_kernels and 150-2-128" > gpurun_out/sanitizer_attn_memcheck.log 2>&1; echo "memcheck rc $?"; tail -4 gpurun_out/sanitizer_attn_memcheck.log
      timeout -k 10 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "attention_kernels and 3-12-128" > gpurun_out/sanitizer_attn_racecheck.log 2>&1; echo "racecheck rc $?"; tail -4 gpurun_out/sanitizer_attn_racecheck.log ;;
    ab_keep)
      timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "resln or fuses_hf_output" 2>&1 | tail -3
      for v in 0 1 0 1; do
        timeout -k 10 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --extras 0 --resln-keep-bits $v > gpurun_out/bench_keep$v.json 2> gpurun_out/bench_keep$v.err; echo "resln-keep-bits=$v rc $?"; python scripts/bench_kernels.py gpurun_out/bench_keep$v.json 2>/dev/null | grep "^value\|resln"
      done ;;
    ab_poly)
      for v in 0 1 0 1; do
        timeout -k 10 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --extras 0 --gelu-poly $v > gpurun_out/bench_poly$v.json 2> gpurun_out/bench_poly$v.err; echo "gelu-poly=$v rc $?"; python scripts/bench_kernels.py gpurun_out/bench_poly$v.json 2>/dev/null | head -1
      done ;;
    ab_links)
      for v in 0 1 0 1; do
        timeout -k 10 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --extras 0 --gelu-links $v > gpurun_out/bench_links$v.json 2> gpurun_out/bench_links$v.err; echo "links=$v rc $?"; python scripts/bench_kernels.py gpurun_out/bench_links$v.json | head -2
      done ;;
    multi2)
      scripts/gpu_run.sh tests2
      TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
      for spec in "2 bert_cls batch" "2 bert_qa samples" "2 bert_qa batch" "1 bert_qa batch"; do
        set -- $spec
        NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout -k 10 900 $TR --nproc-per-node $1 --master-port 2950$1 bench.py --gpus $1 --config $2 --shard $3 --steps 6 --warmup 3 --no-cpu-baseline --extras 0 > gpurun_out/scale_$2_$3_n$1.json 2> gpurun_out/scale_$2_$3_n$1.err; echo "N=$1 $2 $3 rc $?"
        grep -m3 -E "NCCL INFO (comm|Connected|NVLS)" gpurun_out/scale_$2_$3_n$1.err | cut -c1-160
        python scripts/bench_kernels.py gpurun_out/scale_$2_$3_n$1.json 2>/dev/null | head -1; python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('   allreduce', d['grad_allreduce']['launches_per_step'], d['grad_allreduce']['ms_alone'], d['grad_allreduce_bytes_per_step'])" gpurun_out/scale_$2_$3_n$1.json
      done ;;
    multi8)
      TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
      for spec in "8 bert_cls batch" "8 bert_qa samples" "8 bert_qa batch" "4 bert_qa samples" "8 bert_large samples"; do
        set -- $spec
        NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout -k 10 900 $TR --nproc-per-node $1 --master-port 2951$1 bench.py --gpus $1 --config $2 --shard $3 --steps 6 --warmup 3 --no-cpu-baseline --extras 0 > gpurun_out/scale_$2_$3_n$1.json 2> gpurun_out/scale_$2_$3_n$1.err; echo "N=$1 $2 $3 rc $?"
        grep -m2 -E "NCCL INFO (comm|NVLS)" gpurun_out/scale_$2_$3_n$1.err | cut -c1-160
        python scripts/bench_kernels.py gpurun_out/scale_$2_$3_n$1.json 2>/dev/null | head -1; python -c "import json,sys; d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print('   allreduce', d['grad_allreduce']['launches_per_step'], d['grad_allreduce']['ms_alone'], d['grad_allreduce_bytes_per_step'])" gpurun_out/scale_$2_$3_n$1.json
      done ;;
    ab_attn)
      for v in 0 1 0 1; do
        timeout -k 10 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --extras 0 --attention $v > gpurun_out/bench_attn$v.json 2> gpurun_out/bench_attn$v.err; echo "attention=$v rc $?"; python scripts/bench_kernels.py gpurun_out/bench_attn$v.json 2>/dev/null | head -2
      done
      python scripts/bench_kernels.py gpurun_out/bench_attn1.json 2>/dev/null | head -14 ;;
    ncu2)
      timeout -k 10 1500 ncu --set full --clock-control none --import-source on -k regex:"gelu_kernel|dgelu_kernel|bayes_gemm2_kernel|bayes_gemm_kernel|sample_kl_multi_kernel|embedding_|attention_fwd|attention_bwd|split_bf16" -c 14 -f -o gpurun_out/prof_r02 python scripts/profile_r02.py > gpurun_out/ncu_r02.log 2>&1; echo "rc $?"; tail -2 gpurun_out/ncu_r02.log
      rm -f gpurun_out/r02_ncu_full_kernels.md; python scripts/ncu_summary.py gpurun_out/prof_r02.ncu-rep gpurun_out/r02_ncu_full_kernels.md; cat gpurun_out/r02_ncu_full_kernels.md | cut -c1-330
      python scripts/ncu_traffic.py gpurun_out/prof_r02.ncu-rep gpurun_out/traffic.json | head -40
      ls -la gpurun_out/prof_r02.ncu-rep ;;
    sanitize)
      timeout -k 10 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "dgrad_gelu and 2560 or x3_contractions and 200 or attention_kernels and 5-2-16 or attention_kernels and 2-4-64 or fused_wgrad_epilogue and 200 or embedding_and_layernorm" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?"; tail -6 gpurun_out/sanitizer_memcheck.log
      timeout -k 10 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -x -k "attention_kernels and 5-2-16 or attention_kernels and 2-4-64 or resln_vs_torch and 33 or tc_contractions and 128-256-64" > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?"; tail -6 gpurun_out/sanitizer_racecheck.log ;;
    micro)    timeout -k 10 600 python scripts/gpu_microbench.py 2>&1 | tail -40 ;;
    *) echo "unknown stage $stage" ;;
  esac
done
