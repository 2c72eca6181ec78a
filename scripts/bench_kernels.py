"""Print the per-kernel table of a bench.py JSON line (file argument)."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(f"value {d['value']:.1f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  step frac {d.get('step_frac_of_gemm_roofline')}  "
      f"roofline {d['roofline']['achieved']:.0f} ({d['roofline']['frac']:.3f})  clocks {d.get('clocks')}")
sk = d.get("roofline_sample_kl")
if sk:
    print(f"sample_kl {sk['achieved']:.0f} GB/s frac {sk['frac']:.3f}")
tot = 0.0
for k, v in sorted(d.get("kernels", {}).items(), key=lambda kv: -kv[1]["ms_per_step"]):
    tot += v["ms_per_step"]
    print(f"  {k:26s} {v['calls_per_step']:6.0f} calls  {v['ms_per_step']:8.3f} ms  {100 * v['share']:5.1f} %")
print(f"  sum of instrumented kernels {tot:.2f} ms of {d['ms_per_step']:.2f}")
