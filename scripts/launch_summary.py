#!/usr/bin/env python
"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> markdown table by kernel:
    python scripts/launch_summary.py launches.csv "title" > out.md
Kernels whose name contains one of OURS are marked as hand-written (this repo's .so)."""
import collections
import csv
import sys

OURS = ("attn_tc::", "attn::", "colsum_reduce", "bayes_gemm", "bayes_wgrad", "wgrad_reduce", "sample_kl", "layernorm_", "resln_", "bias_grad", "gemm_f32_kernel", "embedding_", "split_bf16x2",
        "clip_adamw", "grad_sumsq", "gelu_bwd_bias_grad", "philox_normal", "dropout_mask_kernel")
CONTRACTIONS = ("bayes_gemm", "bayes_wgrad_kernel")

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
head = rows[0]
iN, iV = head.index("Kernel Name"), head.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[iN].replace("void ", "").replace("<unnamed>::", "").replace("at::", "")
    d = agg.setdefault(name, [0.0, 0])
    d[0] += float(r[iV].replace(",", "")) / 1e6
    d[1] += 1
total = sum(v[0] for v in agg.values())
n = sum(v[1] for v in agg.values())
print(f"{n} launches ({sys.argv[2] if len(sys.argv) > 2 else ''}), total kernel time {total:.1f} ms (cold-cache, serialised by ncu)\n")
print("| ms | share | launches | ours | kernel |\n|---|---|---|---|---|")
for name, (ms, k) in sorted(agg.items(), key=lambda t: -t[1][0])[:45]:
    ours = "yes" if any(o in name for o in OURS) else ""
    print(f"| {ms:.2f} | {100 * ms / total:.1f}% | {k} | {ours} | `{name[:80]}` |")
ours_ms = sum(ms for name, (ms, _) in agg.items() if any(o in name for o in OURS))
con_ms = sum(ms for name, (ms, _) in agg.items() if any(o in name for o in CONTRACTIONS))
print(f"\nhand-written kernels: {ours_ms:.1f} ms = {100 * ours_ms / total:.1f}% of kernel time; "
      f"tcgen05 contractions (bayes_gemm* + bayes_wgrad): {con_ms:.1f} ms = {100 * con_ms / total:.1f}%")
