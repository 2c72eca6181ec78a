#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` report of scripts/profile_r02.py: per-launch DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernels at the shapes bench.py launches them with.
bench.py reads this file for `roofline.traffic` (never a hand-copied constant).
    python scripts/ncu_traffic.py gpurun_out/prof_r02.ncu-rep profiles/traffic.json"""
import csv
import datetime
import io
import json
import os
import subprocess
import sys

KEYS = {  # kernel-name fragment -> (key in traffic.json, shape, algorithmic bytes)
    "bayes_gemm2_gelu_kernel": ("bert_cls:gemm_fwd_ffn_up", "S=4, M=65536, N=3072, K=768 (x, w read; z, y written, bf16)",
                                4 * (65536 * 768 + 3072 * 768 + 2 * 65536 * 3072) * 2),
    "bayes_gemm2_dgelu_kernel": ("bert_cls:gemm_dgrad_ffn_down_gelu", "S=4, M=65536, N=768, K=3072 (gy, w, z read; gz written, bf16)",
                                 4 * (65536 * 768 + 3072 * 768 + 2 * 65536 * 3072) * 2),
    "bayes_gemm2_kernel": ("bert_cls:gemm_fwd_qkv", "S=4, M=65536, N=768, K=768 (x, w read; y written, bf16)",
                           4 * (2 * 65536 * 768 + 768 * 768) * 2),
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    head = rows[0]
    idx = {n: i for i, n in enumerate(head)}
    units = rows[1]

    def val(r, m):
        v = float(r[idx[m]].replace(",", ""))
        u = units[idx[m]].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "us": 1, "ms": 1e3, "ns": 1e-3, "second": 1e6, "s": 1e6}.get(u, 1)

    res = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        for frag, (key, shape, alg) in KEYS.items():
            if frag in name and key not in res:
                rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
                res[key] = {"kernel": name.split("(")[0].replace("void ", ""), "shape": shape,
                            "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
                            "algorithmic_bytes": alg, "duration_us": val(r, "gpu__time_duration.sum"),
                            "report": os.path.basename(rep), "captured": datetime.date.today().isoformat(),
                            "how": "ncu --set full --clock-control none, one launch, scripts/profile_r02.py"}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
