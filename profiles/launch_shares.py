#!/usr/bin/env python
"""Per-kernel share of a step from the ncu launch list next to bench.py's CUDA-event shares.
Usage: python profiles/launch_shares.py profiles/r2_v3_launches.csv profiles/r2_v3_bench.json > profiles/r2_v3_launch_shares.txt"""
import collections
import csv
import json
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
bench = json.load(open(sys.argv[2]))
STEP = {"block_mlp_kernel": "block_mlp", "longconv_tc2_kernel": "longconv", "longconv_tc_kernel": "longconv", "longconv_fast_kernel": "longconv",
        "block_in_kernel": "block_in", "score_pool_kernel": "gemm_score", "head_fused_kernel": "head", "embed_kernel": "embed",
        "encode_kernel": "encode", "embed_in_kernel": "embed_in", "gather_tails_kernel": "block_mlp"}
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"^void ", "", r[4]).replace("clm::", "")
    short = re.sub(r"\(.*", "", name)
    key = next((k for k in STEP if short.startswith(k)), None)
    if key is None:
        continue
    a = agg.setdefault(short, [0, 0.0, STEP[key]])
    a[0] += 1
    a[1] += float(r[14]) / 1e3
tot = sum(a[1] for a in agg.values())
kms = bench["kernel_ms_per_step"]
ktot = sum(kms.values())
print(f"# ncu launch list of `python bench.py --steps 2 --warmup 1 --k3-reads 0 --k5-steps 0 --no-labels --cpu-sample 0` ({sys.argv[1].split('/')[-1]}):")
print("# gpu__time_duration.sum per step kernel (one-time clm_finalize kernels left out).  Cold-cache, serialised launches: the SHARES are what is")
print(f"# comparable with bench.py's CUDA-event shares ({sys.argv[2].split('/')[-1]}, kernel_ms_per_step), not the absolute times.")
print(f"{'kernel':44s} {'launches':>8s} {'avg us':>9s} {'ncu share':>10s} {'bench share':>12s}")
for short, (n, us, key) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{short:44s} {n:8d} {us / n:9.1f} {us / tot:10.3f} {kms.get(key, 0.0) / ktot:12.3f}")
