"""Throughput of the forward by read length at a fixed padded-token budget (the K3 bucket size): where mixed-length
workloads lose against K2.  python profiles/length_sweep.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from chimeralm_b200.engine import Engine  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402

BUDGET = 32 * 8193
eng = Engine(make_state_dict(0), device=0, max_batch=256, max_tokens=32769, token_budget=BUDGET)
print(f"{'T':>6s} {'B':>4s} {'conv kernel':>16s} {'ms/batch':>9s} {'Mtok/s':>8s}   kernel ms per batch")
for T in (tuple(int(a) for a in sys.argv[1:]) or (1025, 1100, 1537, 2049, 3073, 4096, 4097, 5121, 6145, 8193, 8201, 12289, 16385, 20481, 24577, 32769)):
    B = max(1, min(256, BUDGET // T))
    ids = torch.randint(7, 11, (B, T), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        eng.forward(ids)
    n = 8
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng.forward(ids)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    eng.profile_reset()
    eng.profile(True)
    for _ in range(4):
        eng.forward(ids)
    torch.cuda.synchronize()
    prof = {k: round(v[0] / 4, 3) for k, v in eng.profile_read().items()}
    eng.profile(False)
    print(f"{T:6d} {B:4d} {eng.longconv_variant(T):>16s} {ms:9.3f} {B * T / ms / 1e3:8.1f}   {prof}", flush=True)
eng.close()
