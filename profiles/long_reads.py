"""Whole-forward timing at long-read shapes (B x T) with per-kernel breakdown: python profiles/long_reads.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine
from chimeralm_b200.weights import make_state_dict

sd = make_state_dict(0)
for B, T in ((16, 32769), (32, 16385), (64, 4097)):
    eng = Engine(sd, max_batch=B, max_tokens=T)
    ids = torch.randint(7, 11, (B, T), dtype=torch.uint8, device="cuda")
    for chunked in (1, 0):
        eng.set_option("tc_chunked", chunked)
        if not chunked and T <= 8200:
            eng.set_option("tc_conv", 0)
        for _ in range(3):
            eng.forward(ids)
        eng.profile(True)
        eng.profile_reset()
        n = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.forward(ids)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        prof = {k: round(v[0] / n, 3) for k, v in eng.profile_read().items()}
        eng.profile(False)
        print(f"B={B} T={T} conv={eng.longconv_variant(T)}: {ms:.3f} ms/step  {B / ms * 1e3:,.0f} reads/s  {B * T / ms / 1e3:,.1f} M tokens/s  {prof}")
    eng.close()
