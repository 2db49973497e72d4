"""Interleaved comparison of several VALUES of one engine option on K2-sized forwards:
python profiles/ab_values.py option v0 v1 ... -> ms/step of the named kernel and of the step per value, three rounds."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402

opt, vals = sys.argv[1], [int(v) for v in sys.argv[2:]]
B, T = 32, 8193
eng = Engine(make_state_dict(0), max_batch=B, max_tokens=T)
ids = torch.randint(7, 11, (B, T), dtype=torch.uint8, device="cuda")
for _ in range(100):
    eng.forward(ids)
res = {v: [] for v in vals}
n = 20
for rep in range(3):
    for v in vals:
        eng.set_option(opt, v)
        for _ in range(3):
            eng.forward(ids)
        eng.profile(True)
        eng.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.forward(ids)
        e1.record()
        torch.cuda.synchronize()
        prof = {k: x[0] / n for k, x in eng.profile_read().items()}
        eng.profile(False)
        prof["step"] = e0.elapsed_time(e1) / n
        res[v].append(prof)
for v in vals:
    keys = ["step", "block_mlp", "longconv", "block_in"]
    print(f"{opt}={v}: " + "  ".join(f"{k} {sum(r[k] for r in res[v]) / len(res[v]):.3f}" for k in keys))
