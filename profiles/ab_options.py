"""A/B timing of whole forward passes (K2 size) under engine options: python profiles/ab_options.py name=value ..."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine
from chimeralm_b200.weights import make_state_dict

B, T = 32, 8193
eng = Engine(make_state_dict(0), max_batch=B, max_tokens=T)
ids = torch.randint(7, 11, (B, T), dtype=torch.uint8, device="cuda")


def run(label):
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
    prof = {k: round(v[0] / n, 3) for k, v in eng.profile_read().items()}
    eng.profile(False)
    print(f"{label}: {e0.elapsed_time(e1) / n:.3f} ms/step  {prof}")


run("default")
for arg in sys.argv[1:]:
    name, val = arg.split("=")
    eng.set_option(name, int(val))
    run(arg)
    eng.set_option(name, 1 - int(val) if int(val) in (0, 1) else 0)
