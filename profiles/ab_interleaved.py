"""Interleaved A/B of one engine option on K2-sized forwards (clock drift hits both arms alike):
python profiles/ab_interleaved.py option [B T] -> per-kernel ms/step with the option at 1 and at 0, four alternations."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402

# --noprof: no per-kernel events between the launches (they serialise the stream and switch programmatic dependent launch off):
# the step time alone
NOPROF = "--noprof" in sys.argv
if NOPROF:
    sys.argv.remove("--noprof")
opt = sys.argv[1]
B, T = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (32, 8193)
eng = Engine(make_state_dict(0), max_batch=B, max_tokens=T)
ids = torch.randint(7, 11, (B, T), dtype=torch.uint8, device="cuda")
for _ in range(100):   # clocks and power state settle
    eng.forward(ids)
res = {0: [], 1: []}
n = 20
for rep in range(4):
    for val in (1, 0):
        eng.set_option(opt, val)
        for _ in range(3):
            eng.forward(ids)
        if not NOPROF:
            eng.profile(True)
            eng.profile_reset()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            eng.forward(ids)
        e1.record()
        torch.cuda.synchronize()
        prof = {}
        if not NOPROF:
            prof = {k: v[0] / n for k, v in eng.profile_read().items()}
            eng.profile(False)
        prof["step"] = e0.elapsed_time(e1) / n
        res[val].append(prof)
for val in (1, 0):
    keys = sorted(res[val][0], key=lambda k: -res[val][0][k])
    print(f"{opt}={val}: " + "  ".join(f"{k} {sum(r[k] for r in res[val]) / len(res[val]):.3f}" for k in keys))
