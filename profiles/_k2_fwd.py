"""Three K2-sized forwards (ncu target: pick a launch with --launch-skip)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine
from chimeralm_b200.weights import make_state_dict

B, T = 32, 8193
eng = Engine(make_state_dict(0), max_batch=B, max_tokens=T)
ids = torch.randint(7, 11, (B, T), dtype=torch.uint8, device="cuda")
for _ in range(3):
    eng.forward(ids)
torch.cuda.synchronize()
