"""compute-sanitizer target: one launch of every form of the tensor-core convolution on small batches."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402

eng = Engine(make_state_dict(0), device=0, max_batch=8, max_tokens=32769)
for T, B in ((3000, 5), (8193, 3), (16385, 2), (20000, 3), (32769, 1)):
    Tp = (T + 127) // 128 * 128
    vx = torch.zeros(B, 256, Tp, dtype=torch.float16, device="cuda")
    x0 = torch.zeros(B, 256, Tp, dtype=torch.bfloat16, device="cuda")
    vx[..., :T] = torch.randn(B, 256, T, device="cuda").half() * 0.3
    x0[..., :T] = torch.randn(B, 256, T, device="cuda").bfloat16()
    out = eng.longconv_tc(1, vx, x0, T)
    torch.cuda.synchronize()
    print(T, B, float(out.float().abs().mean()))
