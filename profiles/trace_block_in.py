"""Print CTA 0's clock64 timeline of the fused block-in kernel (run on the GPU box)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine, _stream_ptr
from chimeralm_b200.weights import make_state_dict

eng = Engine(make_state_dict(0), max_batch=32, max_tokens=8193)
B, T = 32, 8193
Tp = (T + 63) // 64 * 64
res = torch.randn(B * T + 160, 256, device="cuda")
vx = torch.zeros(B, 256, Tp, dtype=torch.bfloat16, device="cuda")
x0 = torch.zeros_like(vx)
tr = torch.zeros(2, 64, dtype=torch.int64, device="cuda")
for _ in range(2):
    tr.zero_()
    eng._check(eng.lib.clm_block_in_trace(eng.ctx, 1, C.c_void_p(res.data_ptr()), B, T, Tp, C.c_void_p(vx.data_ptr()),
                                          C.c_void_p(x0.data_ptr()), C.c_void_p(tr.data_ptr()), _stream_ptr(eng.device)), "trace")
    torch.cuda.synchronize()
t = tr.cpu()
t0 = int(t[t > 0].min())
print("per tile - mma: [tile start, xn_full seen, (acc_free[0] seen, pass issued) x2]; epilogue: [tile start, (pass start = staging free, set A accumulated, "
      "x0 staged, set B accumulated, first half of set B convolved, staging free again, pass done) x2]")
for role, name in ((0, "mma"), (1, "epilogue(warp2)")):
    v = [int(x) - t0 for x in t[role] if x > 0]
    print(name, len(v))
    print("  ", v)
