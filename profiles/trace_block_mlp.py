"""Print CTA 0's clock64 timeline of the fused block kernel (run on the GPU box)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine, _stream_ptr
from chimeralm_b200.weights import make_state_dict

eng = Engine(make_state_dict(0), max_batch=4, max_tokens=1024)
M = 148 * 128 * 3
y = torch.randn(M, 256, device="cuda").to(torch.bfloat16)
res = torch.randn(M + 128, 256, device="cuda")
tr = torch.zeros(3, 64, dtype=torch.int64, device="cuda")
for _ in range(2):
    tr.zero_()
    eng._check(eng.lib.clm_block_mlp_trace(eng.ctx, 1, C.c_void_p(y.data_ptr()), C.c_void_p(res.data_ptr()), M,
                                           C.c_void_p(tr.data_ptr()), _stream_ptr(eng.device)), "trace")
    torch.cuda.synchronize()
t = tr.cpu()
t0 = int(t[t > 0].min())
names = {0: "producer(x issued per tile)", 1: "mma", 2: "epilogue(warp2)"}
w = [int(x) for x in t[0][32:37]]
print(f"MMA-thread waits (cycles, whole kernel): weights {w[0]}, gelu(h) ready {w[1]}, H drained {w[2]}, tile-level (r_free/xn_full) {w[3]}, total {w[4]}")
t[0][32:37] = 0
for role in range(3):
    v = [int(x) - t0 for x in t[role] if x > 0]
    print(names[role], len(v))
    print("  ", v)
