"""block_mlp (production form, K2 size) under engine options: event timing + CTA 0 timeline per setting.
usage: python profiles/trace_block_mlp_opts.py "name=value,name=value" ...   (one run per argument; "" = defaults)"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from chimeralm_b200.engine import Engine, _stream_ptr
from chimeralm_b200.weights import make_state_dict

B, T = 32, 8193
Tp = (T + 63) // 64 * 64
eng = Engine(make_state_dict(0), max_batch=B, max_tokens=T)
y = torch.randn(B, 256, Tp, device="cuda").to(torch.bfloat16)
res = torch.randn((B * T + 160) * 256, device="cuda")
tr = torch.zeros(3, 64, dtype=torch.int64, device="cuda")
st = _stream_ptr(eng.device)
for setting in sys.argv[1:] or [""]:
    opts = [kv.split("=") for kv in setting.split(",") if kv]
    for k, v in opts:
        eng._check(eng.lib.clm_set_option(eng.ctx, k.encode(), int(v)), "opt")
    args = (eng.ctx, 1, C.c_void_p(y.data_ptr()), C.c_void_p(res.data_ptr()), B, T, Tp, 1)
    for _ in range(3):
        eng._check(eng.lib.clm_block_mlp_cm_trace(*args, None, st), "warm")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng._check(eng.lib.clm_block_mlp_cm_trace(*args, None, st), "run")
    e1.record()
    torch.cuda.synchronize()
    print(f"[{setting}]: {e0.elapsed_time(e1) / 10:.3f} ms per launch ({B * T} tokens)")
    tr.zero_()
    eng._check(eng.lib.clm_block_mlp_cm_trace(*args, C.c_void_p(tr.data_ptr()), st), "trace")
    torch.cuda.synchronize()
    t = tr.cpu().clone()
    w = [int(x) for x in t[0][32:39]]
    print(f"  MMA-thread waits: weights {w[0]}, gelu(h) ready {w[1]}, H drained {w[2]}, tile-level {w[3]}, total {w[4]}; "
          f"issuing fc1 groups {w[5]}, fc2 groups {w[6]} (whole launch, CTA 0)")
    t[0][32:39] = 0
    t0 = int(t[t > 0].min())
    for role, name in ((1, "mma"), (2, "epilogue")):
        v = [int(x) - t0 for x in t[role] if x > 0]
        print(" ", name, v)
    for k, v in opts:
        eng._check(eng.lib.clm_set_option(eng.ctx, k.encode(), 0), "opt")
