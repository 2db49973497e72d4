"""Interleaved A/B of the tensor-core conv kernels on the K2 shape (conv launches only): tc_pipe=1 (two items in flight)
vs tc_pipe=0 (one item per SM).  python profiles/ab_longconv.py [B T]"""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from chimeralm_b200.engine import Engine, _stream_ptr  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402

B, T = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 8193)
eng = Engine(make_state_dict(0), device=0, max_batch=B, max_tokens=T)
Tp = (T + 127) // 128 * 128
vx = (torch.randn(B, 256, Tp, device="cuda") * 0.3).half()
x0 = torch.randn(B, 256, Tp, device="cuda").bfloat16()
out = torch.zeros_like(x0)
st = _stream_ptr(eng.device)
args = (eng.ctx, 1, C.c_void_p(vx.data_ptr()), C.c_void_p(x0.data_ptr()), C.c_void_p(out.data_ptr()), B, T, Tp)


def run(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        eng._check(eng.lib.clm_longconv_tc(*args, st), "tc")
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


run(300)   # clocks up
res = {0: [], 1: []}
for rep in range(4):
    for opt in (1, 0):
        eng.set_option("tc_pipe", opt)
        run(20)
        res[opt].append(run(200))
ref = None
for opt in (1, 0):
    eng.set_option("tc_pipe", opt)
    eng._check(eng.lib.clm_longconv_tc(*args, st), "tc")
    torch.cuda.synchronize()
    if ref is None:
        ref = out.clone()
    else:
        print("max |two-in-flight - one-item| =", (ref.float() - out.float()).abs().max().item())
for opt in (1, 0):
    v = res[opt]
    print(f"B={B} T={T} tc_pipe={opt}: " + " ".join(f"{x * 1e3:.1f}" for x in v) + f" us/launch (median {sorted(v)[len(v) // 2] * 1e3:.1f})")
