"""clock64 timeline of CTA 0 of longconv_tc_kernel + event timing of the launch (with and without tail)."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from chimeralm_b200.engine import Engine, _stream_ptr  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402

B = 16
eng = Engine(make_state_dict(0), device=0, max_batch=B, max_tokens=32769)
for T in (tuple(int(a) for a in sys.argv[1:]) or (8192, 8193)):
    Tp = (T + 127) // 128 * 128
    vx = (torch.randn(B, 256, Tp, device="cuda") * 0.3).half()
    x0 = torch.randn(B, 256, Tp, device="cuda").bfloat16()
    out = torch.zeros_like(x0)
    trace = torch.zeros(2, 64, dtype=torch.int64, device="cuda")
    st = _stream_ptr(eng.device)
    args = (eng.ctx, 1, C.c_void_p(vx.data_ptr()), C.c_void_p(x0.data_ptr()), C.c_void_p(out.data_ptr()), B, T, Tp)
    for _ in range(3):
        eng._check(eng.lib.clm_longconv_tc(*args, st), "tc")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        eng._check(eng.lib.clm_longconv_tc(*args, st), "tc")
    e1.record()
    torch.cuda.synchronize()
    print(f"T={T}: {e0.elapsed_time(e1) / 10:.3f} ms per launch")
    eng._check(eng.lib.clm_longconv_tc_trace(*args, C.c_void_p(trace.data_ptr()), st), "trace")
    torch.cuda.synchronize()
    t = trace.cpu()
    base = int(t[t > 0].min())
    for row, name in ((0, "mma     "), (1, "epilogue")):
        v = [int(x) - base for x in t[row].tolist() if x > 0]
        print(name, v[:30])
