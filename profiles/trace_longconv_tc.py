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
    trace = torch.zeros(9, 64, dtype=torch.int64, device="cuda")   # row 0 MMA issuer, rows 1.. epilogue warps
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
    ew = [i for i in range(1, 9) if int((t[i] > 0).sum()) > 0]
    if len(ew) > 1:   # spread over the eight epilogue warps, per slot: (min .. max) of start / barrier passed / end
        vs = [[int(x) - base for x in t[i].tolist() if x > 0] for i in ew]
        m = min(len(v) for v in vs) // 3
        print("   epilogue warps, per slot: start min..max | wait over min..max | end min..max")
        for k in range(m):
            c = [[v[3 * k + j] for v in vs] for j in range(3)]
            print(f"   slot {k:2d}: {min(c[0]):6d}..{max(c[0]):6d} | {min(c[1]):6d}..{max(c[1]):6d} | {min(c[2]):6d}..{max(c[2]):6d}"
                  f"   work per warp: " + " ".join(f"{e - w:5d}" for w, e in zip(c[1], c[2])))
    for row, name in ((0, "mma     "), (1, "epilogue")):
        v = [int(x) - base for x in t[row].tolist() if x > 0]
        print(name, v)
        # two-items-in-flight kernel: three stamps per executed slot (start, barrier wait over, end)
        order = (["M1 A", "M7 B-", "M3 A", "M1 B", "M5 A", "M3 B", "M7 A", "M5 B"] if row == 0 else  # noqa
                 ["E3 B-", "E1 A", "E4 B-", "E2 A", "E1 B", "E3 A", "E2 B", "E4 A"])
        slots = [o for k in range(8) for o in order if not (k == 0 and o.endswith("B-"))]
        prev_end = None
        for i in range(0, len(v) - 2, 3):
            s0, s1, s2 = v[i:i + 3]
            gap = "" if prev_end is None else f" gap {s0 - prev_end:5d}"
            print(f"   {slots[i // 3]:6s} start {s0:6d} wait {s1 - s0:5d} work {s2 - s1:5d}{gap}")
            prev_end = s2
