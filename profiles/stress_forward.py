"""Stress / determinism run across shapes with copy traffic on a side stream: python profiles/stress_forward.py"""
import sys, torch
sys.path.insert(0, "/root/repo")
from chimeralm_b200.engine import Engine
from chimeralm_b200.weights import make_state_dict
sd = make_state_dict(0)
side = torch.cuda.Stream()
host = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
dev = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
for B, T, n in ((32, 8193, 600), (16, 16385, 200), (8, 32769, 100), (31, 5000, 300), (3, 8200, 300), (64, 4097, 200), (85, 3073, 200),
                (255, 1025, 200), (127, 2049, 200), (102, 2560, 200), (21, 12289, 200), (13, 20000, 100), (64, 4096, 200), (5, 32769, 100),
                (40, 3000, 200), (9, 2064, 200), (70, 200, 200)):   # 56-token tails (gathered only), 16-token tails, short reads
    eng = Engine(sd, device=0, max_batch=B, max_tokens=T)
    ids = torch.randint(7, 11, (B, T), dtype=torch.uint8, device="cuda")
    first = eng.forward(ids).clone()
    bad = 0
    for i in range(n):
        if i % 2 == 0:
            with torch.cuda.stream(side):
                dev.copy_(host, non_blocking=True); host.copy_(dev, non_blocking=True)
        out = eng.forward(ids)
        if i % 20 == 19 and not torch.equal(out, first):
            bad += 1
    torch.cuda.synchronize()
    eng.forward_status()   # the last forward's status word: raises on a token-range / fp16-range flag
    print(f"B={B} T={T}: {n} forwards, conv={eng.longconv_variant(T)}, mismatches={bad}")
    eng.close()
