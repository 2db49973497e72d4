"""Numerical emulation (CPU, torch) of a tensor-core Monarch FFT long conv with fp16 operands and
fp32 accumulation: N = 16384 = 128 x 128, two reads per complex transform, operands rounded to
fp16 before every matrix product - compared with a float64 direct FFT conv on REAL layer inputs
(vx = v * x1 and the implicit filter of each layer, from the oracle).  Test infrastructure (it imports `oracle/`), not
part of the product; run by hand: `python tests/probes/fp16_monarch_fft_emulation.py [g16]`."""
import math
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from chimeralm_b200.config import DEFAULT_CONFIG as cfg  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402
from oracle import hyena_oracle as O  # noqa: E402

torch.manual_seed(0)
sd = {k: torch.as_tensor(v) for k, v in make_state_dict(0).items()}
T = 8192
ids = torch.randint(7, 11, (2, T))
N, R = 16384, 128


def q(x, dt):
    return x.to(dt).to(torch.float64)


def cmatmul(ar, ai, br, bi):  # fp32-accumulate emulated in float64 (error dominated by operand rounding)
    return ar @ br - ai @ bi, ar @ bi + ai @ br


G16 = len(sys.argv) > 1


def monarch_conv(xa, xb, k, dt, s1=1 / 8, s3=1.0):
    """xa, xb: [T] float64 (bf16-representable), k: [T] float64 filter (bias folded in tap 0)."""
    n = torch.arange(R, dtype=torch.float64)
    ang = -2 * math.pi * torch.outer(n, n) / R
    Fr, Fi = q(torch.cos(ang), dt), q(torch.sin(ang), dt)
    tang = -2 * math.pi * torch.outer(n, n) / N  # [k1, n2]
    twr, twi = torch.cos(tang).float().double(), torch.sin(tang).float().double()
    Zr = torch.zeros(R, R, dtype=torch.float64); Zi = torch.zeros(R, R, dtype=torch.float64)
    Zr.view(-1)[:T] = xa; Zi.view(-1)[:T] = xb              # [n1][n2]
    Zr, Zi = q(Zr, dt), q(Zi, dt)
    Ar, Ai = cmatmul(Fr, Fi, Zr, Zi)                          # [k1][n2]
    Ar, Ai = Ar * s1, Ai * s1
    Ar, Ai = Ar * twr - Ai * twi, Ar * twi + Ai * twr
    Ar, Ai = q(Ar, dt), q(Ai, dt)
    Sr, Si = cmatmul(Ar, Ai, Fr, Fi)                          # [k1][k2], freq = k1 + 128 k2
    Sr, Si = Sr * s3, Si * s3
    kp = torch.zeros(N, dtype=torch.float64); kp[:T] = k
    G = torch.fft.fft(kp) / (N * s1 * s3)
    G = G.view(R, R).t()                                      # G[k1][k2] = G[k1 + 128 k2]
    Gr, Gi = (q(G.real, dt), q(G.imag, dt)) if G16 else (G.real.float().double(), G.imag.float().double())
    Pr, Pi = Sr * Gr - Si * Gi, Sr * Gi + Si * Gr
    mx = max(Ar.abs().max().item(), Ai.abs().max().item()), max(Pr.abs().max().item(), Pi.abs().max().item())
    Pr, Pi = q(Pr, dt), q(Pi, dt)
    Br, Bi = cmatmul(Pr, Pi, Fr, -Fi)                         # [k1][n2]
    Br, Bi = Br * twr + Bi * twi, Bi * twr - Br * twi
    mx = mx + (max(Br.abs().max().item(), Bi.abs().max().item()),)
    Br, Bi = q(Br, dt), q(Bi, dt)
    zr, zi = cmatmul(Fr, -Fi, Br, Bi)                         # [n1][n2]
    return zr.reshape(-1)[:T], zi.reshape(-1)[:T], mx


def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)


h = F.embedding(ids, sd[O.BB + "embeddings.word_embeddings.weight"])
for layer in range(cfg.n_layer):
    p = f"{O.BB}layers.{layer}."
    res = h.float()
    x = F.layer_norm(res, (256,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.layer_norm_epsilon)
    u = F.linear(x, sd[p + "mixer.in_proj.weight"], sd[p + "mixer.in_proj.bias"]).transpose(1, 2)
    uc = F.conv1d(u, sd[p + "mixer.short_filter.weight"], sd[p + "mixer.short_filter.bias"], padding=2, groups=768)[..., :T]
    x0, x1, v = uc.split(256, dim=1)
    vx = bf(v * x1)                                           # what block_in emits today (bf16)
    k = O.implicit_filter(sd, layer, T, cfg).transpose(0, 1).double()
    D = sd[p + "mixer.filter_fn.bias"].double()
    worst = {}
    for dt in (torch.float16, torch.bfloat16):
        errs, rels, mxs = [], [], []
        for c in range(0, 256, 37):
            kk = k[c].clone(); kk[0] += D[c]
            ref = torch.fft.irfft(torch.fft.rfft(vx[:, c], n=N) * torch.fft.rfft(kk, n=N), n=N)[..., :T]
            ya, yb, mx = monarch_conv(vx[0, c], vx[1, c], kk, dt)
            e = torch.stack([ya - ref[0], yb - ref[1]])
            errs.append(e.abs().max().item()); rels.append((e.norm() / ref.norm()).item()); mxs.append(mx)
            bferr = (bf(ref) - ref).abs().max().item()
        worst[dt] = (max(errs), max(rels), [max(m[i] for m in mxs) for i in range(3)])
    print(f"layer {layer}: |vx|max {vx.abs().max():.3f} rms {vx.pow(2).mean().sqrt():.3f}  |y|max {ref.abs().max():.3f}  "
          f"bf16 rounding of y: max abs {bferr:.2e}")
    for dt, (e, r, m) in worst.items():
        print(f"    {str(dt):16s} max abs err {e:.2e}  rel L2 err {r:.2e}  stage maxima (A, P, B) {m[0]:.1f} {m[1]:.3f} {m[2]:.3f}")
    h = O.block(sd, layer, h, cfg)
