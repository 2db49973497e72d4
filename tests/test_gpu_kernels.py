"""Kernel-level parity on the B200 (-m gpu): every CUDA kernel is called through the C-ABI and
compared with a plain PyTorch fp32 statement of the same op (or the oracle's function)."""

import pytest
import torch
import torch.nn.functional as F

from chimeralm_b200 import _lib
from chimeralm_b200.config import DEFAULT_CONFIG as CFG

pytestmark = pytest.mark.gpu


def _rand_bf16(shape, gen, scale=1.0):
    return (torch.randn(shape, generator=gen) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M", [1, 128, 1000, 8193])
@pytest.mark.parametrize("NK", [(768, 256), (256, 256), (1024, 256), (256, 1024)])
def test_gemm_bias(engine, M, NK):
    N, K = NK
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A, W = _rand_bf16((M, K), g), _rand_bf16((N, K), g, 0.05)
    bias = torch.randn(N, generator=g) * 0.1
    ref = A.float() @ W.float().T + bias
    out = engine.gemm(A.cuda(), W.cuda(), bias.cuda(), _lib.EPI_BIAS_BF16)
    torch.cuda.synchronize()
    err = (out.float().cpu() - ref).abs().max().item()
    # bf16 output rounding: |ref| ~ 1 -> half-ulp 2^-9 relative
    assert err <= 1e-2 * max(1.0, ref.abs().max().item()), err


def test_gemm_epilogues(engine):
    g = torch.Generator().manual_seed(3)
    M, K = 777, 256
    A = _rand_bf16((M, K), g)
    W1, b1 = _rand_bf16((1024, K), g, 0.05), torch.randn(1024, generator=g) * 0.1
    ref = F.gelu(A.float() @ W1.float().T + b1, approximate="tanh")
    out = engine.gemm(A.cuda(), W1.cuda(), b1.cuda(), _lib.EPI_BIAS_GELU_TANH)
    assert (out.float().cpu() - ref).abs().max().item() <= 1e-2 * max(1.0, ref.abs().max().item())

    W2, b2 = _rand_bf16((256, K), g, 0.05), torch.randn(256, generator=g) * 0.1
    res = torch.randn(M, 256, generator=g)
    ref = A.float() @ W2.float().T + b2 + res
    out = engine.gemm(A.cuda(), W2.cuda(), b2.cuda(), _lib.EPI_BIAS_RES_F32, res=res.cuda())
    assert (out.cpu() - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())

    w2, bb = torch.randn(256, generator=g) * 0.1, 0.37
    ref = F.gelu(A.float() @ W2.float().T + b2) @ w2 + bb
    out = engine.gemm(A.cuda(), W2.cuda(), b2.cuda(), _lib.EPI_SCORE, w2=w2.cuda(), b2=bb)
    assert (out.cpu() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())


def test_implicit_filter(engine, state_dict):
    from oracle import hyena_oracle as O

    for layer in (0, 3):
        L = CFG.max_seq_len
        ref = O.implicit_filter(state_dict, layer, L, CFG).T  # [D, L]
        got = engine.get_filter(layer, L).cpu()
        err = (got - ref).abs().max().item()
        assert err <= 2e-5, (layer, err)


@pytest.mark.parametrize("T", [1, 37, 128, 131, 200, 515, 1000, 2049, 4096, 4100, 8193, 8292, 20000, 32769])
def test_longconv(engine, state_dict, T):
    from oracle import hyena_oracle as O

    B, D = 3, CFG.d_model
    Tp = (T + 63) // 64 * 64
    g = torch.Generator().manual_seed(T)
    vx = torch.zeros(B, D, Tp, dtype=torch.bfloat16)
    x0 = torch.zeros(B, D, Tp, dtype=torch.bfloat16)
    vx[..., :T] = _rand_bf16((B, D, T), g)
    x0[..., :T] = _rand_bf16((B, D, T), g)
    vx[..., T:] = 7.0  # garbage in the pad region must not leak into the result
    layer = 1
    k = O.implicit_filter(state_dict, layer, T, CFG).T
    bias = state_dict[f"{O.BB}layers.{layer}.mixer.filter_fn.bias"]
    ref = O.fftconv(vx[..., :T].float(), k, bias) * x0[..., :T].float()  # CPU fp32 (the oracle's op)
    out = engine.longconv(layer, vx.cuda(), x0.cuda(), T)
    torch.cuda.synchronize()
    err = (out[..., :T].float().cpu() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 1e-2 * max(1.0, scale), (T, err, scale)


@pytest.mark.parametrize("T,B", [(8192, 2), (8193, 3), (8200, 5), (4097, 2), (5000, 3), (8191, 2), (2057, 3), (3073, 2), (4096, 1),
                                 (8201, 2), (12000, 3), (16384, 2), (16385, 3), (20000, 2), (24583, 1), (32768, 2), (32769, 3),
                                 # four reads per transform (2 049 .. 4 096 tokens): every B mod 4, odd row counts
                                 (2049, 1), (2500, 5), (2056, 4), (3500, 8), (3000, 6), (4096, 7), (4000, 10), (2100, 37)])
def test_longconv_tensor_core(engine, state_dict, T, B):
    """Tensor-core FFT conv (fp16 operands, fp32 accumulate) vs the oracle's fp32 rFFT conv: max error within 1e-2 of
    the output scale (same bar as the fp32 kernels) and relative L2 error <= 2e-3 (bf16 output rounding alone is ~1e-3)."""
    from oracle import hyena_oracle as O

    D = CFG.d_model
    Tp = (T + 127) // 128 * 128
    g = torch.Generator().manual_seed(T)
    vx = torch.zeros(B, D, Tp, dtype=torch.float16)
    x0 = torch.zeros(B, D, Tp, dtype=torch.bfloat16)
    vx[..., :T] = _rand_bf16((B, D, T), g).to(torch.float16)
    x0[..., :T] = _rand_bf16((B, D, T), g)
    if 8192 <= T <= 8200:
        vx[..., T:] = 7.0  # garbage in the pad region must not leak (below 8192 tokens the contract is zeros there)
    x0[..., T:] = 3.0
    for layer in (0, 3):
        k = O.implicit_filter(state_dict, layer, T, CFG).T
        bias = state_dict[f"{O.BB}layers.{layer}.mixer.filter_fn.bias"]
        ref = O.fftconv(vx[..., :T].float(), k, bias) * x0[..., :T].float()
        out = engine.longconv_tc(layer, vx.cuda(), x0.cuda(), T)
        torch.cuda.synchronize()
        d = out[..., :T].float().cpu() - ref
        err, scale = d.abs().max().item(), ref.abs().max().item()
        rel = (d.norm() / ref.norm()).item()
        assert err <= 1e-2 * max(1.0, scale), (T, layer, err, scale)
        assert rel <= 2e-3, (T, layer, rel)


@pytest.mark.parametrize("T,B", [(8193, 7), (8200, 4), (8192, 1), (2057, 2), (6000, 33)])
def test_longconv_tensor_core_two_in_flight_matches_one_item_kernel(engine, T, B):
    """longconv_tc2_kernel (two items in flight per SM, default) against longconv_tc_kernel<false> (one item per SM,
    option tc_pipe=0) on the same inputs: the same fp16 operands go through the same products, so the two differ by fp32
    summation order of the tail tokens only.  Odd and large batches exercise the CTA item ranges (1 .. many items, with
    and without a B-type partner for the last item)."""
    D = CFG.d_model
    Tp = (T + 127) // 128 * 128
    g = torch.Generator().manual_seed(1000 + T + B)
    vx = torch.zeros(B, D, Tp, dtype=torch.float16)
    x0 = torch.zeros(B, D, Tp, dtype=torch.bfloat16)
    vx[..., :T] = _rand_bf16((B, D, T), g).to(torch.float16)
    x0[..., :T] = _rand_bf16((B, D, T), g)
    vx, x0 = vx.cuda(), x0.cuda()
    try:
        engine.set_option("tc_pack4", 0)   # (reads of <= 4 096 tokens would otherwise take the four-reads-per-item form)
        outs = []
        for pipe in (1, 0):
            engine.set_option("tc_pipe", pipe)
            outs.append(engine.longconv_tc(2, vx, x0, T)[..., :T].float().clone())
        torch.cuda.synchronize()
    finally:
        engine.set_option("tc_pipe", 1)
        engine.set_option("tc_pack4", 1)
    d = (outs[0] - outs[1]).abs()
    scale = outs[1].abs().max().item()
    assert torch.equal(outs[0][..., : min(T, 8192)], outs[1][..., : min(T, 8192)]), d.max().item()
    assert d.max().item() <= 1e-2 * max(1.0, scale)


@pytest.mark.parametrize("T,B", [(3000, 5), (4096, 4), (2100, 9)])
def test_longconv_tensor_core_four_reads_per_item_matches_two(engine, T, B):
    """Reads of 2 057 .. 4 096 tokens: the four-reads-per-item form (filter truncated to 4 096 taps, option tc_pack4=1,
    default) against the two-reads-per-item form (full 8 192-tap spectrum) - the same convolution up to fp16 rounding of
    two different spectrum tables."""
    D = CFG.d_model
    Tp = (T + 127) // 128 * 128
    g = torch.Generator().manual_seed(77 + T + B)
    vx = torch.zeros(B, D, Tp, dtype=torch.float16)
    x0 = torch.zeros(B, D, Tp, dtype=torch.bfloat16)
    vx[..., :T] = _rand_bf16((B, D, T), g).to(torch.float16)
    x0[..., :T] = _rand_bf16((B, D, T), g)
    vx, x0 = vx.cuda(), x0.cuda()
    try:
        outs = []
        for pack in (1, 0):
            engine.set_option("tc_pack4", pack)
            outs.append(engine.longconv_tc(1, vx, x0, T)[..., :T].float().clone())
        torch.cuda.synchronize()
    finally:
        engine.set_option("tc_pack4", 1)
    rel = ((outs[0] - outs[1]).norm() / outs[1].norm()).item()
    assert rel <= 3e-3, rel   # two bf16 roundings of nearly equal values


def _range_case(name, B, D, T, g):
    """v * x1 inputs far from N(0, 1): what a trained checkpoint (or a [PAD]-heavy batch) can feed the convolution."""
    x = torch.randn(B, D, T, generator=g)
    if name == "large":            # |vx| up to ~4e3 (fp16 tops out at 65 504; a 16 384-point sum of these does not fit)
        x = x * 1e3
    elif name == "dc":             # DC offset 10: the whole signal energy in one spectrum bin
        x = x + 10.0
    elif name == "tiny":           # amplitude 1e-4: below fp16's normal range without scaling
        x = x * 1e-4
    elif name == "mixed":          # every channel its own magnitude, 1e-4 .. 1e3
        x = x * (10.0 ** (torch.rand(1, D, 1, generator=g) * 7 - 4))
    elif name == "pad_prefix":     # a left-padded read: 3/4 of the row is one constant, then data
        x[..., : 3 * T // 4] = 2.5
    else:
        raise ValueError(name)
    return x.to(torch.bfloat16)


@pytest.mark.parametrize("T,B", [(8193, 2), (20000, 2), (32769, 2), (3000, 5)])   # plain, chunked (V-form tables), four reads per item
@pytest.mark.parametrize("case", ["large", "dc", "tiny", "mixed", "pad_prefix"])
def test_longconv_tensor_core_dynamic_range(engine, state_dict, case, T, B):
    """The fp16 tensor-core convolution with per-channel power-of-two input scaling (from the data here, from the
    calibration draw in the forward) and per-(segment, channel) spectrum scaling, against the oracle's fp32 rFFT
    convolution, per channel: relative L2 error <= 3e-3 for every channel (bf16 output rounding alone is ~1.5e-3)."""
    from oracle import hyena_oracle as O

    D = CFG.d_model
    Tp = (T + 127) // 128 * 128
    g = torch.Generator().manual_seed(T + len(case))
    vx = torch.zeros(B, D, Tp, dtype=torch.bfloat16)
    x0 = torch.zeros(B, D, Tp, dtype=torch.bfloat16)
    vx[..., :T] = _range_case(case, B, D, T, g)
    x0[..., :T] = _rand_bf16((B, D, T), g)
    layer = 2
    k = O.implicit_filter(state_dict, layer, T, CFG).T
    bias = state_dict[f"{O.BB}layers.{layer}.mixer.filter_fn.bias"]
    ref = O.fftconv(vx[..., :T].float(), k, bias) * x0[..., :T].float()
    out = engine.longconv_tc_auto(layer, vx.cuda(), x0.cuda(), T)
    d = out[..., :T].float().cpu() - ref
    assert torch.isfinite(out).all()
    rel = (d.pow(2).sum(dim=(0, 2)).sqrt() / ref.pow(2).sum(dim=(0, 2)).sqrt().clamp_min(1e-30))
    print(f"{case} T={T}: per-channel rel L2 max {rel.max():.2e} median {rel.median():.2e}; |ref| max {ref.abs().max():.3g}")
    assert rel.max().item() <= 3e-3, (case, T, rel.max().item(), int(rel.argmax()))
    # the fp32 FFT kernel on the same bf16 inputs is the yardstick: the tensor-core path may not be worse than 2x of it
    out32 = engine.longconv(layer, vx.cuda(), x0.cuda(), T)
    d32 = out32[..., :T].float().cpu() - ref
    rel32 = (d32.pow(2).sum(dim=(0, 2)).sqrt() / ref.pow(2).sum(dim=(0, 2)).sqrt().clamp_min(1e-30))
    assert rel.max().item() <= max(2.0 * rel32.max().item(), 2.5e-3), (rel.max().item(), rel32.max().item())


def test_longconv_tensor_core_reports_overflow(engine):
    """Raw fp16 entry (no input scaling): rows of 3e4 overflow the transform's fp16 intermediates; the auto-scaled entry
    takes the same data in its stride.  The raw entry's flag lives in the unit-level status word and must not leak into
    the next forward's status."""
    from chimeralm_b200._lib import Fp16RangeError  # noqa: F401  (documented error type of the auto entry)

    B, D, T = 2, CFG.d_model, 8193
    Tp = (T + 127) // 128 * 128
    big = torch.full((B, D, Tp), 3.0e4, dtype=torch.float16)
    x0 = torch.ones(B, D, Tp, dtype=torch.bfloat16)
    out = engine.longconv_tc(0, big.cuda(), x0.cuda(), T)
    torch.cuda.synchronize()
    assert not torch.isfinite(out[..., :T].float()).all()
    ok = engine.longconv_tc_auto(0, big.to(torch.bfloat16).cuda(), x0.cuda(), T)
    assert torch.isfinite(ok[..., :T].float()).all()
    ids = torch.randint(7, 11, (2, 300), dtype=torch.uint8)
    engine.forward(ids.cuda(), check=True)   # raises if the unit-level flag had leaked


def test_encode(engine):
    from chimeralm_b200.engine import pack_reads
    from oracle import tokenizer_oracle as TO

    seqs = ["ATCG", "", "ACGTNXacgt", "N" * 50, "ACGT" * 40, "G"]
    for add_cls, pad_left, max_len in ((True, False, 34), (False, True, 100), (True, True, 16), (False, False, 7)):
        ids_ref = [TO.encode(s, max_length=max_len, add_cls=add_cls) for s in seqs]
        padded = TO.collate(ids_ref, padding_side="left" if pad_left else "right")
        T_pad = len(padded[0])
        bases, offs = pack_reads(seqs)
        ids, lens = engine.encode(bases, offs, T_pad, add_cls=add_cls, add_sep=True, pad_left=pad_left,
                                  max_bases=max_len - 1 - int(add_cls))
        assert ids.cpu().tolist() == padded
        assert lens.cpu().tolist() == [len(x) for x in ids_ref]


@pytest.mark.parametrize("M", [128, 1000, 148 * 128 * 2 + 77])
def test_block_mlp_fused(engine, state_dict, M):
    """Fused out_proj+res+LN2+fc1+gelu+fc2+res kernel vs the same ops in fp32 torch (bf16 weights)."""
    from oracle import hyena_oracle as O

    layer = 2
    p = f"{O.BB}layers.{layer}."
    g = torch.Generator().manual_seed(M)
    y = _rand_bf16((M, 256), g)
    res = torch.randn(M, 256, generator=g)
    q = lambda w: w.to(torch.bfloat16).float()
    sd = state_dict
    r1 = y.float() @ q(sd[p + "mixer.out_proj.weight"]).T + sd[p + "mixer.out_proj.bias"] + res
    xn = F.layer_norm(r1, (256,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5).to(torch.bfloat16).float()
    h = F.gelu(xn @ q(sd[p + "mlp.fc1.weight"]).T + sd[p + "mlp.fc1.bias"], approximate="tanh").to(torch.bfloat16).float()
    ref = h @ q(sd[p + "mlp.fc2.weight"]).T + sd[p + "mlp.fc2.bias"] + r1
    out = engine.block_mlp(layer, y.cuda(), res.cuda().clone())
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    assert err <= 2e-2, (M, err)


@pytest.mark.parametrize("B,T", [(1, 70), (2, 300), (3, 1025), (2, 8193), (2, 272), (3, 140), (2, 8200), (2, 145)])
def test_block_in_fused(engine, state_dict, B, T):
    """Fused LN1+in_proj+short conv+gate vs fp32 torch (oracle ops), channel-major outputs.  Tails of 1..16 tokens (1025, 8193,
    272, 140, 8200) ride on the read's last full tile (160 token columns); 145 (17-token tail) and 70 keep a tile of their own."""
    from oracle import hyena_oracle as O

    layer = 1
    p = f"{O.BB}layers.{layer}."
    sd = state_dict
    g = torch.Generator().manual_seed(B * 10000 + T)
    res = torch.randn(B, T, 256, generator=g) * 1.5 + 0.3
    x = F.layer_norm(res, (256,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    u = F.linear(x, sd[p + "mixer.in_proj.weight"], sd[p + "mixer.in_proj.bias"]).transpose(1, 2)
    uc = F.conv1d(u, sd[p + "mixer.short_filter.weight"], sd[p + "mixer.short_filter.bias"], padding=2, groups=768)[..., :T]
    x0_ref, x1_ref, v_ref = uc.split(256, dim=1)
    vx_ref = v_ref * x1_ref
    vx, x0 = engine.block_in(layer, res.reshape(B * T, 256).cuda(), B, T)
    torch.cuda.synchronize()
    e_vx = (vx[..., :T].float().cpu() - vx_ref).abs().max().item()
    e_x0 = (x0[..., :T].float().cpu() - x0_ref).abs().max().item()
    # bf16 operands (K=256) + bf16 output rounding; |u| ~ 0.3, |vx| ~ 0.1
    assert e_x0 <= 2e-2 * max(1.0, x0_ref.abs().max().item()), (e_x0, x0_ref.abs().max().item())
    assert e_vx <= 2e-2 * max(1.0, vx_ref.abs().max().item()), (e_vx, vx_ref.abs().max().item())


@pytest.mark.parametrize("B,T", [(1, 128), (2, 300), (3, 1025), (4, 60), (7, 130), (5, 2112), (3, 1500)])
def test_block_mlp_channel_major_operand(engine, state_dict, B, T):
    """Same fused block tail, fed the conv output channel-major (MN-major UMMA A operand).  Reads ending in a partial tile of
    1..64 tokens share gathered tiles (300 -> 44-token tails, two per tile; 1025 -> 128 per tile; 60: no full tile at all;
    2112 -> 64-token tails); 1500 (92-token tails) and B = 1 keep one partial tile per read."""
    from oracle import hyena_oracle as O

    layer = 0
    p = f"{O.BB}layers.{layer}."
    Tp = (T + 63) // 64 * 64
    g = torch.Generator().manual_seed(B * 977 + T)
    y_cm = torch.zeros(B, 256, Tp, dtype=torch.bfloat16)
    y_cm[..., :T] = _rand_bf16((B, 256, T), g)
    y_cm[..., T:] = 3.0  # pad region must not matter
    res = torch.randn(B * T, 256, generator=g)
    q = lambda w: w.to(torch.bfloat16).float()
    sd = state_dict
    y = y_cm[..., :T].float().transpose(1, 2).reshape(B * T, 256)
    r1 = y @ q(sd[p + "mixer.out_proj.weight"]).T + sd[p + "mixer.out_proj.bias"] + res
    xn = F.layer_norm(r1, (256,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5).to(torch.bfloat16).float()
    h = F.gelu(xn @ q(sd[p + "mlp.fc1.weight"]).T + sd[p + "mlp.fc1.bias"], approximate="tanh").to(torch.bfloat16).float()
    ref = h @ q(sd[p + "mlp.fc2.weight"]).T + sd[p + "mlp.fc2.bias"] + r1
    out = engine.block_mlp_cm(layer, y_cm.cuda(), res.cuda(), T)
    torch.cuda.synchronize()
    err = (out.cpu() - ref).abs().max().item()
    assert err <= 2e-2, (B, T, err)
