"""Host-side model of the tile schedules that fold a read's tail tokens into shared tiles.

The index arithmetic below restates what the kernels and their launchers compute (chimeralm_b200/csrc/block_mlp.cuh
`tile_row` + `gather_tails_kernel`, csrc/api.cu `launch_block_mlp` / `launch_block_in`, csrc/block_in.cuh `ext_L`) and checks
the properties the GPU parity tests rely on for EVERY shape, not just the ones they run: each token row of the batch is owned by
exactly one (tile, row) slot, gathered columns address the right token of the right read, no slot points outside the batch, and
the tile count is what the wave arithmetic in the docs says (K2: 2 080 -> 2 049 / 2 048 tiles)."""
import itertools

import pytest

BM = 128


def block_mlp_plan(B, T, gather=True):
    """launch_block_mlp: (tiles_per_seq, n_full_tiles, num_tiles, L, P)."""
    tps = (T + BM - 1) // BM
    plan = dict(tps=tps, n_full=B * tps, num=B * tps, L=0, P=0)
    if gather and B > 1:
        L = T % BM
        P = BM // L if L > 0 else 0
        n_g = (B + P - 1) // P if P >= 2 else 0
        if 0 < n_g < B:
            plan.update(tps=T // BM, n_full=B * (T // BM), num=B * (T // BM) + n_g, L=L, P=P)
    return plan


def block_mlp_row(plan, B, T, tile, r):
    """block_mlp_kernel::tile_row -> (row, ok)."""
    if tile < plan["n_full"]:
        t = (tile % plan["tps"]) * BM + r
        return (tile // plan["tps"]) * T + t, t < T
    jj = r // plan["L"]
    j = (tile - plan["n_full"]) * plan["P"] + jj
    return j * T + (T - plan["L"]) + (r - jj * plan["L"]), jj < plan["P"] and j < B


def gathered_column_source(plan, B, T, col):
    """gather_tails_kernel: column of yg -> (read, token) or None (zero column)."""
    g, r = col >> 7, col & 127
    jj = r // plan["L"]
    j = g * plan["P"] + jj
    if jj < plan["P"] and j < B:
        return j, (T - plan["L"]) + (r - jj * plan["L"])
    return None


SHAPES = [(32, 8193), (64, 32769), (2, 300), (3, 1025), (4, 60), (7, 130), (5, 2112), (3, 1500), (1, 8193), (255, 1025),
          (40, 3000), (9, 2064), (70, 200), (2, 128), (3, 129), (129, 129), (5, 64), (6, 65)]


@pytest.mark.parametrize("B,T", SHAPES)
def test_block_mlp_tiles_cover_every_row_once(B, T):
    plan = block_mlp_plan(B, T)
    seen = {}
    for tile, r in itertools.product(range(plan["num"]), range(BM)):
        row, ok = block_mlp_row(plan, B, T, tile, r)
        if not ok:
            continue
        assert 0 <= row < B * T
        assert row not in seen, (tile, r, seen[row])
        seen[row] = (tile, r)
        if tile >= plan["n_full"]:   # the operand column the producer loads for this row is that token of that read
            src = gathered_column_source(plan, B, T, (tile - plan["n_full"]) * BM + r)
            assert src is not None and src[0] * T + src[1] == row
    assert len(seen) == B * T
    if plan["L"]:
        # never more tiles than one partial tile per read, and the gathered operand fits the buffer clm_reserve sizes
        assert plan["num"] < B * ((T + BM - 1) // BM)
        assert (plan["num"] - plan["n_full"]) <= (B + 1) // 2


def test_gathering_is_off_where_it_cannot_pay():
    assert block_mlp_plan(1, 8193)["L"] == 0        # one read: one partial tile either way
    assert block_mlp_plan(3, 1500)["L"] == 0        # 92-token tails: one per tile
    assert block_mlp_plan(8, 1024)["L"] == 0        # no tail at all
    assert block_mlp_plan(32, 8193, gather=False)["num"] == 2080


def test_k2_and_k5_tile_counts():
    k2 = block_mlp_plan(32, 8193)
    assert (k2["num"], k2["L"], k2["P"]) == (2049, 1, 128)          # 13.84 waves on 148 SMs instead of 14.05
    k5 = block_mlp_plan(64, 32769)
    assert k5["num"] == 64 * 256 + 1


def block_in_plan(B, T, ext=True):
    """launch_block_in: (tiles_per_seq, num_tiles, ext_L)."""
    L = T % BM
    if ext and T >= BM and 1 <= L <= 16:
        return T // BM, B * (T // BM), L
    tps = (T + BM - 1) // BM
    return tps, B * tps, 0


@pytest.mark.parametrize("B,T", SHAPES + [(2, 272), (3, 140), (2, 8200), (2, 145), (1, 70)])
def test_block_in_tiles_cover_every_token_once(B, T):
    tps, num, ext_L = block_in_plan(B, T)
    seen = set()
    for tile in range(num):
        b, t0 = tile // tps, (tile % tps) * BM
        cols = BM + (16 if ext_L and tile % tps == tps - 1 else 0)   # the read's last tile computes 16 more columns
        for j in range(cols):
            t = t0 + j
            valid = t < T if j < BM else (j - BM) < ext_L            # main columns: TMA bound; extra columns: the ext_L mask
            if valid:
                assert (b, t) not in seen
                seen.add((b, t))
        if cols > BM:
            # the extra store round writes tokens [t0 + 128, t0 + 256): exactly the read's last 128-token row of [.., Tp)
            Tp = (T + BM - 1) // BM * BM
            assert t0 + BM == Tp - BM and t0 + 2 * BM == Tp
    assert len(seen) == B * T
    if ext_L:
        assert num == B * (T // BM)


def test_k2_block_in_tile_count():
    assert block_in_plan(32, 8193) == (64, 2048, 1)
    assert block_in_plan(32, 8193, ext=False) == (65, 2080, 0)
    assert block_in_plan(2, 145)[2] == 0             # 17-token tail: a tile of its own


def test_block0_table_lookup_equals_the_oracle_ops():
    """DESIGN.md 4.3a in the oracle's own arithmetic (fp32, CPU): block 0's LayerNorm1 + in_proj depend on the token id alone, so
    a 16-row table U[id] followed by the causal 3-tap filter and the gate gives exactly what the oracle's embedding ->
    layer_norm -> linear -> conv1d -> split -> gate gives (chimeralm_b200/csrc/embed_in.cuh restated with torch ops; the
    identity, not the CUDA kernel, is what this pins - the kernel is held to the oracle by tests/test_gpu_forward.py)."""
    import torch
    import torch.nn.functional as F

    from chimeralm_b200.config import DEFAULT_CONFIG as cfg
    from chimeralm_b200.weights import make_state_dict, perturb_norms
    from oracle import hyena_oracle as O

    sd = {k: torch.as_tensor(v) for k, v in make_state_dict(3).items()}
    perturb_norms(sd, seed=4)   # non-trivial LayerNorm affines
    p = f"{O.BB}layers.0."
    D = cfg.d_model
    E = sd[O.BB + "embeddings.word_embeddings.weight"]
    g = torch.Generator().manual_seed(11)
    ids = torch.randint(0, 16, (3, 301), generator=g)
    ids[0, :40] = 4   # a [PAD] prefix like a left-padded batch
    # the oracle's path (hyena_oracle.block / hyena_operator up to the first gate)
    x = F.layer_norm(F.embedding(ids, E).float(), (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.layer_norm_epsilon)
    u = F.linear(x, sd[p + "mixer.in_proj.weight"], sd[p + "mixer.in_proj.bias"]).transpose(1, 2)
    uc = F.conv1d(u, sd[p + "mixer.short_filter.weight"], sd[p + "mixer.short_filter.bias"], padding=2, groups=3 * D)[..., :ids.shape[1]]
    x0_ref, x1_ref, v_ref = uc.split(D, dim=1)
    # the table form: U[id][ch], u = 0 before the read, out = w0 u[t-2] + w1 u[t-1] + w2 u[t] + cb
    U = F.linear(F.layer_norm(E[:16].float(), (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], cfg.layer_norm_epsilon),
                 sd[p + "mixer.in_proj.weight"], sd[p + "mixer.in_proj.bias"])          # [16, 768]
    w = sd[p + "mixer.short_filter.weight"].reshape(3 * D, 3)
    cb = sd[p + "mixer.short_filter.bias"]
    ut = U[ids]                                                                           # [B, T, 768]
    z = torch.zeros_like(ut[:, :1])
    um1 = torch.cat([z, ut[:, :-1]], 1)
    um2 = torch.cat([z, z, ut[:, :-2]], 1)
    out = (w[:, 0] * um2 + w[:, 1] * um1 + w[:, 2] * ut + cb).transpose(1, 2)
    x0, x1, v = out.split(D, dim=1)
    assert (x0 - x0_ref).abs().max().item() <= 1e-5
    assert (v * x1 - v_ref * x1_ref).abs().max().item() <= 1e-5
