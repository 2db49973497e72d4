"""End-to-end parity of the CUDA forward against the CPU oracle (-m gpu), stage by stage and on
final logits, on identical padded input_ids and identical seeded weights.

Tolerances: bf16 GEMM operands / bf16 inter-kernel activations with fp32 accumulation, fp32
residual stream, fp16-operand tensor-core FFT (fp32 accumulate) and fp32 pooling/head.  Residual
stream |err| <= 5e-2 (values ~ N(0,1) scaled by the block outputs), logits |err| <= 1e-3 (measured
2-3e-4; north_star: "stated bf16 tolerance", SURVEY.md 8(d) allows up to 2e-2), labels must agree
wherever |oracle margin| > 2 * 1e-3.  Label agreement proper is measured in test_gpu_labels.py."""

from pathlib import Path

import pytest
import torch

from chimeralm_b200.config import DEFAULT_CONFIG as CFG

pytestmark = pytest.mark.gpu

LOGIT_TOL = 1e-3
ROOT = Path(__file__).resolve().parents[1]


def _ids(B, T, seed, pad_left=0):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(7, 11, (B, T), generator=g)
    ids[:, -1] = 1
    for b in range(B):
        n = int(torch.randint(0, pad_left + 1, (1,), generator=g)) if pad_left else 0
        ids[b, :n] = 4
    return ids


def _resid(engine, B, T):
    """Residual stream from the device's blocked (R32) layout -> [B, T, D]."""
    from chimeralm_b200.engine import r32_to_rows

    M = B * T
    raw = engine.debug_copy("resid", ((M + 31) // 32 * 32 * CFG.d_model,), torch.float32)
    return r32_to_rows(raw, M).reshape(B, T, CFG.d_model).cpu()


def _oracle_states(sd, ids):
    """Residual stream after every block + final hidden + logits, from the oracle."""
    import torch.nn.functional as F
    from oracle import hyena_oracle as O

    with torch.inference_mode():
        h = F.embedding(ids, sd[O.BB + "embeddings.word_embeddings.weight"])
        states = [h.clone()]
        for i in range(CFG.n_layer):
            h = O.block(sd, i, h, CFG)
            states.append(h.clone())
        hf = F.layer_norm(h, (CFG.d_model,), sd[O.BB + "ln_f.weight"], sd[O.BB + "ln_f.bias"], CFG.layer_norm_epsilon)
        logits, attn = O.head(sd, hf, return_attention=True)
    return states, hf, logits, attn


@pytest.mark.parametrize("B,T", [(2, 301), (3, 1025)])
def test_stagewise_residual_stream(engine, state_dict, B, T):
    ids = _ids(B, T, seed=B * 1000 + T, pad_left=40)
    states, hf, logits_ref, _ = _oracle_states(state_dict, ids)
    engine.reserve(B, T)
    try:
        engine.set_debug_stop(0, 0)
        engine.forward(ids.cuda())
        got = _resid(engine, B, T)
        assert torch.equal(got, states[0]), "embedding gather must be exact"
        for l in range(CFG.n_layer):
            engine.set_debug_stop(l, 9)
            engine.forward(ids.cuda())
            got = _resid(engine, B, T)
            err = (got - states[l + 1]).abs().max().item()
            assert err <= 5e-2, (l, err)
    finally:
        engine.set_debug_stop(-1, -1)
    logits = engine.forward(ids.cuda()).cpu()
    assert (logits - logits_ref).abs().max().item() <= LOGIT_TOL


@pytest.mark.parametrize("B,T", [(1, 64), (4, 2049), (2, 8193), (3, 5000), (3, 3000)])
def test_logits_and_labels(engine, state_dict, B, T):
    from oracle import hyena_oracle as O

    ids = _ids(B, T, seed=T, pad_left=T // 3)
    ref = O.forward(state_dict, ids, CFG)
    logits, labels = engine.forward(ids.to(torch.uint8).cuda(), return_labels=True)
    logits, labels = logits.cpu(), labels.cpu()
    err = (logits - ref).abs().max().item()
    print(f"logits max|err| B={B} T={T} ({engine.longconv_variant(T)}): {err:.3e}")
    assert err <= LOGIT_TOL, err
    margin = ref[:, 1] - ref[:, 0]
    decided = margin.abs() > 2 * LOGIT_TOL
    assert torch.equal(labels[decided].long(), ref.argmax(1)[decided])
    # int64 ids (the reference's dtype) take the same path
    logits64 = engine.forward(ids.cuda()).cpu()
    assert torch.equal(logits64, logits)


def test_k1_bam_anchor_end_to_end(tmp_path, state_dict):
    """K1 (BASELINE.json configs[0]): `predict` on the reference's test BAM, batch 12, file order,
    Hub-flavour tokens, through the mirrored Trainer/BamDataModule/PredictionWriter stack.  The
    first batch (12 reads padded to 32 769 tokens: chunked FFT path, 11x left padding) is
    checked against the oracle on identical padded ids; every read must get a prediction line."""
    from pathlib import Path

    from chimeralm_b200.callbacks import PredictionWriter, load_predictions_from_folder
    from chimeralm_b200.data import BamDataModule, Trainer
    from chimeralm_b200.model import ClassificationLit
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model
    from oracle import hyena_oracle as O
    from oracle import tokenizer_oracle as TO
    from chimeralm_b200.bam import parse_bam_file

    bam = Path(__file__).parent / "golden" / "test_chimric_reads.bam"
    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    model = ClassificationLit(state_dict, device=0, max_batch=12, max_tokens=32769)
    dm = BamDataModule(tok, batch_size=12, predict_data_path=bam, engine=model.engine)
    out = tmp_path / "pred"
    trainer = Trainer(callbacks=[PredictionWriter(out, "batch")])
    res = trainer.predict(model, dataloaders=dm, return_predictions=True)
    preds = load_predictions_from_folder(out)
    recs = list(parse_bam_file(bam))
    assert len(preds) == 100 and set(preds) == {r["id"] for r in recs}
    assert sorted(p.name for p in out.glob("*.txt")) == sorted(f"0_{i}.txt" for i in range(9))
    # every batch vs the oracle on identical padded ids (token ids from the oracle tokenizer: bit-exact check of the GPU
    # encoder too); 6 of the 9 batches pad to 32 769 tokens (chunked FFT path, up to 60x left padding)
    dm.setup("predict")
    worst = 0.0
    for bi, batch in enumerate(dm.predict_dataloader()):
        chunk = recs[12 * bi: 12 * bi + 12]
        ids_ref = TO.collate([TO.encode(r["seq"], max_length=32769, add_cls=False) for r in chunk], padding_side="left")
        assert batch["input_ids"].cpu().tolist() == ids_ref, bi
        logits = model.forward(batch["input_ids"]).cpu()
        ref = O.forward(state_dict, torch.tensor(ids_ref), CFG)
        err = (logits - ref).abs().max().item()
        worst = max(worst, err)
        print(f"K1 batch {bi} ({len(ids_ref)} x {len(ids_ref[0])} tokens): logits max|err| {err:.3e}")
        assert err <= LOGIT_TOL, (bi, err)
        margin = ref[:, 1] - ref[:, 0]
        decided = margin.abs() > 2 * LOGIT_TOL
        got = torch.tensor([preds[r["id"]] for r in chunk])
        assert torch.equal(got[decided], ref.argmax(1)[decided]), bi
    assert bi == 8
    assert model.engine.tc_fallbacks == 0
    model.engine.close()


def test_mixed_lengths_bucketed(engine, state_dict):
    """K3-style ragged batch: reads of different lengths, left-padded to the batch maximum by the GPU
    encoder; logits vs oracle on the same padded ids."""
    from chimeralm_b200.engine import pack_reads
    from oracle import hyena_oracle as O
    from oracle import tokenizer_oracle as TO

    g = torch.Generator().manual_seed(11)
    lens = [1000, 1733, 2950, 400, 2999, 1, 0, 2048]
    seqs = ["".join("ACGT"[int(x)] for x in torch.randint(0, 4, (n,), generator=g)) for n in lens]
    ids_ref = TO.collate([TO.encode(s, max_length=32769, add_cls=False) for s in seqs], padding_side="left")
    T = len(ids_ref[0])
    bases, offs = pack_reads(seqs)
    ids, lens_dev = engine.encode(bases, offs, T, add_cls=False, add_sep=True, pad_left=True, max_bases=32768)
    assert ids.cpu().tolist() == ids_ref and lens_dev.cpu().tolist() == [n + 1 for n in lens]
    logits = engine.forward(ids).cpu()
    ref = O.forward(state_dict, torch.tensor(ids_ref), CFG)
    assert (logits - ref).abs().max().item() <= LOGIT_TOL


def test_repeated_forward_under_copy_traffic_is_stable_and_deterministic(state_dict):
    """Stress for the persistent kernels' mbarrier protocols (a lapped single-phase barrier in the tensor-core conv once
    showed up only as a rare launch failure in the BAM predict loop): many K2-shaped forwards back to back while another
    stream keeps the copy engines and HBM busy; every result must be bit-identical to the first."""
    from chimeralm_b200.engine import Engine

    B, T = 32, 8193
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    try:
        ids = _ids(B, T, seed=7).to(torch.uint8).cuda()
        assert eng.longconv_variant(T) == "fft_tensor_core"
        first = eng.forward(ids).clone()
        side = torch.cuda.Stream()
        host = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
        dev = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
        for i in range(150):
            if i % 3 == 0:
                with torch.cuda.stream(side):
                    dev.copy_(host, non_blocking=True)
                    host.copy_(dev, non_blocking=True)
            out = eng.forward(ids)
            if i % 25 == 24:
                assert torch.equal(out, first), i
        torch.cuda.synchronize()
        assert torch.equal(out, first)
    finally:
        eng.close()


def test_attention_weights_export(state_dict):
    """`save_attention=True` exposes softmax_t(scores) of the last forward, shape [B, T, 1] like the
    reference's `BinarySequenceClassifier.attention_weights` (oracle head pinned on the reference's own golden)."""
    from chimeralm_b200.model import ClassificationLit

    B, T = 3, 700
    ids = _ids(B, T, seed=11, pad_left=100)
    _, _, logits_ref, attn_ref = _oracle_states(state_dict, ids)
    model = ClassificationLit(state_dict, device=0, max_batch=B, max_tokens=T, save_attention=True)
    try:
        logits = model.forward(ids.cuda()).cpu()
        attn = model.attention_weights.cpu()
        assert attn.shape == (B, T, 1)
        assert (logits - logits_ref).abs().max().item() <= LOGIT_TOL
        ref = attn_ref.reshape(B, T)
        assert torch.allclose(attn[..., 0].sum(1), torch.ones(B), atol=1e-4)
        assert (attn[..., 0] - ref).abs().max().item() <= 2e-2 * ref.max().item()
        with pytest.raises(Exception):
            model.engine.attention_weights(B, T + 1)
    finally:
        model.engine.close()


def _synthetic_bam(path, n=96, seed=3):
    import numpy as np

    from chimeralm_b200.bam import BamWriter, make_record, minimal_header

    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    w = BamWriter(path, minimal_header())
    for i in range(n):
        w.write(make_record(f"read_{i:04d}", acgt[rng.integers(0, 4, int(rng.integers(200, 3000)))].tobytes(), sa_tag=(i % 7 != 0)))
    w.close()


@pytest.mark.parametrize("gpus", [1, 2])
def test_cli_predict_subprocess(tmp_path, gpus):
    """`python -m chimeralm_b200 predict` end to end in a fresh process; with 2 GPUs the per-rank workers are spawned
    (they must be importable from the children: the worker lives in predict_worker.py, not in __main__) and the union of the
    per-rank files equals the single-GPU result."""
    import subprocess
    import sys

    from chimeralm_b200.callbacks import load_predictions_from_folder

    if torch.cuda.device_count() < gpus:
        pytest.skip(f"needs {gpus} GPUs")
    bam = tmp_path / "in.bam"
    _synthetic_bam(bam)
    outs = {}
    for g in sorted({1, gpus}):
        out = tmp_path / f"pred{g}"
        r = subprocess.run([sys.executable, "-m", "chimeralm_b200", "predict", str(bam), "-o", str(out), "-b", "16", "--gpus", str(g)],
                           capture_output=True, text=True, timeout=600, cwd=str(ROOT))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[g] = load_predictions_from_folder(out)
    assert len(outs[1]) == 96 - 14   # every 7th record has no SA tag
    assert outs[gpus] == outs[1]


def test_full_size_batch_permutation_invariance(state_dict):
    """Size-independent property at the K2 size (32 reads x 8193 tokens, too big for the CPU oracle in a test): reads are
    independent, so permuting the batch permutes the logits.  Not bit-exact by construction - the tensor-core conv carries two
    reads in one complex transform, and a read's fp16 operand roundings depend on its partner - but far inside the logit
    tolerance; the same reads against the oracle are covered at B = 2 by test_logits_and_labels."""
    from chimeralm_b200.engine import Engine

    B, T = 32, 8193
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    try:
        ids = _ids(B, T, seed=123, pad_left=2000).to(torch.uint8).cuda()
        base = eng.forward(ids).clone()
        g = torch.Generator().manual_seed(5)
        for _ in range(2):
            perm = torch.randperm(B, generator=g).cuda()
            out = eng.forward(ids[perm].contiguous())
            assert (out - base[perm]).abs().max().item() <= 1e-3
        # an odd batch (the last read has no partner in the conv) and a batch of one
        assert (eng.forward(ids[:31].contiguous()) - base[:31]).abs().max().item() <= 1e-3
        assert (eng.forward(ids[7:8].contiguous()) - base[7:8]).abs().max().item() <= 1e-3
    finally:
        eng.close()


@pytest.mark.parametrize("option,T", [("fused_score_pool", 1500), ("fused_head", 1500), ("tc_conv", 8193), ("tc_chunked", 9000),
                                      ("tc_pipe", 8193), ("tc_pack4", 3000), ("tc_pipe_chunked", 16385), ("tc_pipe_chunked", 20000),
                                      ("mlp_helpers_high", 700), ("mlp_gather_tails", 8193), ("mlp_gather_tails", 700), ("mlp_gather_tails", 1281),
                                      ("mlp_gather_tails", 40),
                                      ("embed_in", 8193), ("embed_in", 700), ("embed_in", 40), ("embed_in", 3000), ("embed_in", 9000),
                                      ("in_ext_tail", 8193), ("in_ext_tail", 1296), ("in_ext_tail", 2060), ("in_ext_tail", 9000),
                                      ("pdl", 8193), ("pdl", 700), ("pdl", 20000),
                                      ("embed_res", 8193), ("embed_res", 700), ("embed_res", 40), ("embed_res", 1281),
                                      ("fused_mlp", 700), ("fused_in", 700), ("fast_conv", 700), ("mlp_epi16", 1500), ("mlp_pp", 8193), ("mlp_early_res", 8193), ("mlp_fc2_lag", 8193),
                                      ("skip_dead_res", 700), ("in_2cta", 700), ("in_2cta", 8193)])
def test_kernel_variants_agree(state_dict, option, T):
    """Every `clm_set_option` switch selects a different kernel for the same math (fused vs unfused, tensor-core vs fp32 FFT,
    8 vs 16 epilogue warps): flipping it must not move the logits by more than the parity tolerance."""
    from chimeralm_b200.engine import Engine

    from chimeralm_b200._lib import ChimeraLMNativeError

    B = 3
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    try:
        ids = _ids(B, T, seed=T, pad_left=T // 4).to(torch.uint8).cuda()
        base = eng.forward(ids).clone()
        default_on = option not in ("mlp_epi16", "mlp_pp", "in_2cta", "mlp_helpers_high")
        try:
            eng.set_option(option, {"mlp_fc2_lag": 2}.get(option, 0 if default_on else 1))
        except ChimeraLMNativeError as e:
            if "not compiled in" in str(e):
                pytest.skip(f"{option}: experiment kernel, product build (-DCLM_EXPERIMENTS adds it)")
            raise
        other = eng.forward(ids).clone()
        eng.set_option(option, {"mlp_early_res": 33, "mlp_fc2_lag": 1}.get(option, 1 if default_on else 0))
        again = eng.forward(ids)
        assert (other - base).abs().max().item() <= LOGIT_TOL, option
        assert torch.equal(again, base), option
    finally:
        eng.close()


def test_two_models_on_one_device_do_not_share_constants(state_dict):
    """The block-tail biases live in a __constant__ bank, which is per device, not per context: two models with different
    weights on the same GPU must each see their own (the bank is re-bound when the launching model changes)."""
    from chimeralm_b200.engine import Engine
    from chimeralm_b200.weights import make_state_dict

    sd_b = {k: torch.as_tensor(v) for k, v in make_state_dict(1).items()}
    B, T = 2, 600
    ids = _ids(B, T, seed=9).to(torch.uint8).cuda()
    a = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    ref_a = a.forward(ids).clone()
    b = Engine(sd_b, device=0, max_batch=B, max_tokens=T)   # finalizing b used to overwrite a's biases
    try:
        ref_b = b.forward(ids).clone()
        assert not torch.allclose(ref_a, ref_b)
        for _ in range(2):
            assert torch.equal(a.forward(ids), ref_a)
            assert torch.equal(b.forward(ids), ref_b)
    finally:
        a.close()
        b.close()
    c = Engine(sd_b, device=0, max_batch=B, max_tokens=T)
    try:
        assert torch.equal(c.forward(ids), ref_b)
    finally:
        c.close()


def test_k2_full_batch_vs_oracle(state_dict):
    """K2 at its full size - 32 reads x 8 193 tokens, the benchmark's step - against the CPU oracle on the same ids
    (no padding: K2 reads all have 8 192 bases), plus the same batch through the host entry point clm_predict_host."""
    from chimeralm_b200 import synth
    from chimeralm_b200.engine import Engine
    from oracle import hyena_oracle as O

    B, L = 32, 8192
    reads = synth.uniform_reads(B, L, synth.K2_SEED)
    ids = torch.from_numpy(synth.pad_left_ids(list(reads)))
    assert ids.shape == (B, L + 1)
    ref = O.forward(state_dict, ids, CFG)
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=L + 1)
    try:
        assert eng.longconv_variant(L + 1) == "fft_tensor_core"
        logits = eng.forward(ids.cuda(), check=True).cpu()
        err = (logits - ref).abs().max().item()
        print(f"K2 full batch 32 x 8193: logits max|err| {err:.3e}")
        assert err <= LOGIT_TOL, err
        bases = torch.from_numpy(reads.reshape(-1).copy()).pin_memory()
        offs = torch.arange(0, (B + 1) * L, L, dtype=torch.int64).pin_memory()
        lo, la = eng.predict_host(bases, offs, L + 1, add_cls=False, add_sep=True, pad_left=True, max_bases=32768)
        assert torch.equal(lo, logits)
        assert torch.equal(la.long(), (logits[:, 1] > logits[:, 0]).long())
        assert eng.tc_fallbacks == 0 and eng.native_tc_fallbacks == 0
    finally:
        eng.close()


def test_predict_step_contract(state_dict):
    """`predict_step` returns exactly `(logits, batch["labels"])` like the reference (basic_module.py:177-187); the
    device labels are an attribute, not a third element."""
    from chimeralm_b200.model import ClassificationLit

    B, T = 3, 400
    ids = _ids(B, T, seed=5, pad_left=50)
    model = ClassificationLit(state_dict, device=0, max_batch=B, max_tokens=T)
    try:
        batch = {"input_ids": ids.cuda(), "labels": torch.full((B,), -1, dtype=torch.int64), "id": torch.zeros(B, 256, dtype=torch.int8)}
        out = model.predict_step(batch, 0)
        assert isinstance(out, tuple) and len(out) == 2
        logits, labels = out
        assert logits.shape == (B, 2) and logits.dtype == torch.float32
        assert labels is batch["labels"]
        torch.cuda.synchronize()
        assert torch.equal(model.last_device_labels.cpu().long(), logits.argmax(1).cpu())
        assert torch.equal(model.forward(ids.cuda()), logits)
    finally:
        model.engine.close()


def test_token_id_out_of_range_is_reported(state_dict):
    """The reference's nn.Embedding raises IndexError for an id outside the table; here the forward's status word says
    so (embed_kernel flags it, the last kernel of the forward publishes it)."""
    from chimeralm_b200.engine import Engine

    B, T = 2, 300
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    try:
        ids = _ids(B, T, seed=1).to(torch.int32)
        good = eng.forward(ids.cuda(), check=True).clone()
        bad = ids.clone()
        bad[1, 17] = 16
        with pytest.raises(IndexError):
            eng.forward(bad.cuda(), check=True)
        neg = ids.clone()
        neg[0, 3] = -1
        eng.forward(neg.cuda())          # asynchronous form: the status is read after a synchronise
        seq = eng.last_seq
        torch.cuda.synchronize()
        with pytest.raises(IndexError):
            eng.forward_status(seq)
        # the word is cleared: the next clean forward is clean, and identical to the first
        assert torch.equal(eng.forward(ids.cuda(), check=True), good)
    finally:
        eng.close()


def test_failed_reserve_leaves_a_safe_context(state_dict):
    """a clm_reserve that runs out of memory must not leave dangling workspaces behind the old limits."""
    from chimeralm_b200._lib import ChimeraLMNativeError
    from chimeralm_b200.engine import Engine

    B, T = 2, 500
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    try:
        ids = _ids(B, T, seed=2).to(torch.uint8).cuda()
        good = eng.forward(ids, check=True).clone()
        with pytest.raises(ChimeraLMNativeError):
            eng.reserve(8192, 32769)     # ~1.7 TB of workspaces
        rc = eng.lib.clm_forward(eng.ctx, ids.data_ptr(), 2, B, T, good.data_ptr(), None, None)
        assert rc < 0, "a forward after a failed reserve must be refused, not run on freed memory"
        eng.reserve(B, T)
        assert torch.equal(eng.forward(ids, check=True), good)
    finally:
        eng.close()


def _trained_magnitude_weights(sd):
    """Weights of the magnitudes a trained checkpoint may have (random init keeps every activation tiny): in_proj x 6
    (v * x1 grows 36x, into the hundreds), per-layer filters scaled up / down by 50x, a large bias skip."""
    out = {k: v.clone() for k, v in sd.items()}
    fscale = [50.0, 0.02, 8.0, 0.3]
    for i in range(CFG.n_layer):
        p = f"net.backbone.backbone.layers.{i}.mixer."
        out[p + "in_proj.weight"] *= 6.0
        out[p + "in_proj.bias"] *= 6.0
        out[p + "filter_fn.implicit_filter.6.weight"] *= fscale[i]
        out[p + "filter_fn.bias"] *= 1.0 if i % 2 else 30.0
    return out


@pytest.mark.parametrize("B,T", [(2, 8193), (1, 20000), (1, 32769)])
def test_forward_with_trained_weight_magnitudes(state_dict, B, T):
    """ADVICE r1 / VERDICT weak #3: the fp16 tensor-core convolution with activations and filters far from the
    random-init magnitudes, full forward vs the oracle at 8 193, 20 000 and 32 769 tokens.  The per-channel input scale
    (calibration draw) and the per-segment spectrum scale keep every fp16 operand in range, so the tensor-core path must
    be as close to the oracle as the fp32-convolution path is."""
    from chimeralm_b200.engine import Engine
    from oracle import hyena_oracle as O

    sd = _trained_magnitude_weights(state_dict)
    ids = _ids(B, T, seed=T + 1, pad_left=T // 2)
    ref = O.forward(sd, ids, CFG)
    eng = Engine(sd, device=0, max_batch=B, max_tokens=T)
    try:
        assert eng.longconv_variant(T) == "fft_tensor_core"
        tc = eng.forward(ids.to(torch.uint8).cuda(), check=True).cpu()
        assert eng.tc_fallbacks == 0, "the batch left the fp16 range although the scales were calibrated"
        eng.set_option("tc_conv", 0)
        fp = eng.forward(ids.to(torch.uint8).cuda(), check=True).cpu()
        e_tc, e_fp = (tc - ref).abs().max().item(), (fp - ref).abs().max().item()
        vx = eng.debug_copy("vx", (B * 256 * ((T + 127) // 128 * 128),), torch.bfloat16).float().abs().max().item()
        print(f"trained-magnitude weights, {B} x {T}: |logits| {ref.abs().max():.3f}  max|v*x1| (last layer) {vx:.1f}  "
              f"err tensor-core conv {e_tc:.3e}, fp32 conv {e_fp:.3e}")
        assert torch.isfinite(tc).all()
        assert e_tc <= max(LOGIT_TOL, 2.0 * e_fp), (e_tc, e_fp)
    finally:
        eng.close()


def test_fp16_range_overflow_switches_to_fp32_conv(state_dict):
    """Automatic switch: with the input scale pushed 2^13 too high (test hook) the tensor-core convolution overflows,
    the forward's status says so, and both synchronous entry points redo the batch with the fp32 FFT kernel."""
    from chimeralm_b200._lib import Fp16RangeError
    from chimeralm_b200.engine import Engine

    B, T, L = 2, 8193, 8192
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    try:
        ids = _ids(B, T, seed=77).to(torch.uint8).cuda()
        eng.set_option("tc_conv", 0)
        want = eng.forward(ids, check=True).clone()
        eng.set_option("tc_conv", 1)
        eng.set_option("tc_scale_shift", 13)
        eng.forward(ids)
        seq = eng.last_seq
        torch.cuda.synchronize()
        with pytest.raises(Fp16RangeError):
            eng.forward_status(seq)
        got = eng.forward(ids, check=True)
        assert eng.tc_fallbacks == 1 and torch.equal(got, want)
        # clm_predict_host does the same on its own
        from chimeralm_b200 import synth

        reads = synth.uniform_reads(B, L, 9)
        bases = torch.from_numpy(reads.reshape(-1).copy()).pin_memory()
        offs = torch.arange(0, (B + 1) * L, L, dtype=torch.int64).pin_memory()
        kw = dict(add_cls=False, add_sep=True, pad_left=True, max_bases=32768)
        lo, _ = eng.predict_host(bases, offs, T, **kw)
        assert eng.native_tc_fallbacks == 1 and torch.isfinite(lo).all()
        eng.set_option("tc_scale_shift", 0)
        eng.set_option("tc_conv", 0)
        lo_fp, _ = eng.predict_host(bases, offs, T, **kw)
        assert torch.equal(lo, lo_fp)
        eng.set_option("tc_conv", 1)
        lo_tc, _ = eng.predict_host(bases, offs, T, **kw)
        assert eng.native_tc_fallbacks == 1 and (lo_tc - lo_fp).abs().max().item() <= LOGIT_TOL
    finally:
        eng.close()


def test_predict_host_submit_wait_pipeline(state_dict):
    """The two-halves form of clm_predict_host: three batches in flight, results identical to the one-at-a-time call,
    a fourth submit without a wait is refused, and a batch that overflows the fp16 convolution is
    redone in fp32 inside its own wait without disturbing the batches queued behind it."""
    from chimeralm_b200 import synth
    from chimeralm_b200._lib import ChimeraLMNativeError
    from chimeralm_b200.engine import Engine

    B, L = 4, 8192
    T = L + 1
    eng = Engine(state_dict, device=0, max_batch=B, max_tokens=T)
    kw = dict(add_cls=False, add_sep=True, pad_left=True, max_bases=32768)
    try:
        reads = [synth.uniform_reads(B, L, 100 + i) for i in range(5)]
        bases = [torch.from_numpy(r.reshape(-1).copy()).pin_memory() for r in reads]
        offs = torch.arange(0, (B + 1) * L, L, dtype=torch.int64).pin_memory()
        want = [eng.predict_host(b, offs, T, **kw)[0].clone() for b in bases]
        lo = [torch.empty(B, 2).pin_memory() for _ in range(5)]
        la = [torch.empty(B, dtype=torch.uint8).pin_memory() for _ in range(5)]
        tickets = [eng.predict_host_submit(bases[i], offs, T, logits_out=lo[i], labels_out=la[i], **kw) for i in range(3)]
        with pytest.raises(ChimeraLMNativeError):
            eng.predict_host_submit(bases[3], offs, T, logits_out=lo[3], labels_out=la[3], **kw)
        eng.predict_host_wait(tickets[0])
        tickets.append(eng.predict_host_submit(bases[3], offs, T, logits_out=lo[3], labels_out=la[3], **kw))
        for t in tickets[1:]:
            eng.predict_host_wait(t)
        with pytest.raises(ChimeraLMNativeError):
            eng.predict_host_wait(tickets[0])        # already waited for
        for i in range(4):
            assert torch.equal(lo[i], want[i]), i
            assert torch.equal(la[i].long(), (want[i][:, 1] > want[i][:, 0]).long())
        # overflow in the middle of the queue: batch 1 is submitted with the input scale pushed up (test hook)
        eng.set_option("tc_conv", 0)
        fp = [eng.predict_host(b, offs, T, **kw)[0].clone() for b in bases[:3]]
        eng.set_option("tc_conv", 1)
        t0 = eng.predict_host_submit(bases[0], offs, T, logits_out=lo[0], labels_out=la[0], **kw)
        torch.cuda.synchronize()
        eng.set_option("tc_scale_shift", 13)
        t1 = eng.predict_host_submit(bases[1], offs, T, logits_out=lo[1], labels_out=la[1], **kw)
        torch.cuda.synchronize()
        eng.set_option("tc_scale_shift", 0)
        t2 = eng.predict_host_submit(bases[2], offs, T, logits_out=lo[2], labels_out=la[2], **kw)
        for t in (t0, t1, t2):
            eng.predict_host_wait(t)
        assert eng.native_tc_fallbacks == 1
        assert torch.equal(lo[0], want[0]) and torch.equal(lo[2], want[2]) and torch.equal(lo[1], fp[1])
    finally:
        eng.close()


def test_bucketed_predict_flow_matches_per_read_forward(tmp_path, state_dict):
    """`--bucket` through the mirrored stack (BamDataModule -> Trainer -> PredictionWriter): length-sorted batches under a
    token budget, vectorised name rows, pinned staging ring.  Every read gets exactly one line, and a read's label equals
    the label of a forward on the identical padded batch."""
    from chimeralm_b200.callbacks import PredictionWriter, load_predictions_from_folder, resume_read_names
    from chimeralm_b200.data import BamDataModule, Trainer
    from chimeralm_b200.model import ClassificationLit
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    bam = tmp_path / "in.bam"
    _synthetic_bam(bam, n=120, seed=5)
    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    model = ClassificationLit(state_dict, device=0, max_batch=8, max_tokens=3001)
    try:
        dm = BamDataModule(tok, batch_size=8, predict_data_path=bam, engine=model.engine, bucket_by_length=True)
        out = tmp_path / "pred"
        res = Trainer(callbacks=[PredictionWriter(out, "batch")]).predict(model, dataloaders=dm, return_predictions=True)
        preds = load_predictions_from_folder(out)
        n_kept = len(dm.data_predict)
        assert len(preds) == n_kept == sum(len(i) for i, _ in res)
        assert sorted(int(j) for i, _ in res for j in i) == list(range(n_kept))
        lens = dm.data_predict.lengths
        budget = 8 * (int(lens.max()) + 1)
        for batch in dm.predict_dataloader():
            B_, T_ = batch["input_ids"].shape
            assert B_ * T_ <= budget and B_ <= 64
            logits = model.forward(batch["input_ids"]).cpu()
            names = resume_read_names(batch["id"])
            assert [preds[n] for n in names] == logits.argmax(1).tolist()
    finally:
        model.engine.close()


def test_single_sequence_predictor(state_dict):
    """`ChimeraLMPredictor.predict` (reference chimeralm/ui.py:36-79 without Gradio): validation messages, and the
    probabilities of one sequence against softmax(oracle logits) on the reference's tokenisation of it (max_length 32 768:
    the web path keeps 32 767 bases + [SEP])."""
    from chimeralm_b200.model import ClassificationLit
    from chimeralm_b200.predictor import ChimeraLMPredictor
    from oracle import hyena_oracle as O
    from oracle import tokenizer_oracle as TO

    model = ClassificationLit(state_dict, device=0, max_batch=1, max_tokens=32769)
    try:
        pr = ChimeraLMPredictor(model)
        assert pr.predict("  ")[0] == "Please enter a DNA sequence"
        assert pr.predict("ACGTX")[0].startswith("Invalid characters")
        g = torch.Generator().manual_seed(3)
        seq = "".join("ACGTN"[int(x)] for x in torch.randint(0, 5, (1500,), generator=g)).lower()
        name, conf, breakdown = pr.predict(seq)
        ids = torch.tensor([TO.encode(seq.upper(), max_length=32768, add_cls=False)])
        ref = torch.softmax(O.forward(state_dict, ids, CFG), dim=-1)[0]
        assert name == ["Biological", "Chimeric Artifact"][int(ref.argmax())]
        assert abs(conf - ref.max().item()) <= 1e-3
        assert breakdown == {"Biological": f"{ref[0].item():.3f}", "Chimeric Artifact": f"{ref[1].item():.3f}"}
    finally:
        model.engine.close()


@pytest.mark.parametrize("pooling", ["mean", "max", "cls"])
@pytest.mark.parametrize("fused", [True, False])
def test_other_pooling_modes(state_dict, pooling, fused):
    """`BinarySequenceClassifier.pooling_type` other than attention (components/hyena.py:97-115,134-136, mask None): mean over
    all positions, max over the sequence, the first position.  The oracle head is pinned on the reference class for each
    mode; the state dict carries no scorer weights then.  Both the fused tail and the unfused path."""
    import dataclasses

    from chimeralm_b200.engine import Engine
    from oracle import hyena_oracle as O

    cfg = dataclasses.replace(CFG, pooling_type=pooling)
    sd = {k: v for k, v in state_dict.items() if ".head.attention." not in k}
    B, T = 3, 1000
    ids = _ids(B, T, seed=13, pad_left=300)
    ref = O.forward(sd, ids, cfg)
    eng = Engine(sd, device=0, cfg=cfg, max_batch=B, max_tokens=T)
    try:
        if not fused:
            eng.set_option("fused_score_pool", 0)
            eng.set_option("fused_head", 0)
        logits = eng.forward(ids.to(torch.uint8).cuda(), check=True).cpu()
        err = (logits - ref).abs().max().item()
        print(f"pooling={pooling} fused={fused}: logits max|err| {err:.3e}")
        assert err <= LOGIT_TOL, err   # measured 1.3e-4 (mean), 3.6e-4 (max: ONE bf16-rounded position per channel), 2.2e-4 (cls)
    finally:
        eng.close()
    with pytest.raises(ValueError):
        Engine(sd, device=0, cfg=dataclasses.replace(CFG, pooling_type="median"), max_batch=1, max_tokens=64)
