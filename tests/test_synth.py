"""Host-side workload logic (no GPU): synthetic K2/K3/K5 generators, length bucketing, token-balanced dealing
(replaces the reference's file-order batches + rank::world sampler, chimeralm/data/bam.py:142-146,287-299) and the
label-agreement golden fixture."""

from pathlib import Path

import numpy as np
import pytest

from chimeralm_b200 import synth

GOLDEN = Path(__file__).parent / "golden" / "label_agreement.npz"


def test_k3_length_law_and_determinism():
    lens = synth.k3_lengths(200_000)
    assert lens.min() >= 1000 and lens.max() <= 32768
    assert 7000 < lens.mean() < 8600            # log-normal(ln 6000, 0.75) clipped: mean ~ 8 kb
    assert (lens == 32768).mean() < 0.03 and (lens == 1000).mean() < 0.02
    flat, offs = synth.k3_reads(500)
    assert np.array_equal(np.diff(offs), lens[:500]) and len(flat) == offs[-1]
    assert set(np.unique(flat)) <= set(b"ACGT")
    flat2, offs2 = synth.k3_reads(500)
    assert np.array_equal(flat, flat2) and np.array_equal(offs, offs2)
    # the first n lengths do not depend on n (the bases are drawn after the lengths, so they do)
    assert np.array_equal(synth.k3_lengths(1000)[:500], lens[:500])


def test_bucketing_covers_every_read_once_and_bounds_padding():
    lens = synth.k3_lengths(20_000)
    batches = synth.bucket_batches(lens, 32)
    seen = np.concatenate(batches)
    assert len(seen) == len(lens) and len(np.unique(seen)) == len(lens)
    assert all(len(b) <= 32 for b in batches)
    padded = sum(len(b) * (int(lens[b].max()) + 1) for b in batches)
    real = int(lens.sum()) + len(lens)
    assert padded / real < 1.01                 # sorted runs of 32: under 1 % padding on 20 k reads
    order = np.arange(len(lens)).reshape(-1, 32)
    padded_fo = sum(32 * (int(lens[b].max()) + 1) for b in order)
    assert padded_fo / real > 2.0               # file order (the reference's policy) pads more than 2x
    capped = synth.bucket_batches(lens, 64, max_tokens_per_batch=32 * 8193)
    assert all(len(b) * (int(lens[b].max()) + 1) <= 32 * 8193 or len(b) == 1 for b in capped)
    assert sum(len(b) for b in capped) == len(lens)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_lpt_dealing_balances_tokens(world):
    lens = synth.k3_lengths(20_000)
    batches = synth.bucket_batches(lens, 32)
    costs = [synth.batch_cost(len(b), int(lens[b].max()) + 1) for b in batches]
    ranks, loads = synth.deal_lpt(costs, world)
    assert sorted(sum(ranks, [])) == list(range(len(batches)))
    assert loads.max() / loads.mean() < 1.01    # LPT on ~600 batches: within 1 % of perfect
    # the reference's sampler deals reads round-robin; on length-sorted batches dealt round-robin the last ranks always
    # get the longer batch of each round
    rr = np.zeros(world)
    for i, c in enumerate(costs):
        rr[i % world] += c
    assert loads.max() <= rr.max() + 1e-9


def test_label_sets_are_deterministic_and_disjoint():
    ev = synth.label_eval_batches()
    assert sum(len(s) for s, _ in ev) == 2028 and len(ev) == 64
    assert max(max(len(x) for x in s) for s, _ in ev) == 32768
    ev2 = synth.label_eval_batches()
    assert all(np.array_equal(a, b) for (s1, _), (s2, _) in zip(ev, ev2) for a, b in zip(s1, s2))
    cal = synth.label_calibration_batches()
    assert sum(len(s) for s, _ in cal) == 256 + 256 + 32 + 12
    ids = synth.pad_left_ids(ev[-1][0])
    assert ids.shape == (12, 32769) and (ids[:, -1] == 1).all() and set(np.unique(ids)) <= {1, 4, 7, 8, 9, 10}


def test_label_golden_fixture_is_consistent_with_the_generator():
    g = np.load(GOLDEN)
    n = int(g["n_reads"])
    assert n == 2028 == len(g["classes"]) == len(g["logits_probe"]) == len(g["logits_centred"])
    classes = np.concatenate([c for _, c in synth.label_eval_batches()])
    assert np.array_equal(classes, g["classes"])
    m = g["logits_probe"][:, 1] - g["logits_probe"][:, 0]
    assert 0.3 < (m > 0).mean() < 0.7           # both labels occur
    assert g["probe_w"].shape == (2, 512) and g["probe_b"].shape == (2,) and g["centred_b"].shape == (2,)
    mc = g["logits_centred"][:, 1] - g["logits_centred"][:, 0]
    assert 0.2 < (mc > 0).mean() < 0.8
