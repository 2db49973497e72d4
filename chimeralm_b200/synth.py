"""Synthetic predict workloads (no network: datasets and checkpoints are unavailable offline).

Definitions follow SURVEY.md 8(d) / BASELINE.json `configs`:
  K2  100 000 reads x 8 192 bases, iid uniform over ACGT, rng seed 20251018, batch 32
  K3  1 000 000 reads, length = clip(round(exp(N(ln 6000, 0.75^2))), 1000, 32768), seed 20251019,
      bases drawn from the same generator after the lengths; bucketed by length
  K4  K3 dealt to 2/4/8 ranks, token-balanced (see `deal_lpt`)
  K5  reads of exactly 32 768 bases, batch 64 per GPU
plus the label-agreement set (`label_reads`): reads of two composition classes, so that a head
whose output layer was fitted on a calibration draw (`fit_probe_head`) separates them with a
bimodal margin - the stand-in for a trained checkpoint when measuring label agreement.

Everything here is host-side numpy; nothing touches the GPU or the oracle.
"""

from __future__ import annotations

import math

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
K2_SEED, K3_SEED, K5_SEED = 20251018, 20251019, 20251020
K3_MU, K3_SIGMA, K3_MIN, K3_MAX = math.log(6000.0), 0.75, 1000, 32768


def uniform_reads(n: int, length: int, seed: int) -> np.ndarray:
    """[n, length] uint8 ASCII, iid uniform over ACGT (K2, K5)."""
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, size=(n, length))]


def k3_lengths(n: int, seed: int = K3_SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return np.clip(np.rint(np.exp(rng.normal(K3_MU, K3_SIGMA, size=n))), K3_MIN, K3_MAX).astype(np.int64)


def k3_reads(n: int, seed: int = K3_SEED):
    """(flat uint8 bases, int64 offsets[n+1]) of the first `n` reads of the K3 stream."""
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(np.exp(rng.normal(K3_MU, K3_SIGMA, size=n))), K3_MIN, K3_MAX).astype(np.int64)
    offsets = np.zeros(n + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    flat = ACGT[rng.integers(0, 4, size=int(offsets[-1]))]
    return flat, offsets


# ------------------------------------------------------------------------------------------
# Length bucketing and token-balanced dealing (replaces the reference's file-order batches and
# Lightning's rank::world sampler, chimeralm/data/bam.py:142-146,287-299)

def batch_cost(B: int, T: int) -> float:
    """Model of one batch's device time in arbitrary units: B * T * (c1 + c2 * log2 T).  The dense
    layers are linear in tokens (c1); the FFT convolution adds the log term.  The
    ratio c2 / c1 = 0.035 is fitted to the measured K2 / K5 steps (the 16 385- and 32 769-token conv
    costs per token are ~1.3x and ~1.6x the 8 193-token one while the GEMM kernels stay flat)."""
    return B * T * (1.0 + 0.035 * math.log2(max(T, 2)))


def bucket_batches(lengths: np.ndarray, batch: int, n_special: int = 1, max_tokens_per_batch: int | None = None):
    """Sort reads by length and cut consecutive runs into batches of <= `batch` reads (and, when
    given, <= max_tokens_per_batch padded tokens).  Returns a list of index arrays; every batch is
    padded to its longest member, so padding waste is what remains of the sort."""
    order = np.argsort(lengths, kind="stable")
    out, i, n = [], 0, len(order)
    while i < n:
        j = min(n, i + batch)
        if max_tokens_per_batch is not None:
            while j > i + 1 and (j - i) * (int(lengths[order[j - 1]]) + n_special) > max_tokens_per_batch:
                j -= 1
        if (j - i) > 1 and (j - i) % 2 == 1 and j < n:
            j -= 1   # even batches: the long convolution carries TWO reads per transform (real and imaginary part)
        out.append(order[i:j])
        i = j
    return out


def deal_lpt(costs, world: int):
    """Greedy longest-processing-time assignment: batches in decreasing cost order, each to the
    least-loaded rank.  Returns (per-rank lists of batch indices in decreasing cost, per-rank load)."""
    loads = np.zeros(world, np.float64)
    ranks = [[] for _ in range(world)]
    for b in np.argsort(-np.asarray(costs, np.float64), kind="stable"):
        r = int(np.argmin(loads))
        ranks[r].append(int(b))
        loads[r] += float(costs[b])
    return ranks, loads


# ------------------------------------------------------------------------------------------
# Label-agreement set

def label_reads(n: int, seed: int, len_lo: int = 1000, len_hi: int = 4000, fixed_len: int | None = None):
    """`n` reads of two composition classes: class 0 GC fraction ~ U(0.25, 0.40), class 1 ~ U(0.60, 0.75);
    lengths log-uniform in [len_lo, len_hi] (or all `fixed_len`).  Returns (list of uint8 arrays, classes)."""
    rng = np.random.default_rng(seed)
    cls = rng.integers(0, 2, size=n)
    if fixed_len is not None:
        lens = np.full(n, fixed_len, np.int64)
    else:
        lens = np.rint(np.exp(rng.uniform(math.log(len_lo), math.log(len_hi), size=n))).astype(np.int64)
    seqs = []
    for c, L in zip(cls, lens):
        gc = rng.uniform(0.60, 0.75) if c else rng.uniform(0.25, 0.40)
        p = np.array([(1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2])
        seqs.append(ACGT[rng.choice(4, size=int(L), p=p)])
    return seqs, cls


def pad_left_ids(seqs, pad_id: int = 4, sep_id: int = 1) -> np.ndarray:
    """Hub-flavour tokens of a batch (`ids + [SEP]`, left padding with [PAD]): uint8 [B, Tmax + 1].
    Same rule as chimeralm_b200.tokenizer / the oracle tokenizer for ACGT-only reads."""
    lut = np.full(256, 6, np.uint8)
    for ch, v in zip(b"ACGTN", (7, 8, 9, 10, 11)):
        lut[ch] = v
    T = max(len(s) for s in seqs) + 1
    ids = np.full((len(seqs), T), pad_id, np.uint8)
    for b, s in enumerate(seqs):
        ids[b, T - 1 - len(s): T - 1] = lut[s]
        ids[b, T - 1] = sep_id
    return ids


LABEL_CALIB_SEED, LABEL_EVAL_SEED = 101, 202


def label_calibration_batches(batch: int = 32):
    """The draw the probe head is fitted on (disjoint seeds from the evaluation set, same regimes: a trained model has
    seen padded batches of every length): 256 reads of 1-4 kb in length-bucketed batches, 256 in draw order (heavy
    left padding), 32 reads of 8 192 bases, one batch of 12 long reads padded to T = 32 769."""
    out = []
    seqs, cls = label_reads(256, LABEL_CALIB_SEED)
    lens = np.array([len(s) for s in seqs])
    out += [([seqs[i] for i in idx], cls[idx]) for idx in bucket_batches(lens, batch)]
    seqs, cls = label_reads(256, LABEL_CALIB_SEED + 1)
    out += [(seqs[i:i + batch], cls[i:i + batch]) for i in range(0, 256, batch)]
    seqs, cls = label_reads(32, LABEL_CALIB_SEED + 2, fixed_len=8192)
    out.append((seqs, cls))
    s1, c1 = label_reads(6, LABEL_CALIB_SEED + 3, fixed_len=32768)
    s2, c2 = label_reads(6, LABEL_CALIB_SEED + 4, 20000, 32768)
    out.append((s1 + s2, np.concatenate([c1, c2])))
    return out


def label_eval_batches(batch: int = 32):
    """The label-agreement evaluation set, as the batches `predict` would form (left-padded to the batch maximum):
      A  1 952 reads of 1-4 kb: 31 length-bucketed batches + 30 batches in draw order (heavy padding),
      B  64 reads of exactly 8 192 bases (T = 8 193, the K2 shape), 2 batches,
      C  one batch of 12 long reads padded to T = 32 769 (6 of 32 768 bases, 6 of 20-32 kb).
    Returns a list of (list of uint8 base arrays, int class array); 2 028 reads in 64 batches."""
    out = []
    seqs, cls = label_reads(1952, LABEL_EVAL_SEED)
    n_b = 31 * batch
    lens = np.array([len(s) for s in seqs[:n_b]])
    for idx in bucket_batches(lens, batch):
        out.append(([seqs[i] for i in idx], cls[idx]))
    for i in range(n_b, len(seqs), batch):
        out.append((seqs[i:i + batch], cls[i:i + batch]))
    seqs, cls = label_reads(64, LABEL_EVAL_SEED + 1, fixed_len=8192)
    for i in range(0, 64, batch):
        out.append((seqs[i:i + batch], cls[i:i + batch]))
    s1, c1 = label_reads(6, LABEL_EVAL_SEED + 2, fixed_len=32768)
    s2, c2 = label_reads(6, LABEL_EVAL_SEED + 3, 20000, 32768)
    out.append((s1 + s2, np.concatenate([c1, c2])))
    return out
