"""Seeded random-init weights with the reference's state-dict key layout.

Pretrained checkpoints (`yangliz5/chimeralm`, chimeralm/models/lm.py:14) are not
reachable offline, so parity and benchmarks use random-init weights of the named
architecture.  Key names follow what `ClassificationLit(net=HyenaDna(...))`
produces (chimeralm/models/basic_module.py:39, components/hyena.py:237-238):
``net.backbone.backbone.*`` for the HF HyenaDNAModel and ``net.head.*`` for
BinarySequenceClassifier (components/hyena.py:50-74,161-166).

Init rules restate HF `_init_weights` for the backbone and the
PyTorch defaults for the head; values are drawn from a numpy Generator so the same
seed gives the same tensors on every box, independent of torch's RNG.
"""

from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch

from .config import DEFAULT_CONFIG, HyenaConfig

BACKBONE_PREFIX = "net.backbone.backbone."
HEAD_PREFIX = "net.head."


def _normal(rng, shape, std):
    return torch.from_numpy((rng.standard_normal(shape) * std).astype(np.float32))


def _uniform(rng, shape, bound):
    return torch.from_numpy(rng.uniform(-bound, bound, shape).astype(np.float32))


def _linear_default(rng, out_f, in_f):
    """nn.Linear default init: kaiming_uniform(a=sqrt(5)) == U(-1/sqrt(in), 1/sqrt(in))."""
    b = 1.0 / math.sqrt(in_f)
    return _uniform(rng, (out_f, in_f), b), _uniform(rng, (out_f,), b)


def positional_embedding(cfg: HyenaConfig):
    """HyenaPositionalEmbedding tables: z [1,Lmax,emb_dim], t [1,Lmax,1]."""
    L = cfg.max_seq_len
    t = torch.linspace(0, 1, L)[None, :, None]
    bands = (cfg.emb_dim - 1) // 2
    t_rescaled = torch.linspace(0, L - 1, L)[None, :, None]
    w = 2 * math.pi * t_rescaled / L
    f = torch.linspace(1e-4, bands - 1, bands)[None, None]
    z = torch.exp(-1j * f * w)
    z = torch.cat([t, z.real, z.imag], dim=-1)
    return z.float().contiguous(), t.float().contiguous()


def make_state_dict(seed: int = 0, cfg: HyenaConfig = DEFAULT_CONFIG,
                    head_logit_gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    """Build a full `ClassificationLit`-style state dict (fp32, CPU).

    `head_logit_gain` rescales `output_layer` so that random-init margins are not
    degenerate; 1.0 keeps the plain PyTorch default init.
    """
    rng = np.random.default_rng(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    D, std = cfg.d_model, cfg.initializer_range
    P = BACKBONE_PREFIX
    sd[P + "embeddings.word_embeddings.weight"] = _normal(rng, (cfg.vocab_rows, D), std)
    z, t = positional_embedding(cfg)
    resc = std / math.sqrt(2 * cfg.n_layer)
    for i in range(cfg.n_layer):
        L = f"{P}layers.{i}."
        sd[L + "mixer.in_proj.weight"] = _normal(rng, (cfg.inner_width, D), std)
        sd[L + "mixer.in_proj.bias"] = torch.zeros(cfg.inner_width)
        sd[L + "mixer.out_proj.weight"] = _normal(rng, (D, D), resc)
        sd[L + "mixer.out_proj.bias"] = torch.zeros(D)
        kb = 1.0 / math.sqrt(cfg.short_filter_order)  # Conv1d default, fan_in = 1*k
        sd[L + "mixer.short_filter.weight"] = _uniform(rng, (cfg.inner_width, 1, cfg.short_filter_order), kb)
        sd[L + "mixer.short_filter.bias"] = _uniform(rng, (cfg.inner_width,), kb)
        sd[L + "mixer.filter_fn.bias"] = _normal(rng, (D,), 1.0)
        sd[L + "mixer.filter_fn.pos_emb.z"] = z.clone()
        sd[L + "mixer.filter_fn.pos_emb.t"] = t.clone()
        freq = cfg.activation_freq * torch.ones(1, cfg.filter_order)
        dims = [cfg.emb_dim] + [cfg.filter_order] * (cfg.num_inner_mlps + 1)
        for j in range(cfg.num_inner_mlps + 1):
            sd[L + f"mixer.filter_fn.implicit_filter.{2 * j}.weight"] = _normal(rng, (dims[j + 1], dims[j]), std)
            sd[L + f"mixer.filter_fn.implicit_filter.{2 * j}.bias"] = torch.zeros(dims[j + 1])
            sd[L + f"mixer.filter_fn.implicit_filter.{2 * j + 1}.freq"] = freq.clone()
        last = 2 * (cfg.num_inner_mlps + 1)
        sd[L + f"mixer.filter_fn.implicit_filter.{last}.weight"] = _normal(rng, (D, cfg.filter_order), std)
        max_decay = math.log(cfg.target) / cfg.fast_decay_pct
        min_decay = math.log(cfg.target) / cfg.slow_decay_pct
        sd[L + "mixer.filter_fn.modulation.deltas"] = torch.linspace(min_decay, max_decay, D)[None, None].float()
        for n in ("norm1", "norm2"):
            sd[L + n + ".weight"] = torch.ones(D)
            sd[L + n + ".bias"] = torch.zeros(D)
        sd[L + "mlp.fc1.weight"] = _normal(rng, (cfg.d_inner, D), std)
        sd[L + "mlp.fc1.bias"] = torch.zeros(cfg.d_inner)
        sd[L + "mlp.fc2.weight"] = _normal(rng, (D, cfg.d_inner), resc)
        sd[L + "mlp.fc2.bias"] = torch.zeros(D)
    sd[P + "ln_f.weight"] = torch.ones(D)
    sd[P + "ln_f.bias"] = torch.zeros(D)

    H, Hh = cfg.head_hidden, cfg.head_hidden // 2
    Q = HEAD_PREFIX
    for name, (o, i) in (
        ("attention.0", (Hh, D)), ("attention.2", (1, Hh)),
        ("classifier.0", (H, D)), ("classifier.3", (H, H)),
        ("classifier.6.layers.0", (H, H)), ("classifier.6.layers.3", (H, H)),
        ("output_layer", (cfg.num_classes, H)),
    ):
        w, b = _linear_default(rng, o, i)
        sd[Q + name + ".weight"] = w
        sd[Q + name + ".bias"] = b
    if head_logit_gain != 1.0:
        sd[Q + "output_layer.weight"] *= head_logit_gain
        sd[Q + "output_layer.bias"] *= head_logit_gain
    return sd


def perturb_norms(sd, seed: int = 1, scale: float = 0.1):
    """Give LayerNorm affine params and zero-init biases non-trivial values.

    HF init leaves LN weight=1/bias=0 and Linear bias=0, which would hide indexing
    bugs in bias/affine handling; tests use this to make every parameter matter.
    """
    rng = np.random.default_rng(seed)
    out = OrderedDict()
    for k, v in sd.items():
        if k.endswith("pos_emb.z") or k.endswith("pos_emb.t") or k.endswith(".freq") or k.endswith("deltas"):
            out[k] = v.clone()
        elif (".norm" in k or "ln_f" in k) and k.endswith(".weight"):
            out[k] = v + _normal(rng, tuple(v.shape), scale)
        elif k.endswith(".bias") and float(v.abs().sum()) == 0.0:
            out[k] = _normal(rng, tuple(v.shape), scale * 0.2)
        else:
            out[k] = v.clone()
    return out


def save_lightning_ckpt(sd, path) -> None:
    """Write the checkpoint layout `--ckpt` expects: {"state_dict": ...} (scripts/model2hub.py:33)."""
    torch.save({"state_dict": dict(sd)}, str(path))


def load_checkpoint(path) -> "OrderedDict[str, torch.Tensor]":
    """Load a Lightning `.ckpt` ({"state_dict": ...}), a bare state dict `.pt`, or `.safetensors`."""
    path = str(path)
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file  # optional dependency

        raw = load_file(path)
    else:
        raw = torch.load(path, map_location="cpu", weights_only=True)
        if isinstance(raw, dict) and "state_dict" in raw:
            raw = raw["state_dict"]
    sd = OrderedDict()
    for k, v in raw.items():
        # Hub safetensors saved by PyTorchModelHubMixin carry the same `net.*` names.
        if not k.startswith("net."):
            k = "net." + k
        sd[k] = v.detach().float().contiguous()
    return sd
