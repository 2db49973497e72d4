"""Architecture constants of the `chimeralm predict` model.

The named architecture is fixed by the reference at
``chimeralm/models/lm.py:39-61`` (HyenaDna backbone "hyenadna-small-32k-seqlen"
+ BinarySequenceClassifier(256, 512, num_layers=2, attention pooling, gelu,
residual)).  The backbone's hyper-parameters are those of the HF Hub repo
``LongSafari/hyenadna-small-32k-seqlen-hf`` (``config.json``), which is NOT
vendored in the reference (call site ``chimeralm/models/components/hyena.py:237``);
they are restated here as fields so that every recalled value is configurable
(SURVEY.md Appendix A.1).
"""

from __future__ import annotations

from dataclasses import dataclass, asdict


@dataclass(frozen=True)
class HyenaConfig:
    d_model: int = 256
    n_layer: int = 4
    d_inner: int = 1024
    vocab_size: int = 12
    pad_vocab_size_multiple: int = 8
    max_seq_len: int = 32770
    hyena_order: int = 2
    filter_order: int = 64
    emb_dim: int = 5
    short_filter_order: int = 3
    num_inner_mlps: int = 2
    activation_freq: float = 10.0
    layer_norm_epsilon: float = 1e-5
    initializer_range: float = 0.02
    # HyenaExponentialModulation defaults
    fast_decay_pct: float = 0.3
    slow_decay_pct: float = 1.5
    target: float = 1e-2
    shift: float = 0.05
    # head (chimeralm/models/lm.py:46-55)
    head_hidden: int = 512
    head_num_layers: int = 2
    num_classes: int = 2
    # BinarySequenceClassifier.pooling_type (components/hyena.py:22): "attention" is what ChimeraLM uses (lm.py:46-55);
    # "mean", "max" and "cls" are the head's other modes (no scorer weights in the state dict then)
    pooling_type: str = "attention"

    @property
    def vocab_rows(self) -> int:
        v, m = self.vocab_size, self.pad_vocab_size_multiple
        return v if v % m == 0 else v + m - v % m

    @property
    def inner_width(self) -> int:
        return self.d_model * (self.hyena_order + 1)

    def to_dict(self) -> dict:
        return asdict(self)


DEFAULT_CONFIG = HyenaConfig()

# Token id layout, identical in the in-repo CharacterTokenizer
# (chimeralm/data/tokenizer.py:227-239) and the Hub HyenaDNA tokenizer.
CLS_ID, SEP_ID, BOS_ID, MASK_ID, PAD_ID, RESERVED_ID, UNK_ID = 0, 1, 2, 3, 4, 5, 6
BASE_IDS = {"A": 7, "C": 8, "G": 9, "T": 10, "N": 11}

# Hub tokenizer (what `chimeralm predict` loads, chimeralm/__main__.py:267):
# ids + [SEP], left padding, model_max_length 32770 -> max_len_single_sentence 32769.
HUB_MODEL_MAX_LENGTH = 32770
MAX_ID_LENGTH = 256  # chimeralm/data/tokenizer.py:94
