"""Host-side mirror of the reference tokenizer interface.

Mirrors `chimeralm/data/tokenizer.py` (reference): `CharacterTokenizer` (:190-327),
`load_tokenizer_from_hyena_model` (:36-55), `tokenize_and_align_labels_and_quals_ids`
(:85-114) and `DataCollator.torch_call` (:136-187) — same names, argument meaning and
outputs — without depending on `transformers`.  Single strings are encoded with a numpy
lookup table; whole batches go through the CUDA encoder (`Engine.encode` /
`clm_encode_batch`), which is the path `predict` uses.
"""

from __future__ import annotations

import numpy as np
import torch

from .config import BASE_IDS, CLS_ID, HUB_MODEL_MAX_LENGTH, MAX_ID_LENGTH, PAD_ID, SEP_ID, UNK_ID

_VOCAB = {"[CLS]": 0, "[SEP]": 1, "[BOS]": 2, "[MASK]": 3, "[PAD]": 4, "[RESERVED]": 5, "[UNK]": 6, **BASE_IDS}
_INV_VOCAB = {v: k for k, v in _VOCAB.items()}
_LUT = np.full(256, UNK_ID, dtype=np.uint8)
for _ch, _i in BASE_IDS.items():
    _LUT[ord(_ch)] = _i

MODEL_SEQ_INPUT = "input_ids"
MODEL_LABEL_INPUT = "labels"
SEQ_FEATURE, ID_FEATURE = "seq", "id"


class CharacterTokenizer:
    """Character tokenizer with the reference's id layout.

    `add_cls=True, padding_side="right"` is the in-repo flavour (reference :198,297-306);
    `add_cls=False, padding_side="left"` is the Hub HyenaDNA flavour `predict` loads.
    """

    model_input_names = [MODEL_SEQ_INPUT]

    def __init__(self, model_max_length: int | None = None, padding_side: str = "right", *, add_cls: bool = True,
                 add_sep: bool = True, emit_attention_mask: bool = False):
        if padding_side not in ("left", "right"):
            raise ValueError(f"padding_side must be 'left' or 'right', got {padding_side!r}")
        self.characters = ("A", "C", "G", "T", "N")
        self.model_max_length = model_max_length if model_max_length is not None else int(1e30)
        self.padding_side = padding_side
        self.add_cls, self.add_sep = add_cls, add_sep
        self.emit_attention_mask = emit_attention_mask
        self.pad_token_id, self.cls_token_id, self.sep_token_id, self.unk_token_id = PAD_ID, CLS_ID, SEP_ID, UNK_ID
        self.all_special_tokens = ["[BOS]", "[SEP]", "[UNK]", "[CLS]", "[PAD]", "[MASK]"]

    # ---- properties the reference code reads
    @property
    def vocab_size(self) -> int:
        return len(_VOCAB)

    def get_vocab(self) -> dict:
        return dict(_VOCAB)

    @property
    def num_special_tokens(self) -> int:
        return int(self.add_cls) + int(self.add_sep)

    @property
    def max_len_single_sentence(self) -> int:
        return self.model_max_length - self.num_special_tokens

    # ---- encoding
    def encode_array(self, text: str | bytes, max_length: int | None = None) -> np.ndarray:
        raw = text.encode("utf-8", "replace") if isinstance(text, str) else bytes(text)
        if isinstance(text, str) and len(raw) != len(text):
            # non-ASCII characters are single tokens ([UNK]) in the reference, not several bytes
            raw = bytes(ord(c) if ord(c) < 128 else 0 for c in text)
        ids = _LUT[np.frombuffer(raw, dtype=np.uint8)]
        ns = self.num_special_tokens
        if max_length is not None and ids.size + ns > max_length:
            ids = ids[: max(max_length - ns, 0)]
        parts = ([np.array([CLS_ID], np.uint8)] if self.add_cls else []) + [ids] + (
            [np.array([SEP_ID], np.uint8)] if self.add_sep else [])
        return np.concatenate(parts) if parts else ids

    def encode(self, text: str, *, truncation: bool = False, max_length: int | None = None, **_) -> list[int]:
        if truncation and max_length is None and self.model_max_length < int(1e30):
            max_length = self.model_max_length
        return self.encode_array(text, max_length if truncation else None).tolist()

    def __call__(self, text, truncation: bool = False, max_length: int | None = None, padding=False, **_) -> dict:
        if isinstance(text, (list, tuple)):
            rows = [self.encode(t, truncation=truncation, max_length=max_length) for t in text]
            if padding:
                rows = self._pad_rows(rows)
            out = {MODEL_SEQ_INPUT: rows}
        else:
            out = {MODEL_SEQ_INPUT: self.encode(text, truncation=truncation, max_length=max_length)}
        if self.emit_attention_mask:
            v = out[MODEL_SEQ_INPUT]
            out["attention_mask"] = [[int(i != PAD_ID) for i in r] for r in v] if v and isinstance(v[0], list) else [1] * len(v)
        return out

    def _pad_rows(self, rows, length: int | None = None):
        tmax = max(len(r) for r in rows) if length is None else length
        if self.padding_side == "left":
            return [[PAD_ID] * (tmax - len(r)) + list(r) for r in rows]
        return [list(r) + [PAD_ID] * (tmax - len(r)) for r in rows]

    def pad(self, features, padding=True, max_length=None, pad_to_multiple_of=None, return_tensors=None):
        """`tokenizer.pad` as DataCollatorWithPadding calls it: pad to the longest in the batch."""
        rows = [f[MODEL_SEQ_INPUT] for f in features]
        rows = [r.tolist() if hasattr(r, "tolist") else list(r) for r in rows]
        tmax = max(len(r) for r in rows)
        if pad_to_multiple_of:
            tmax = (tmax + pad_to_multiple_of - 1) // pad_to_multiple_of * pad_to_multiple_of
        padded = self._pad_rows(rows, tmax)
        batch = {MODEL_SEQ_INPUT: padded}
        if self.emit_attention_mask or "attention_mask" in features[0]:
            n = [len(r) for r in rows]
            batch["attention_mask"] = ([[0] * (tmax - k) + [1] * k for k in n] if self.padding_side == "left"
                                       else [[1] * k + [0] * (tmax - k) for k in n])
        if return_tensors == "pt":
            batch = {k: torch.tensor(v, dtype=torch.int64) for k, v in batch.items()}
        return batch

    # ---- decoding
    def convert_ids_to_tokens(self, ids):
        return [_INV_VOCAB[int(i)] for i in ids]

    def decode(self, token_ids, *, skip_special_tokens: bool = True, **_) -> str:
        if isinstance(token_ids, dict):
            token_ids = token_ids[MODEL_SEQ_INPUT]
        if isinstance(token_ids, torch.Tensor):
            token_ids = token_ids.tolist()
        if token_ids and isinstance(token_ids[0], list):
            token_ids = token_ids[0]
        toks = self.convert_ids_to_tokens(token_ids)
        if skip_special_tokens:
            toks = [t for t in toks if t not in self.all_special_tokens]
        return "".join(toks)


def load_tokenizer_from_hyena_model(model_name: str) -> CharacterTokenizer:
    """Offline equivalent of reference `load_tokenizer_from_hyena_model` (:36-55): the Hub
    HyenaDNA tokenizer = same vocabulary, `ids + [SEP]`, left padding."""
    max_lengths = {
        "hyenadna-tiny-1k-seqlen": 1024,
        "hyenadna-small-32k-seqlen": 32768,
        "hyenadna-medium-160k-seqlen": 160000,
        "hyenadna-medium-450k-seqlen": 450000,
        "hyenadna-large-1m-seqlen": 1_000_000,
    }
    if model_name not in max_lengths:
        raise ValueError(f"Model name {model_name} not found in available models.")
    return CharacterTokenizer(model_max_length=max_lengths[model_name] + 2, padding_side="left", add_cls=False,
                              emit_attention_mask=True)


assert load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen").model_max_length == HUB_MODEL_MAX_LENGTH


def encode_read_name(rid: str, max_id_length: int = MAX_ID_LENGTH) -> list[int]:
    """Reference :108-111: [len] + ord(chars), cut / zero-padded to 256."""
    new_id = [len(rid)] + [ord(c) for c in rid]
    return new_id[:max_id_length] if len(new_id) > max_id_length else new_id + [0] * (max_id_length - len(new_id))


def encode_read_name_rows(name_rows: np.ndarray, max_id_length: int = MAX_ID_LENGTH) -> np.ndarray:
    """The same rows for a whole batch, without a Python loop per read: `name_rows` is uint8 [n, stride] of
    NUL-terminated names (what the native BAM reader fills).  Returns int8 [n, max_id_length] exactly equal to
    `np.array([encode_read_name(name) ...]).astype(np.int8)` (the collator's cast, reference :167-168)."""
    n, stride = name_rows.shape
    is_nul = name_rows == 0
    lens = np.where(is_nul.any(axis=1), is_nul.argmax(axis=1), stride)
    out = np.zeros((n, max_id_length), np.uint8)
    w = min(stride, max_id_length - 1)
    out[:, 1 : 1 + w] = np.where(np.arange(w)[None, :] < lens[:, None], name_rows[:, :w], 0)
    out[:, 0] = lens.astype(np.uint8)          # len <= 254 for BAM names; int64 -> int8 wraps the same way
    return out.view(np.int8)


def names_to_rows(names, stride: int = MAX_ID_LENGTH) -> np.ndarray:
    """list[str] -> uint8 [n, stride] NUL-padded (names longer than stride - 1 are cut like the reference's slice)."""
    if not len(names):
        return np.zeros((0, stride), np.uint8)
    fixed = np.array([x.encode("ascii", "replace") for x in names], dtype=f"S{stride - 1}")
    rows = np.zeros((len(names), stride), np.uint8)
    rows[:, : stride - 1] = fixed.view(np.uint8).reshape(len(names), stride - 1)
    return rows


def tokenize_and_align_labels_and_quals_ids(data, tokenizer, max_length, *, include_qual=False, seq_feature=SEQ_FEATURE,
                                            id_feature=ID_FEATURE, max_id_length=MAX_ID_LENGTH, **_):
    """Per-example tokenise fn of the predict dataset (reference :85-114)."""
    if include_qual:
        raise NotImplementedError("quality scores are a train-only branch of the reference (out of scope)")
    out = tokenizer(data[seq_feature], truncation=True, max_length=max_length, padding=True)
    out.update({"id": encode_read_name(data[id_feature], max_id_length), MODEL_LABEL_INPUT: -1})
    return out


class DataCollator:
    """`DataCollator.torch_call` (reference :136-187) for the predict dataset."""

    def __init__(self, tokenizer, padding=True, max_length=None, pad_to_multiple_of=None):
        self.tokenizer, self.padding, self.max_length, self.pad_to_multiple_of = tokenizer, padding, max_length, pad_to_multiple_of

    def torch_call(self, features):
        label_name = "label" if "label" in features[0] else "labels"
        labels = [f[label_name] for f in features] if label_name in features[0] else None
        stripped = [{k: v for k, v in f.items() if k not in ("input_quals", label_name, "id")} for f in features]
        batch = self.tokenizer.pad(stripped, padding=self.padding, max_length=self.max_length,
                                   pad_to_multiple_of=self.pad_to_multiple_of, return_tensors="pt")
        if "id" in features[0]:
            ids = [f["id"].tolist() if hasattr(f["id"], "tolist") else list(f["id"]) for f in features]
            batch["id"] = torch.tensor(ids, dtype=torch.int8)  # names >= 128 chars overflow, as in the reference
        if labels is not None:
            batch[label_name] = torch.tensor(labels, dtype=torch.int64)
        return batch

    __call__ = torch_call
