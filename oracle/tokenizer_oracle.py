"""ORACLE (test infrastructure, never shipped): plain-Python restatement of the
reference's character tokenisation, collation and prediction-file rules.

Pinned against the reference itself: `oracle/make_golden.py` imports
`/root/reference/chimeralm/data/tokenizer.py` by file path and records its outputs in
`tests/golden/tokenizer_golden.json`; `tests/test_oracle.py` checks every function here
against them and against the reference's own KATs
(`tests/test_tokenzier.py:3-17`: "ATCG" -> [0,7,10,8,9,1];
`tests/test_data_module.py:56-73`: left-padded (12, 98) batch).
"""

from __future__ import annotations

VOCAB = {"[CLS]": 0, "[SEP]": 1, "[BOS]": 2, "[MASK]": 3, "[PAD]": 4, "[RESERVED]": 5, "[UNK]": 6,
         "A": 7, "C": 8, "G": 9, "T": 10, "N": 11}  # chimeralm/data/tokenizer.py:227-239


def encode(seq: str, *, max_length: int | None, add_cls: bool, add_sep: bool = True) -> list[int]:
    """`tokenizer(seq, truncation=True, max_length=max_length)["input_ids"]`.

    chimeralm/data/tokenizer.py:264-268 (`list(text)`, dict lookup, unknown -> [UNK]=6),
    :297-306 (in-repo flavour: [CLS] + ids + [SEP]); Hub flavour: ids + [SEP]
    (SURVEY.md A.9).  `max_length` counts the special tokens; truncation keeps the first
    bases (HF `truncation=True` == longest_first on a single sequence, right side).
    """
    ids = [VOCAB.get(ch, VOCAB["[UNK]"]) for ch in seq]
    n_special = int(add_cls) + int(add_sep)
    if max_length is not None and len(ids) + n_special > max_length:
        ids = ids[: max(max_length - n_special, 0)]
    return ([VOCAB["[CLS]"]] if add_cls else []) + ids + ([VOCAB["[SEP]"]] if add_sep else [])


def encode_read_name(name: str, max_id_length: int = 256) -> list[int]:
    """chimeralm/data/tokenizer.py:108-111: [len(name)] + ord(chars), cut/padded to 256."""
    new_id = [len(name)] + [ord(c) for c in name]
    if len(new_id) > max_id_length:
        return new_id[:max_id_length]
    return new_id + [0] * (max_id_length - len(new_id))


def collate(list_of_ids: list[list[int]], *, padding_side: str, pad_id: int = 4) -> list[list[int]]:
    """DataCollator.torch_call -> tokenizer.pad(padding=True): pad to the longest member of
    the batch with [PAD]=4 on `padding_side` (chimeralm/data/tokenizer.py:152-159)."""
    tmax = max(len(x) for x in list_of_ids)
    out = []
    for x in list_of_ids:
        fill = [pad_id] * (tmax - len(x))
        out.append(fill + x if padding_side == "left" else x + fill)
    return out


def resume_read_name(row: list[int]) -> str:
    """chimeralm/models/callbacks.py:38-63 on one int8[256] row."""
    if not row:
        return ""
    n = row[0]
    if n <= 0 or n >= len(row):
        raise ValueError("Invalid read name data")
    return "".join(chr(b) for b in row[1 : 1 + n] if 32 <= b <= 126)


def prediction_lines(names: list[str], labels: list[int]) -> list[str]:
    """chimeralm/models/callbacks.py:137-139: "{read_name}\\t{label}\\n"."""
    return [f"{n}\t{int(l)}\n" for n, l in zip(names, labels)]
