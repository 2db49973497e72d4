"""Generate the golden fixtures under tests/golden/ from the REFERENCE ITSELF.

Runs only in the build container (needs /root/reference, which the GPU box lacks):

    python oracle/make_golden.py

It imports, by file path and without copying any source,
  * /root/reference/chimeralm/data/tokenizer.py  (CharacterTokenizer, DataCollator,
    tokenize_and_align_labels_and_quals_ids) and
  * /root/reference/chimeralm/models/components/hyena.py (BinarySequenceClassifier)
and records their outputs on seeded inputs.  `import chimeralm` as a package is
impossible offline (gradio/lightning/pysam missing), and the backbone is HF-Hub remote
code that is not in the reference tree, so no backbone outputs can be generated.
"""

from __future__ import annotations

import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference")
OUT = ROOT / "tests" / "golden"


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def tokenizer_golden():
    m = _load("ref_tokenizer", "chimeralm/data/tokenizer.py")
    rng = np.random.default_rng(7)
    alphabet = np.array(list("ACGTNacgtnRYKM=*-X"))
    probs = np.array([0.22, 0.22, 0.22, 0.22, 0.04] + [0.08 / 13] * 13)
    seqs = ["ATCG", "", "A", "ACGTNXacgt", "N" * 7]
    for n in (5, 31, 96, 97, 98, 99, 130, 257):
        seqs.append("".join(rng.choice(alphabet, size=n, p=probs)))
    cases = []
    for mml, side in ((None, "right"), (100, "left"), (100, "right"), (34, "left")):
        tok = m.CharacterTokenizer(model_max_length=mml, padding_side=side)
        max_len = tok.max_len_single_sentence if mml is not None else None
        enc = []
        for s in seqs:
            if mml is None:
                enc.append(tok(s)["input_ids"])
            else:
                enc.append(tok(s, truncation=True, max_length=max_len, padding=True)["input_ids"])
        entry = {"model_max_length": mml, "padding_side": side, "max_len_single_sentence": max_len,
                 "pad_token_id": tok.pad_token_id, "input_ids": enc}
        if mml is not None:
            feats = [dict(m.tokenize_and_align_labels_and_quals_ids({"seq": s, "id": f"read-{i}/{len(s)}"}, tok, max_len))
                     for i, s in enumerate(seqs)]
            batch = m.DataCollator(tok).torch_call(feats)
            entry["collated_input_ids"] = batch["input_ids"].tolist()
            entry["collated_id"] = batch["id"].tolist()
            entry["collated_labels"] = batch["labels"].tolist()
            entry["collated_dtypes"] = {k: str(v.dtype) for k, v in batch.items()}
        cases.append(entry)
    names = ["r", "read/1", "a" * 127, "m64011_190830_220126/1/ccs;x|y", "tab\tname", "b" * 126 + "\x7f"]
    name_rows = [m.tokenize_and_align_labels_and_quals_ids({"seq": "A", "id": n}, m.CharacterTokenizer(model_max_length=10), 8)["id"]
                 for n in names]
    return {"source": "reference chimeralm/data/tokenizer.py imported by path (transformers %s)" % __import__("transformers").__version__,
            "seqs": seqs, "cases": cases, "names": names, "name_rows": name_rows,
            "kat_ATCG": m.CharacterTokenizer().encode("ATCG")}


def head_golden():
    from chimeralm_b200.weights import HEAD_PREFIX, make_state_dict, perturb_norms

    m = _load("ref_hyena", "chimeralm/models/components/hyena.py")
    head = m.BinarySequenceClassifier(input_dim=256, hidden_dim=512, num_layers=2, dropout=0.1,
                                      pooling_type="attention", activation="gelu", use_residual=True,
                                      save_attention=True).eval()
    out = {}
    for tag, seed, shape in (("a", 0, (3, 37, 256)), ("b", 5, (2, 1025, 256))):
        sd = perturb_norms(make_state_dict(seed), seed + 1)
        hsd = {k[len(HEAD_PREFIX):]: v for k, v in sd.items() if k.startswith(HEAD_PREFIX)}
        missing, unexpected = head.load_state_dict(hsd, strict=True)
        assert not missing and not unexpected
        g = torch.Generator().manual_seed(100 + seed)
        hidden = torch.randn(shape, generator=g)
        with torch.inference_mode():
            logits = head(hidden, None)
        out[f"{tag}_seed"] = np.array(seed)
        out[f"{tag}_shape"] = np.array(shape)
        out[f"{tag}_logits"] = logits.numpy()
        out[f"{tag}_attn"] = head.attention_weights.squeeze(-1).numpy()
    # the head's other pooling modes (components/hyena.py:97-115,134-136): no scorer in the module, same classifier weights
    for pooling in ("mean", "max", "cls"):
        alt = m.BinarySequenceClassifier(input_dim=256, hidden_dim=512, num_layers=2, dropout=0.1, pooling_type=pooling,
                                         activation="gelu", use_residual=True).eval()
        sd = perturb_norms(make_state_dict(0), 1)
        hsd = {k[len(HEAD_PREFIX):]: v for k, v in sd.items() if k.startswith(HEAD_PREFIX) and ".attention." not in k}
        alt.load_state_dict(hsd, strict=True)
        hidden = torch.randn((3, 37, 256), generator=torch.Generator().manual_seed(100))
        with torch.inference_mode():
            out[f"a_logits_{pooling}"] = alt(hidden, None).numpy()
    out["head_param_count"] = np.array(sum(p.numel() for p in head.parameters()))
    out["head_keys"] = np.array(sorted(head.state_dict().keys()))
    return out


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "tokenizer_golden.json").write_text(json.dumps(tokenizer_golden()))
    np.savez_compressed(OUT / "head_golden.npz", **head_golden())
    print("wrote", sorted(p.name for p in OUT.iterdir()))


if __name__ == "__main__":
    main()
