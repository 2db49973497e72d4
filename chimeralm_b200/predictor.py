"""Single-sequence prediction, the non-UI half of the reference's web predictor (`chimeralm/ui.py:13-79`,
`ChimeraLMPredictor.predict`): clean + validate the sequence, tokenise with the Hub-flavour tokenizer, one forward,
softmax over the two logits, class name + confidence + per-class breakdown.  Gradio itself is out of scope."""

from __future__ import annotations

import torch

from .model import ChimeraLM, ClassificationLit
from .tokenizer import load_tokenizer_from_hyena_model

CLASS_NAMES = ["Biological", "Chimeric Artifact"]


class ChimeraLMPredictor:
    def __init__(self, model: ClassificationLit | None = None, *, ckpt: str | None = None, device: int = 0, seed: int = 0):
        if model is None:
            model = (ChimeraLM.from_pretrained(ckpt, device=device, max_batch=1, max_tokens=32769) if ckpt
                     else ChimeraLM.new(seed=seed, device=device, max_batch=1, max_tokens=32769))
        self.model = model.eval()
        self.tokenizer = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
        self.device = self.model.device

    def predict(self, sequence: str) -> tuple[str, float, dict]:
        """Same return contract as the reference: (prediction or message, confidence, {class: "0.123"})."""
        if not sequence or not sequence.strip():
            return "Please enter a DNA sequence", 0.0, {}
        sequence = sequence.strip().upper()
        if not all(c in "ACGTN" for c in sequence):
            return "Invalid characters in sequence. Only A, C, G, T, N are allowed.", 0.0, {}
        try:
            ids = self.tokenizer(sequence, truncation=True, padding=True, max_length=32768)["input_ids"]
            input_ids = torch.tensor([ids], dtype=torch.int64, device=self.device)
            logits = self.model(input_ids, None)
            probabilities = torch.softmax(logits, dim=-1)
            predicted_class = int(torch.argmax(probabilities, dim=-1).item())
            confidence = probabilities[0][predicted_class].item()
            breakdown = {name: f"{probabilities[0][i].item():.3f}" for i, name in enumerate(CLASS_NAMES)}
            return CLASS_NAMES[predicted_class], confidence, breakdown
        except Exception as e:  # noqa: BLE001 - reference behaviour: report, do not raise (ui.py:77-79)
            return f"Prediction failed: {e}", 0.0, {}
