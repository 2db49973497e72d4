"""Prediction outputs: mirror of `chimeralm/models/callbacks.py` (reference).

`resume_read_name` (:38-63) and `PredictionWriter.write_on_batch_end` (:79-150) keep the
reference's names, argument order, file naming (`{global_rank}_{batch_idx}.txt`, :134), line
format (`name\\tlabel\\n`, :137-139) and its log-and-continue error behaviour, without
Lightning: `trainer` only needs a `global_rank` attribute.
"""

from __future__ import annotations

import logging
from pathlib import Path
from typing import Any

import numpy as np
import torch

logger = logging.getLogger(__name__)


def resume_read_name(bytes_data) -> str:
    """int8[256] row `[len, ord(c)..., 0...]` -> read name (printable ASCII 32..126 only)."""
    if isinstance(bytes_data, torch.Tensor):
        if bytes_data.numel() == 0:
            return ""
        bytes_data = bytes_data.tolist()
    elif not bytes_data:
        return ""
    try:
        n = bytes_data[0]
        if n <= 0 or n >= len(bytes_data):
            raise ValueError("Invalid read name length")
        return "".join(chr(b) for b in bytes_data[1 : 1 + n] if 32 <= b <= 126)
    except (IndexError, TypeError, ValueError) as e:
        raise ValueError("Invalid read name data") from e


def resume_read_names(batch_ids) -> list:
    """`resume_read_name` for a whole `[B, 256]` int8 batch in one numpy pass; an entry is the name, or the ValueError the
    scalar function would have raised for that row (bad length byte)."""
    a = batch_ids.numpy() if isinstance(batch_ids, torch.Tensor) else np.asarray(batch_ids)
    if a.ndim != 2 or a.shape[1] < 2:
        return [_try_resume(r) for r in batch_ids]
    a = a.astype(np.int64)
    n = a[:, 0]
    keep = (np.arange(1, a.shape[1])[None, :] <= n[:, None]) & (a[:, 1:] >= 32) & (a[:, 1:] <= 126)
    chars = np.where(keep, a[:, 1:], 0).astype(np.uint8)
    out = []
    for i in range(a.shape[0]):
        if n[i] <= 0 or n[i] >= a.shape[1]:
            out.append(ValueError("Invalid read name data"))
        else:
            out.append(chars[i][keep[i]].tobytes().decode("ascii"))
    return out


def _try_resume(row):
    try:
        return resume_read_name(row)
    except ValueError as e:
        return e


class PredictionWriter:
    """Writes one `{rank}_{batch_idx}.txt` per batch (write_interval == "batch")."""

    def __init__(self, output_dir: str | Path, write_interval: str = "batch") -> None:
        self.interval = write_interval
        self.output_dir = Path(output_dir)  # Path(None) raises TypeError, as in the reference (N2)

    def write_on_batch_end(self, trainer: Any, pl_module: Any, prediction: Any, batch_indices: Any,
                           batch: dict[str, Any], batch_idx: int, dataloader_idx: int) -> None:
        try:
            if prediction is None or len(prediction) == 0:
                logger.warning(f"Empty prediction for batch {batch_idx}, dataloader {dataloader_idx}")
                return
            if "id" not in batch:
                logger.error(f"Missing 'id' key in batch {batch_idx}, dataloader {dataloader_idx}")
                return
            pred_tensor = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
            if pred_tensor is None or pred_tensor.numel() == 0:
                logger.warning(f"Empty prediction tensor for batch {batch_idx}")
                return
            # labels = argmax(dim=1) of the logits, no softmax (reference :107)
            predictions_cpu = pred_tensor.argmax(dim=1).cpu()
            batch_ids = batch["id"]
            if len(predictions_cpu) != len(batch_ids):
                logger.error(f"Size mismatch: predictions={len(predictions_cpu)}, batch_ids={len(batch_ids)} for batch {batch_idx}")
                return
            read_names = []
            for i, name in enumerate(resume_read_names(batch_ids)):   # same rule as resume_read_name per row, vectorised
                if isinstance(name, Exception):   # reference behaviour: log and continue
                    logger.error(f"Error processing read name at index {i}: {name}")
                    read_names.append(f"error_read_{i}")
                    continue
                if not name:
                    name = f"unknown_read_{i}"
                    logger.warning(f"Empty read name for index {i} in batch {batch_idx}")
                read_names.append(name)
            if not self.output_dir.exists():
                self.output_dir.mkdir(parents=False, exist_ok=True)
            output_file = self.output_dir / f"{trainer.global_rank}_{batch_idx}.txt"
            try:
                lines = [f"{n}\t{p}\n" for n, p in zip(read_names, predictions_cpu.tolist(), strict=True)]
                with output_file.open("w") as f:
                    f.writelines(lines)
            except OSError as e:
                logger.error(f"Failed to write predictions to {output_file}: {e}")
        except Exception as e:  # noqa: BLE001
            logger.error(f"Critical error in write_on_batch_end for batch {batch_idx}: {e}")


def load_predicts(path: Path | str) -> dict[str, int]:
    """`chimeralm/__main__.py:26-61`: parse one prediction file (exactly two tab-separated fields)."""
    predicts: dict[str, int] = {}
    try:
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(f"File not found: {path}")
        with path.open(encoding="utf-8") as f:
            for line_num, line in enumerate(f, 1):
                line = line.strip()
                if not line:
                    continue
                parts = line.split("\t")
                if len(parts) != 2:
                    raise ValueError(f"Invalid line format at line {line_num}: {line}")
                predicts[parts[0]] = int(parts[1])
    except Exception as e:
        raise ValueError(f"Error reading file {path}: {e}") from e
    return predicts


def load_predictions_from_folder(path: Path | str) -> dict[str, int]:
    """`chimeralm/__main__.py:64-69`: union over every `*.txt` (later files override)."""
    predictions: dict[str, int] = {}
    for file in Path(path).glob("*.txt"):
        predictions.update(load_predicts(file))
    return predictions
