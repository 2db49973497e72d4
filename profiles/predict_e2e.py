"""End-to-end `chimeralm predict` on a synthetic BAM: ingest thread scaling, then the whole
CLI flow (BamDataModule -> Trainer.predict -> PredictionWriter) timed by phase.

    python profiles/predict_e2e.py [n_reads] [read_len] [batch]
"""
import os
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from chimeralm_b200.bam import BamWriter, make_record, minimal_header  # noqa: E402
from chimeralm_b200.ingest import read_bam_flat  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    bs = int(sys.argv[3]) if len(sys.argv) > 3 else 32
    tmp = Path(tempfile.mkdtemp())
    bam = tmp / "synth.bam"
    rng = np.random.default_rng(0)
    t = time.time()
    w = BamWriter(bam, minimal_header())
    acgt = np.frombuffer(b"ACGT", np.uint8)
    for i in range(n):
        w.write(make_record(f"read_{i:07d}", acgt[rng.integers(0, 4, L)].tobytes()))
    w.close()
    print(f"synthetic BAM: {n} reads x {L} bases, {bam.stat().st_size / 1e6:.1f} MB, written in {time.time() - t:.1f}s; "
          f"host cores {os.cpu_count()}")
    for nt in (1, 2, 4, 8, 16, 0):
        t = time.time()
        names, flat, offs = read_bam_flat(bam, 32768, n_threads=nt)
        dt = time.time() - t
        print(f"ingest threads={nt or 'all'}: {dt:.3f}s  {len(names) / dt:,.0f} reads/s  {flat.size / dt / 1e6:,.0f} Mbases/s")

    import torch

    if not torch.cuda.is_available():
        return
    from chimeralm_b200.callbacks import PredictionWriter, load_predictions_from_folder
    from chimeralm_b200.data import BamDataModule, Trainer
    from chimeralm_b200.model import ChimeraLM
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    t = time.time()
    model = ChimeraLM.new(seed=0, device=0, max_batch=bs, max_tokens=L + 1)
    print(f"model build + finalize: {time.time() - t:.2f}s")
    for rep in range(2):
        out = tmp / f"pred{rep}"
        out.mkdir()
        dm = BamDataModule(train_data_path=Path("dummy.bam"), tokenizer=tok, predict_data_path=bam, batch_size=bs,
                           engine=model.engine, num_workers=0)
        tr = Trainer(callbacks=[PredictionWriter(output_dir=out)])
        t0 = time.time()
        dm.setup("predict")
        t1 = time.time()
        tr.predict(model=model, dataloaders=_Loaded(dm), return_predictions=False)
        torch.cuda.synchronize()
        t2 = time.time()
        preds = load_predictions_from_folder(out)
        print(f"run {rep}: setup(ingest) {t1 - t0:.3f}s  predict loop {t2 - t1:.3f}s  -> {n / (t2 - t1):,.0f} reads/s in the loop, "
              f"{n / (t2 - t0):,.0f} reads/s BAM->labels; {len(preds)} predictions written")
    ref_preds = preds
    for rep in range(2):
        out = tmp / f"pred_stream{rep}"
        out.mkdir()
        dm = BamDataModule(train_data_path=Path("dummy.bam"), tokenizer=tok, predict_data_path=bam, batch_size=bs,
                           engine=model.engine, num_workers=0, streaming=True)
        tr = Trainer(callbacks=[PredictionWriter(output_dir=out)])
        t0 = time.time()
        tr.predict(model=model, dataloaders=dm, return_predictions=False)
        torch.cuda.synchronize()
        t2 = time.time()
        preds = load_predictions_from_folder(out)
        print(f"streaming run {rep}: {t2 - t0:.3f}s -> {n / (t2 - t0):,.0f} reads/s BAM->labels; {len(preds)} predictions, "
              f"identical to the load-all run: {preds == ref_preds}")


class _Loaded:
    """Hands Trainer.predict an already set-up data module's loader (so setup is timed apart)."""

    def __init__(self, dm):
        self.dm = dm

    def __iter__(self):
        return iter(self.dm.predict_dataloader())


if __name__ == "__main__":
    main()
