"""K3-style workload: reads with log-normal lengths clip(round(exp(N(ln 6000, 0.75^2))), 1000, 32768) in a BAM,
predicted through the CLI-equivalent flow (BamDataModule -> Trainer.predict -> PredictionWriter), file order vs length
bucketing.  python profiles/k3_mixed_lengths.py [n_reads] [batch]"""
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from chimeralm_b200.bam import BamWriter, make_record, minimal_header  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8000
    bs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    import torch

    from chimeralm_b200.callbacks import PredictionWriter, load_predictions_from_folder
    from chimeralm_b200.data import BamDataModule, Trainer
    from chimeralm_b200.model import ChimeraLM
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    rng = np.random.default_rng(20251019)
    lens = np.clip(np.round(np.exp(rng.normal(np.log(6000), 0.75, n))), 1000, 32768).astype(int)
    tmp = Path(tempfile.mkdtemp())
    bam = tmp / "k3.bam"
    acgt = np.frombuffer(b"ACGT", np.uint8)
    t = time.time()
    w = BamWriter(bam, minimal_header())
    for i, L in enumerate(lens):
        w.write(make_record(f"synth-{i:09d}", acgt[rng.integers(0, 4, L)].tobytes()))
    w.close()
    print(f"K3 BAM: {n} reads, {lens.sum() / 1e6:.1f} Mbases (mean {lens.mean():.0f}, max {lens.max()}), {bam.stat().st_size / 1e6:.0f} MB, "
          f"written in {time.time() - t:.1f}s")
    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    model = ChimeraLM.new(seed=0, device=0, max_batch=bs, max_tokens=32769)
    results = {}
    for label, kw in (("file order, streaming", dict(streaming=True)), ("length-bucketed", dict(bucket_by_length=True))):
        for rep in range(2):
            out = tmp / f"pred_{label[:4]}_{rep}"
            out.mkdir()
            dm = BamDataModule(train_data_path=Path("dummy.bam"), tokenizer=tok, predict_data_path=bam, batch_size=bs,
                               engine=model.engine, num_workers=0, **kw)
            tr = Trainer(callbacks=[PredictionWriter(output_dir=out)])
            t0 = time.time()
            tr.predict(model=model, dataloaders=dm, return_predictions=False)
            torch.cuda.synchronize()
            dt = time.time() - t0
        results[label] = load_predictions_from_folder(out)
        print(f"{label:24s}: {dt:.2f}s  {n / dt:,.0f} reads/s  {lens.sum() / dt / 1e6:,.1f} Mbases/s (BAM -> label files, second run)")
    a, b = results.values()
    print(f"labels agree between the two batchings: {sum(a[k] == b[k] for k in a)} / {len(a)}")


if __name__ == "__main__":
    main()
