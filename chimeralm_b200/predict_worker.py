"""Per-rank body of `chimeralm predict` (one process per GPU).

Lives in its own module, not in `__main__`: `torch.multiprocessing.spawn` pickles the target function by qualified name and
the children of `python -m chimeralm_b200` cannot import it from `__main__`.
"""

from __future__ import annotations

import logging
import os
from pathlib import Path

log = logging.getLogger("chimeralm")


def predict_rank(rank: int, world: int, data_path: Path, output_path: Path, batch_size: int, ckpt_path, seed_weights,
                  max_sample, bucket: bool, port: int, num_workers: int = 0):
    import torch

    from .callbacks import PredictionWriter
    from .data import BamDataModule, Trainer
    from .model import ChimeraLM
    from .tokenizer import load_tokenizer_from_hyena_model

    dist = None
    if world > 1:
        import torch.distributed as dist

        # explicit TCP rendezvous on the loopback: independent of any RANK / MASTER_* / TORCHELASTIC_* variables the parent
        # process may carry (run from inside a torchrun worker, the env:// rendezvous would wait on torchrun's own store)
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                                device_id=torch.device("cuda", rank))
    tokenizer = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    if ckpt_path is not None:
        model = ChimeraLM.from_pretrained(str(ckpt_path), device=rank)
    else:
        model = ChimeraLM.new(seed=seed_weights, device=rank)
    datamodule = BamDataModule(train_data_path=Path("dummy.bam"), tokenizer=tokenizer, predict_data_path=data_path,
                               batch_size=batch_size, max_predict_samples=max_sample, engine=model.engine,
                               bucket_by_length=bucket, streaming=not bucket, num_workers=num_workers)
    callbacks = [PredictionWriter(output_dir=output_path, write_interval="batch")]
    trainer = Trainer(accelerator="gpu", devices=world, callbacks=callbacks, logger=False, rank=rank, world_size=world)
    trainer.predict(model=model, dataloaders=datamodule, return_predictions=False)
    if dist is not None:
        # the path's single exchange: gather (read index, label) to every rank; rank 0 reports
        pairs = [(i, int(l)) for idx, labs in trainer.last_results for i, l in zip(idx, labs.tolist())]
        mine = torch.tensor(pairs, dtype=torch.int32, device=f"cuda:{rank}").reshape(-1, 2)
        counts = [torch.zeros(1, dtype=torch.int64, device=mine.device) for _ in range(world)]
        dist.all_gather(counts, torch.tensor([mine.shape[0]], device=mine.device))
        mx = int(max(c.item() for c in counts))
        pad = torch.full((mx, 2), -1, dtype=torch.int32, device=mine.device)
        pad[: mine.shape[0]] = mine
        out = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(out, pad)
        if rank == 0:
            allp = torch.cat([o[: int(c.item())] for o, c in zip(out, counts)]).cpu()
            log.info(f"Gathered {allp.shape[0]} predictions from {world} ranks")
        dist.destroy_process_group()
