"""CPU multi-process test (-m "not gpu"): the N>1 predict path's host logic with gloo, world_size 2.
Each rank takes its shard of the golden BAM exactly as `BamDataModule` deals it (samples rank::W),
produces fake labels = f(read index), writes its `{rank}_{batch}.txt` files through PredictionWriter and
joins the single end-of-run all_gather; rank 0 checks that the union covers every read once."""

import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
BAM = ROOT / "tests" / "golden" / "test_chimric_reads.bam"


def _worker(rank, world, port, outdir, bucket=False):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chimeralm_b200.callbacks import PredictionWriter
    from chimeralm_b200.data import BamDataModule
    from chimeralm_b200.tokenizer import load_tokenizer_from_hyena_model

    tok = load_tokenizer_from_hyena_model("hyenadna-small-32k-seqlen")
    dm = BamDataModule(tok, batch_size=12, predict_data_path=BAM, rank=rank, world_size=world, bucket_by_length=bucket)
    dm.setup("predict")
    writer = PredictionWriter(outdir, "batch")

    class Tr:
        global_rank = rank

    pairs = []
    tokens = 0
    for bi, batch in enumerate(dm.predict_dataloader()):
        tokens += batch["input_ids"].numel()
        labels = torch.tensor([int(i) % 2 for i in batch["indices"]])
        logits = torch.stack([1.0 - labels.float(), labels.float()], dim=1)
        writer.write_on_batch_end(Tr(), None, (logits, batch["labels"]), None, batch, bi, 0)
        pairs += [(int(i), int(l)) for i, l in zip(batch["indices"], labels)]
    mine = torch.tensor(pairs, dtype=torch.int32).reshape(-1, 2)
    counts = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([mine.shape[0]]))
    mx = int(max(c.item() for c in counts))
    pad = torch.full((mx, 2), -1, dtype=torch.int32)
    pad[: mine.shape[0]] = mine
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    loads = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(loads, torch.tensor([tokens]))
    if rank == 0 and bucket:   # K4 policy: padded tokens per rank within 25 % of each other on 100 reads of 0.5 - 33 k tokens
        ld = [int(x.item()) for x in loads]
        assert max(ld) / (sum(ld) / world) < 1.25, ld
    if rank == 0:
        allp = torch.cat([o[: int(c.item())] for o, c in zip(out, counts)])
        idx = sorted(allp[:, 0].tolist())
        assert idx == list(range(100)), "every read exactly once across ranks"
        assert all(int(l) == int(i) % 2 for i, l in allp.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_write_gather(tmp_path):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path / "pred")), nprocs=2, join=True)
    from chimeralm_b200.bam import parse_bam_file
    from chimeralm_b200.callbacks import load_predictions_from_folder

    preds = load_predictions_from_folder(tmp_path / "pred")
    names = [r["id"] for r in parse_bam_file(BAM)]
    assert set(preds) == set(names) and len(preds) == 100
    assert all(preds[n] == i % 2 for i, n in enumerate(names))
    files = sorted(p.name for p in (tmp_path / "pred").glob("*.txt"))
    # per-device batch = 12 // 2 = 6 -> 50 reads per rank -> 9 files per rank named {rank}_{batch}.txt
    assert files == sorted(f"{r}_{b}.txt" for r in range(2) for b in range(9))


def test_two_rank_bucketed_lpt_dealing_and_gather(tmp_path):
    """The K4 policy on CPU with gloo: length-bucketed batches under a token budget, dealt to the two ranks by LPT; every
    read is predicted exactly once, the per-rank padded-token loads are balanced, the files of both ranks carry all reads."""
    port = 29900 + os.getpid() % 90
    mp.spawn(_worker, args=(2, port, str(tmp_path / "pred"), True), nprocs=2, join=True)
    from chimeralm_b200.bam import parse_bam_file
    from chimeralm_b200.callbacks import load_predictions_from_folder

    preds = load_predictions_from_folder(tmp_path / "pred")
    names = [r["id"] for r in parse_bam_file(BAM)]
    assert set(preds) == set(names) and len(preds) == 100
    assert all(preds[n] == i % 2 for i, n in enumerate(names))
    assert {p.name.split("_")[0] for p in (tmp_path / "pred").glob("*.txt")} == {"0", "1"}
