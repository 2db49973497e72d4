"""Predict-side data module: mirror of `chimeralm/data/bam.py::BamDataModule` (reference
:78-109,129-174,287-299) without Lightning / HF datasets / pysam.

Batching policy is the reference's: file order, `batch_size // world_size` reads per device,
`shuffle=False`, last batch short, every batch padded to its longest member on the
tokenizer's `padding_side`.  Under data parallelism rank r takes samples r, r+W, r+2W, ...
(Lightning's unrepeated distributed sampler in predict mode).  `bucket_by_length=True` is the
B200-side option: reads are sorted by length and cut into batches under a padded-token budget
(`batch_size x longest read`, so short reads travel many to a batch), and the batches are dealt to
the ranks by greedy longest-processing-time on a cost model of the step, so every GPU gets the
same amount of work instead of the same number of reads (`chimeralm_b200.synth.bucket_batches /
deal_lpt`; replaces the reference's policy at chimeralm/data/bam.py:142-146,287-299).

When an `Engine` is attached the batch's `input_ids` are produced on the GPU by
`clm_encode_batch` from raw bases staged in pinned memory (uint8 ids, device tensor);
without one the module collates on the host exactly like `DataCollator.torch_call`.
"""

from __future__ import annotations

from pathlib import Path

import numpy as np
import torch

from . import synth
from .ingest import read_bam_flat
from .tokenizer import DataCollator, encode_read_name, encode_read_name_rows, names_to_rows


class PredictDataset:
    """Kept reads in file order, already truncated to `max_bases`: names plus ONE flat uint8 array
    of ASCII bases with int64 offsets (the layout the native ingest writes and the CUDA encoder
    reads); `seqs[i]` is a view into it."""

    def __init__(self, names: list[str], flat: np.ndarray, offsets: np.ndarray):
        self.names, self.flat, self.offsets = names, flat, offsets
        self.lengths = np.diff(offsets)
        self._seqs = None

    @classmethod
    def from_lists(cls, names: list[str], seqs: list[np.ndarray]):
        offsets = np.zeros(len(seqs) + 1, np.int64)
        np.cumsum([len(s) for s in seqs], out=offsets[1:])
        return cls(names, np.concatenate(seqs) if offsets[-1] else np.zeros(0, np.uint8), offsets)

    @property
    def id_rows(self) -> np.ndarray:
        """int8 [n, 256] read-name rows of the whole dataset (`encode_read_name` per read, vectorised)."""
        if getattr(self, "_id_rows", None) is None:
            if any(len(x) > 254 for x in self.names):   # longer than a BAM name can be: the scalar rule keeps len(name)
                self._id_rows = np.array([encode_read_name(x) for x in self.names], dtype=np.int64).astype(np.int8).reshape(-1, 256)
            else:
                self._id_rows = encode_read_name_rows(names_to_rows(self.names))
        return self._id_rows

    @property
    def seqs(self) -> list[np.ndarray]:
        if self._seqs is None:
            o = self.offsets
            self._seqs = [self.flat[o[i] : o[i + 1]] for i in range(len(self.names))]
        return self._seqs

    def __len__(self):
        return len(self.names)


_UPPER = np.arange(256, dtype=np.uint8)
_UPPER[ord("a") : ord("z") + 1] -= 32


def read_parquet_records(path, max_records=None):
    """Predict input as the authors ran large evaluations (`chimeralm/data/fq.py:152-181`, HF `load_dataset("parquet")`):
    columns `id` and `seq` (a `qual` column, if present, is ignored exactly like `input_quals` is)."""
    import pyarrow.parquet as pq

    n = 0
    pf = pq.ParquetFile(str(path))
    for rg in range(pf.num_row_groups):
        tbl = pf.read_row_group(rg, columns=["id", "seq"])
        for name, seq in zip(tbl.column("id").to_pylist(), tbl.column("seq").to_pylist()):
            yield name, np.frombuffer(seq.encode("ascii", "replace"), dtype=np.uint8)
            n += 1
            if max_records is not None and n >= max_records:
                return


def read_fastq_bytes(path):
    """FASTQ(.gz) records as (name, uint8 bases), upper-cased like `pyfastx.Fastx(uppercase=True)`
    (`chimeralm/data/only_fq.py:22-41`); the name is the header up to the first whitespace."""
    import gzip

    op = gzip.open if str(path).endswith(".gz") else open
    with op(str(path), "rb") as f:
        while True:
            h = f.readline()
            if not h:
                return
            s = f.readline().rstrip(b"\r\n")
            f.readline()
            f.readline()
            yield h[1:].split()[0].decode("ascii", "replace"), _UPPER[np.frombuffer(s, dtype=np.uint8)]


class BamDataModule:
    def __init__(self, tokenizer, train_data_path=None, batch_size: int = 12, val_data_path=None, test_data_path=None,
                 predict_data_path=None, num_workers: int = 1, max_train_samples=None, max_val_samples=None,
                 max_test_samples=None, max_predict_samples=None, *, pin_memory: bool = False, engine=None,
                 bucket_by_length: bool = False, rank: int = 0, world_size: int = 1, streaming: bool = False,
                 prefetch_batches: int = 3):
        self.tokenizer = tokenizer
        self.batch_size = batch_size
        self.predict_data_path = predict_data_path
        self.max_predict_samples = max_predict_samples
        self.engine = engine
        self.bucket_by_length = bucket_by_length
        self.num_workers = max(0, int(num_workers or 0))  # ingest threads; 0 = all host cores
        self.rank, self.world_size = rank, world_size
        # streaming: batches are decoded by a producer thread straight from the BAM into a ring of
        # pinned buffers while the GPU works on earlier ones (file order, so no length bucketing)
        self.streaming = streaming and not bucket_by_length
        self.prefetch_batches = max(1, prefetch_batches)
        self.batch_size_per_device = batch_size
        self.data_collator = DataCollator(tokenizer)
        self.data_predict: PredictDataset | None = None

    @property
    def num_classes(self) -> int:
        return 2

    def setup(self, stage: str | None = None) -> None:
        if self.batch_size % self.world_size != 0:
            raise RuntimeError(f"Batch size ({self.batch_size}) is not divisible by the number of devices ({self.world_size}).")
        self.batch_size_per_device = self.batch_size // self.world_size
        if stage != "predict":
            raise NotImplementedError("chimeralm_b200 implements the predict stage only (training is out of scope)")
        if not self.predict_data_path:
            raise ValueError("Predict data path is required for prediction stage.")
        path = Path(self.predict_data_path)
        max_bases = self.tokenizer.max_len_single_sentence - self.tokenizer.num_special_tokens
        if self.streaming and path.suffix not in (".fq", ".fastq", ".gz", ".parquet"):
            self.data_predict = None  # read on the fly by _stream_batches
            return
        self.streaming = False
        if path.suffix in (".fq", ".fastq", ".gz", ".parquet"):
            names, seqs = [], []
            for name, seq in (read_parquet_records(path) if path.suffix == ".parquet" else read_fastq_bytes(path)):
                names.append(name)
                seqs.append(seq[:max_bases])
                if self.max_predict_samples is not None and len(names) >= self.max_predict_samples:
                    break
            self.data_predict = PredictDataset.from_lists(names, seqs)
        else:
            # Native ingest: BGZF inflate on `num_workers` threads (0 = all cores), is_chimeric filter
            # and base decode in C++ (csrc/bam_ingest.cpp) - no per-read Python object.
            names, flat, offsets = read_bam_flat(path, max_bases, self.max_predict_samples, chimeric_only=True,
                                                 n_threads=self.num_workers)
            self.data_predict = PredictDataset(names, flat, offsets)

    # -------------------------------------------------------------------------------------
    def _rank_batches(self) -> list[np.ndarray]:
        """Index arrays of this rank's batches, in the order they are run."""
        n = len(self.data_predict)
        bs = self.batch_size_per_device
        if not self.bucket_by_length:
            # the reference's policy: file order, Lightning's unrepeated sampler (rank r takes reads r, r + W, ...), fixed size
            idx = np.arange(self.rank, n, self.world_size)
            return [idx[i : i + bs] for i in range(0, len(idx), bs)]
        lens = np.asarray(self.data_predict.lengths, dtype=np.int64)
        ns = self.tokenizer.num_special_tokens
        budget = bs * (int(lens.max()) + ns) if n else 0
        batches = synth.bucket_batches(lens, max(bs, min(8 * bs, 256)), n_special=ns, max_tokens_per_batch=budget)
        costs = [synth.batch_cost(len(b), int(lens[b].max()) + ns) for b in batches]
        ranks, _ = synth.deal_lpt(costs, self.world_size)
        return [batches[i] for i in ranks[self.rank]]

    def max_batch_shape(self) -> tuple[int, int, int]:
        """(max reads, max tokens, max padded tokens) over this rank's batches: what the engine has to reserve."""
        ns = self.tokenizer.num_special_tokens
        lens = self.data_predict.lengths
        shapes = [(len(b), int(lens[b].max()) + ns) for b in self._rank_batches()] or [(1, 1)]
        return max(s[0] for s in shapes), max(s[1] for s in shapes), max(s[0] * s[1] for s in shapes)

    def _stream_batches(self):
        """Producer/consumer loader: a thread pulls this rank's reads from the native BAM reader
        (`clm_bam_next` with `clm_bam_set_shard`) into pinned buffers `prefetch_batches` ahead."""
        import queue
        import threading

        from .ingest import NAME_STRIDE, NativeBamReader

        tok, bs, W, rank = self.tokenizer, self.batch_size_per_device, self.world_size, self.rank
        ns = tok.num_special_tokens
        max_bases = tok.max_len_single_sentence - ns
        pin = self.engine is not None
        depth = self.prefetch_batches + 3  # queue + one being filled + one in flight + one being written
        ring = [(torch.empty(bs * max_bases, dtype=torch.uint8, pin_memory=pin),
                 torch.empty(bs + 1, dtype=torch.int64, pin_memory=pin),
                 np.empty((bs, NAME_STRIDE), np.uint8)) for _ in range(depth)]
        q: queue.Queue = queue.Queue(maxsize=self.prefetch_batches)
        limit = self.max_predict_samples
        stop = threading.Event()

        def put(item) -> bool:
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.1)
                    return True
                except queue.Full:
                    continue
            return False

        def produce():
            try:
                with NativeBamReader(self.predict_data_path, self.num_workers) as rd:
                    rd.set_shard(rank, W)
                    done, k = 0, 0
                    while True:
                        want = bs
                        if limit is not None:  # global cap: this rank owns indices rank, rank+W, ... < limit
                            mine = max(0, (limit - rank + W - 1) // W)
                            want = min(bs, mine - done)
                        if want <= 0:
                            break
                        bases, offs, nm = ring[k % depth]
                        n = rd.next_block(want, max_bases, bases.numpy(), offs.numpy(), nm, True)
                        if n == 0:
                            break
                        if not put((k, n, done)):
                            return
                        done += n
                        k += 1
                put(None)
            except BaseException as e:  # noqa: BLE001 - surfaced on the consumer side
                put(e)

        th = threading.Thread(target=produce, name="clm-bam-ingest", daemon=True)
        th.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                k, n, done = item
                bases, offs, nm = ring[k % depth]
                o = offs.numpy()
                T = int(np.diff(o[: n + 1]).max()) + ns
                id_rows = encode_read_name_rows(nm[:n])   # one numpy pass, no per-read Python
                batch = {"id": torch.from_numpy(id_rows), "labels": torch.full((n,), -1, dtype=torch.int64),
                         "indices": np.arange(rank + W * done, rank + W * (done + n), W)}
                nb = int(o[n])
                if self.engine is not None:
                    dev = self.engine.device
                    ids, _ = self.engine.encode(bases[: max(nb, 1)].to(dev, non_blocking=True),
                                                offs[: n + 1].to(dev, non_blocking=True), T, add_cls=tok.add_cls,
                                                add_sep=tok.add_sep, pad_left=tok.padding_side == "left",
                                                max_bases=max_bases)
                    batch["input_ids"] = ids
                else:
                    b = bases.numpy()
                    feats = [{"input_ids": tok.encode_array(b[o[i] : o[i + 1]].tobytes(), max_length=tok.max_len_single_sentence)}
                             for i in range(n)]
                    batch["input_ids"] = tok.pad(feats, return_tensors="pt")["input_ids"]
                yield batch
        finally:
            stop.set()
            th.join(timeout=5)

    def predict_dataloader(self):
        if self.streaming:
            yield from self._stream_batches()
            return
        ds = self.data_predict
        tok = self.tokenizer
        ns = tok.num_special_tokens
        batches = self._rank_batches()
        ring = None
        if self.engine is not None and batches:
            # persistent pinned staging, three deep: batch k + 1 is assembled while batch k's copy may still be in flight and
            # batch k - 1 has been consumed (Trainer.predict synchronises on batch k - 1 before asking for batch k + 1)
            nb_max = max(int(ds.lengths[b].sum()) for b in batches)
            ring = [(torch.empty(max(nb_max, 1), dtype=torch.uint8).pin_memory(),
                     torch.empty(max(len(b) for b in batches) + 1, dtype=torch.int64).pin_memory()) for _ in range(3)]
            B, T, budget = self.max_batch_shape()
            self.engine.reserve(B, T, budget)
        id_rows_all = ds.id_rows
        for k, sel in enumerate(batches):
            lens = ds.lengths[sel]
            T = int(lens.max()) + ns
            batch = {"id": torch.from_numpy(id_rows_all[sel]), "labels": torch.full((len(sel),), -1, dtype=torch.int64),
                     "indices": sel}
            if self.engine is not None:
                bases, offs = ring[k % 3]
                o = offs.numpy()
                o[0] = 0
                np.cumsum(lens, out=o[1 : len(sel) + 1])
                fb = bases.numpy()
                if len(sel) and sel[-1] - sel[0] == len(sel) - 1 and np.all(np.diff(sel) == 1):
                    fb[: o[len(sel)]] = ds.flat[ds.offsets[sel[0]] : ds.offsets[sel[-1] + 1]]   # consecutive reads: one copy
                else:
                    for j, i in enumerate(sel):
                        fb[o[j] : o[j + 1]] = ds.flat[ds.offsets[i] : ds.offsets[i + 1]]
                nb = int(o[len(sel)])
                dev = self.engine.device
                ids, _ = self.engine.encode(bases[: max(nb, 1)].to(dev, non_blocking=True),
                                            offs[: len(sel) + 1].to(dev, non_blocking=True), T, add_cls=tok.add_cls,
                                            add_sep=tok.add_sep, pad_left=tok.padding_side == "left",
                                            max_bases=tok.max_len_single_sentence - ns)
                batch["input_ids"] = ids
            else:
                feats = [{"input_ids": tok.encode_array(ds.seqs[i].tobytes(), max_length=tok.max_len_single_sentence)} for i in sel]
                batch["input_ids"] = tok.pad(feats, return_tensors="pt")["input_ids"]
            yield batch


class Trainer:
    """The slice of `lightning.Trainer` that `chimeralm predict` uses (chimeralm/__main__.py:307-317)."""

    def __init__(self, accelerator="gpu", devices=1, callbacks=None, deterministic=True, logger=False, rank: int = 0,
                 world_size: int = 1):
        self.callbacks = list(callbacks or [])
        self.global_rank, self.world_size = rank, world_size

    def predict(self, model, dataloaders=None, return_predictions: bool = False, ckpt_path=None):
        from ._lib import Fp16RangeError
        from .weights import load_checkpoint

        if ckpt_path is not None:
            model.load_state_dict(load_checkpoint(ckpt_path))
            if hasattr(dataloaders, "engine") and dataloaders.engine is not None:
                dataloaders.engine = model.engine
        dm = dataloaders
        if hasattr(dm, "setup"):
            dm.rank, dm.world_size = self.global_rank, self.world_size
            dm.setup("predict")
            loader = dm.predict_dataloader()
        else:
            loader = dm
        model.eval()
        results = []

        def to_host(t):
            return torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t, non_blocking=True) if t.is_cuda else t

        def finish(item):
            # Batch k's outputs are consumed (callbacks, files) only after batch k+1 was launched, so
            # the device never waits for the host side of the loop.
            batch_idx, batch, pred, labels, seq, ev = item
            if ev is not None:
                ev.synchronize()
                try:
                    model.engine.forward_status(seq)
                except Fp16RangeError:   # this batch left the fp16 range of the tensor-core conv: redo it in fp32
                    logits, dev_labels = model.engine.forward(batch["input_ids"], return_labels=True, check=True)
                    pred, labels = (logits.cpu(), pred[1]), dev_labels.cpu()
            for cb in self.callbacks:
                cb.write_on_batch_end(self, model, pred, None, batch, batch_idx, 0)
            if return_predictions or self.world_size > 1:
                results.append((batch.get("indices"), labels if labels is not None else pred[0].argmax(1).cpu()))

        pending = None
        for batch_idx, batch in enumerate(loader):
            pred = model.predict_step(batch, batch_idx)
            labels, seq, ev = None, 0, None
            if isinstance(pred, (tuple, list)) and pred[0].is_cuda:
                dev = pred[0].device
                labels, seq = getattr(model, "last_device_labels", None), getattr(model, "last_forward_seq", 0)
                pred = tuple(to_host(t) if isinstance(t, torch.Tensor) else t for t in pred)
                labels = to_host(labels) if labels is not None else None
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
            if pending is not None:
                finish(pending)
            pending = (batch_idx, batch, pred, labels, seq, ev)
        if pending is not None:
            finish(pending)
        self.last_results = results
        return results if return_predictions else None
