#!/usr/bin/env python
"""Headline benchmark: `chimeralm predict` throughput (reads/s) on synthetic 8 kb reads.

Contract (see the task brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON
line.  A step = one batch of BATCH reads through the predict hot path (tokenise -> forward ->
labels).  `value` is measured with the read bytes already resident in HBM; `e2e` goes through
the C-ABI host entry point (`clm_predict_host`) with pinned HOST buffers, H2D and D2H inside
the timed region.  `roofline` is for the dominant kernel class, timed live with CUDA events
on the launching stream (`clm_profile_*`).  `cpu_baseline` is the reference-equivalent CPU
predict path (the oracle port: per-base Python tokeniser + collate + fp32 eager PyTorch
forward + argmax) on a bounded sample.

`--impl reference` times that CPU path alone, on the same workload definition.
"""

from __future__ import annotations

import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

READ_LEN = 8192          # K2: fixed 8 kb reads
BATCH = 32               # K2: batch 32
SEED = 20251018
N_DISTINCT = 8           # distinct synthetic batches rotated through the timed region
F_TOK = 6_423_040        # dense FLOP/token
CONV_BYTES_TOK_LAYER = 1536  # long-conv algorithmic bytes/token/layer (bf16 vx in, x0 in, y*x0 out)


def synth_reads(n, length, seed):
    rng = np.random.default_rng(seed)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, length))]


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "src": "fallback"}


class ClockSampler:
    """`nvidia-smi -lms 20` on one GPU.  Started early (nvidia-smi needs up to a second to come up on an 8-GPU box, longer
    than a short timed region), rows are time-stamped and only those taken inside [mark_start, mark_end] are reported."""
    QUERY = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        self.t0 = self.t1 = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def mark_start(self):
        self.t0 = datetime.datetime.now()

    def mark_end(self):
        self.t1 = datetime.datetime.now()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in Path(self.f.name).read_text().strip().splitlines() if r.count(",") >= 9]
        os.unlink(self.f.name)
        if not rows:
            return out

        def when(r):
            try:
                return datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f")
            except ValueError:
                return None

        inside = [r for r in rows if self.t0 and self.t1 and when(r) and self.t0 <= when(r) <= self.t1]
        if inside:
            out["window"] = "timed region"
        else:  # region shorter than the sampling period: the samples taken under load closest to it (warm-up + region)
            before_end = [r for r in rows if self.t1 and when(r) and when(r) <= self.t1]
            inside = before_end[-3:] or rows[-3:]
            out["window"] = "last samples up to the end of the timed region (region shorter than the 20 ms sampling period)"
        rows = inside
        sm = sorted(float(r[2]) for r in rows)
        out["sm_mhz"] = sm[len(sm) // 2]
        out["sm_max_mhz"] = float(rows[0][3])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        out["reasons"] = [n for i, n in enumerate(names) if any(r[6 + i].strip().lower().startswith("active") for r in rows)]
        out["samples"] = len(rows)
        return out


def workload_config(L, T, B, world):
    tokens_per_step = B * T
    return {"workload": f"K2: synthetic {L} b reads, T={T} tokens (ids+[SEP]), batch {B} per GPU, random-init ChimeraLM",
            "read_len": L, "batch_per_gpu": B, "parallelism": f"dp{world} (reads sharded, final all_gather of labels)",
            "l2": f"per-step activation working set ~{tokens_per_step * 256 * 12 / 1e6:.0f} MB >> 126 MB L2; {N_DISTINCT} distinct input batches rotate"}


def cpu_reference_path(sd, cfg, seqs_ascii: np.ndarray, batch: int = 12):
    """Reference-equivalent CPU predict on `seqs_ascii` [n, L] (chimeralm/__main__.py:248-319
    pipeline restated): per-read Python tokeniser -> collate (left pad) -> fp32 forward -> argmax."""
    from oracle import hyena_oracle, tokenizer_oracle

    n = seqs_ascii.shape[0]
    t0 = time.perf_counter()
    rows = [tokenizer_oracle.encode(bytes(r).decode(), max_length=32769, add_cls=False) for r in seqs_ascii]
    t_tok = time.perf_counter() - t0
    labels = []
    t1 = time.perf_counter()
    for i in range(0, n, batch):
        ids = torch.tensor(tokenizer_oracle.collate(rows[i:i + batch], padding_side="left"), dtype=torch.int64)
        labels += hyena_oracle.predict_labels(sd, ids, cfg).tolist()
    t_fwd = time.perf_counter() - t1
    return n / (t_tok + t_fwd), t_tok, t_fwd, labels


def run_reference(args, sd, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_per_step = 12
    reads = synth_reads(n_per_step * max(1, min(args.steps + args.warmup, 4)), READ_LEN, SEED)
    for w in range(args.warmup):
        cpu_reference_path(sd, cfg, reads[:n_per_step])
    t0 = time.perf_counter()
    for s in range(args.steps):
        off = (s % (reads.shape[0] // n_per_step)) * n_per_step
        cpu_reference_path(sd, cfg, reads[off:off + n_per_step])
    dt = time.perf_counter() - t0
    val = n_per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": "predict_reads_per_s", "value": val, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.read_len, args.read_len + 1, args.batch, max(1, args.gpus)),
        "bases_per_s": val * READ_LEN,
        "cpu_baseline": {"value": val, "unit": "reads/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"each step = {n_per_step} reads x {READ_LEN} b of the same synthetic workload through the CPU predict "
                                   f"path (python tokeniser, collate, fp32 eager forward, argmax), batch 12 (the reference CLI default); "
                                   f"{args.steps} steps; rank 0 only"},
        "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--read-len", type=int, default=READ_LEN)
    ap.add_argument("--cpu-sample", type=int, default=12, help="reads in the CPU baseline sample (0 = skip)")
    args = ap.parse_args()

    from chimeralm_b200.config import DEFAULT_CONFIG as cfg
    from chimeralm_b200.weights import make_state_dict

    sd = make_state_dict(0)
    if args.impl == "reference":
        run_reference(args, sd, cfg)
        return

    from chimeralm_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    sampler = ClockSampler(local_rank) if rank == 0 else None   # started now, rows outside the timed region are dropped
    B, L = args.batch, args.read_len
    T = L + 1
    eng = Engine(sd, device=local_rank, max_batch=B, max_tokens=T)

    # synthetic reads: each rank owns its own shard (weak scaling; reads are independent)
    reads = synth_reads(N_DISTINCT * B, L, SEED + rank)
    offsets_h = torch.arange(0, (B + 1) * L, L, dtype=torch.int64)
    dev_batches = [torch.from_numpy(reads[i * B:(i + 1) * B].reshape(-1).copy()).to(dev) for i in range(N_DISTINCT)]
    host_batches = [torch.from_numpy(reads[i * B:(i + 1) * B].reshape(-1).copy()).pin_memory() for i in range(N_DISTINCT)]
    offsets_d = offsets_h.to(dev)
    offsets_p = offsets_h.pin_memory()
    enc = dict(add_cls=False, add_sep=True, pad_left=True, max_bases=32768)

    def step_resident(i):
        ids, _ = eng.encode(dev_batches[i % N_DISTINCT], offsets_d, T, **enc)
        return eng.forward(ids, return_labels=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (`value`)
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    if sampler:
        sampler.mark_start()
    eng.profile_reset()
    eng.profile(True)
    launches0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    all_labels = []
    ev0.record()
    for i in range(args.steps):
        _, labels = step_resident(i)
        all_labels.append(labels)
    my_labels = torch.cat(all_labels)
    if dist is not None:  # the path's only exchange: final gather of predictions
        gathered = [torch.empty_like(my_labels) for _ in range(world)]
        dist.all_gather(gathered, my_labels)
    ev1.record()
    barrier()
    if sampler:
        sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - launches0
    prof = eng.profile_read()
    eng.profile(False)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())

    # ---------------- end to end through the C-ABI host entry point (`e2e`)
    lo = torch.empty(B, 2, dtype=torch.float32).pin_memory()
    la = torch.empty(B, dtype=torch.uint8).pin_memory()
    for i in range(args.warmup):
        eng.predict_host(host_batches[i % N_DISTINCT], offsets_p, T, logits_out=lo, labels_out=la, **enc)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        eng.predict_host(host_batches[i % N_DISTINCT], offsets_p, T, logits_out=lo, labels_out=la, **enc)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    reads_per_s = world * B * args.steps / (ms_max / 1e3)
    tokens_per_step = B * T
    pk = peaks()
    # dominant kernel class by device time
    dom = max(prof.items(), key=lambda kv: kv[1][0])
    dom_name, (dom_ms, dom_n) = dom
    total_prof_ms = sum(v[0] for v in prof.values())
    per_launch_s = dom_ms / dom_n / 1e3
    # algorithmic work per token per launch
    flop_per_tok = {"gemm_in_proj": 2 * 256 * 768, "gemm_out_proj": 2 * 256 * 256, "gemm_fc1": 2 * 256 * 1024,
                    "gemm_fc2": 2 * 1024 * 256, "gemm_score": 2 * 256 * 256 + 2 * 256,
                    "block_in": 2 * 256 * 768, "block_mlp": 2 * 256 * 256 + 2 * 256 * 1024 + 2 * 1024 * 256}
    bytes_per_tok = {"longconv": CONV_BYTES_TOK_LAYER, "layernorm": 1024 + 512, "shortconv_gate": 1536 + 1024,
                     "transpose": 1024, "pool": 516, "embed": 1 + 1024 + 512, "encode": 2, "head": 0}
    traffic = {}
    tp = ROOT / "profiles" / "r1_traffic.json"
    if tp.exists() and B == BATCH and L == READ_LEN:
        traffic = {k: v for k, v in json.loads(tp.read_text()).items() if not k.startswith("_")}

    conv_variant = eng.longconv_variant(T)
    if conv_variant == "fft_tensor_core":
        # Monarch FFT on tcgen05 (csrc/longconv_tc.cuh): per item (one channel of two reads) 16 MMAs 128x128x16 + 2 x 16 MMAs
        # 128x256x16 + 32 MMAs 128xN7x16 (N7 = 80 when T > 8192, else 64); 256 * ceil(B / 2) items per launch.
        n7 = 80 if T > 8192 else 64
        item_flop = 2 * 16 * (16 * 128 * 128 + 2 * 16 * 128 * 256 + 32 * 128 * n7)
        flop_per_tok["longconv"] = item_flop * 256 * ((B + 1) // 2) / tokens_per_step

    def roofline_of(name, ms, n):
        sec = ms / n / 1e3
        if name in flop_per_tok:
            ach = flop_per_tok[name] * tokens_per_step / sec / 1e12
            r = {"bound": "tensor", "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"]}
        else:
            ach = bytes_per_tok.get(name, 0) * tokens_per_step / sec / 1e9
            r = {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"]}
        r.update({"kernel": name, "traffic": traffic.get(name), "avg_launch_ms": ms / n, "share_of_step": ms / total_prof_ms})
        return r

    roof = roofline_of(dom_name, dom_ms, dom_n)
    roof["peak_source"] = pk["src"]
    if dom_name == "longconv" and conv_variant != "fft_tensor_core":
        roof["note"] = ("HBM roofline per SURVEY 8(d)'s compulsory-traffic model (1536 B/token/layer); the kernel itself is bound by "
                        "fp32 FFT instruction issue (ncu: issue-active ~50%, DRAM ~9%), see profiles/r1_v6_longconv_fast.txt")
    rooflines = {k: roofline_of(k, v[0], v[1]) for k, v in prof.items() if k in ("longconv", "block_mlp", "block_in", "gemm_score")}
    dense_frac = F_TOK * (reads_per_s / world) * T / 1e12 / pk["tflops"]

    cpu = None
    if args.cpu_sample > 0:
        torch.set_num_threads(os.cpu_count() or 1)
        sample = synth_reads(args.cpu_sample, L, SEED)
        v, t_tok, t_fwd, cpu_labels = cpu_reference_path(sd, cfg, sample)
        cpu = {"value": v, "unit": "reads/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{args.cpu_sample} reads x {L} b, batch 12; tokenise {t_tok:.2f}s + forward {t_fwd:.2f}s",
               "torch": torch.__version__}

    line = {
        "metric": "predict_reads_per_s", "value": reads_per_s, "unit": "reads/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(L, T, B, world),
        "dtype_note": "bf16 GEMM operands, fp16 operands in the tensor-core FFT conv, fp32 accumulation / residual / pooling / head",
        "bases_per_s": reads_per_s * L,
        "tokens_per_s": reads_per_s * T,
        "dense_tensor_frac_of_peak": dense_frac,
        "e2e": {"value": world * B * args.steps / e2e_s, "unit": "reads/s", "h2d_bytes_per_step": B * L + (B + 1) * 8,
                "d2h_bytes_per_step": B * 2 * 4 + B, "api": "clm_predict_host (C-ABI, pinned host buffers)"},
        "gpu_launches": int(launches),
        "roofline": roof,
        "rooflines_top_kernels": rooflines,
        "longconv_kernel": conv_variant,
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])},
        "cpu_baseline": cpu,
        "clocks": clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
