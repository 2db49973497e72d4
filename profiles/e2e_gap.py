"""Is `e2e` (clm_predict_host_submit/wait from pinned host memory) really slower than the device-resident step, or is the
difference clock drift between two passes of a bench?  Alternates the two loops (A B A B ...) on the K2 shape and prints
each pass's reads/s next to the SM clock sampled during it.
    python profiles/e2e_gap.py [rounds] [steps]"""
import subprocess
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from chimeralm_b200 import synth  # noqa: E402
from chimeralm_b200.engine import Engine  # noqa: E402
from chimeralm_b200.weights import make_state_dict  # noqa: E402


def sm_clock():
    r = subprocess.run(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"], capture_output=True, text=True)
    return r.stdout.strip()


def main():
    rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    B, L = 32, 8192
    T = L + 1
    eng = Engine(make_state_dict(0), device=0, max_batch=B, max_tokens=T)
    reads = synth.uniform_reads(8 * B, L, synth.K2_SEED)
    offs = torch.arange(0, (B + 1) * L, L, dtype=torch.int64)
    dev = [torch.from_numpy(reads[i * B:(i + 1) * B].reshape(-1).copy()).cuda() for i in range(8)]
    host = [torch.from_numpy(reads[i * B:(i + 1) * B].reshape(-1).copy()).pin_memory() for i in range(8)]
    offs_d, offs_p = offs.cuda(), offs.pin_memory()

    def resident():
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(steps):
            ids, _ = eng.encode(dev[i % 8], offs_d, T, **bench.ENC)
            eng.forward(ids, return_labels=True)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / 1e3

    def e2e():
        t0 = time.perf_counter()
        bench.e2e_pipelined(eng, [(host[i % 8], offs_p, T, B) for i in range(steps)], B)
        return time.perf_counter() - t0

    def e2e_sync():
        lo = torch.empty(B, 2).pin_memory()
        t0 = time.perf_counter()
        for i in range(steps):
            eng.predict_host(host[i % 8], offs_p, T, logits_out=lo, **bench.ENC)
        return time.perf_counter() - t0

    for _ in range(10):
        resident()
    for r in range(rounds):
        for name, fn in (("resident", resident), ("e2e pipelined", e2e), ("e2e one at a time", e2e_sync)):
            s = fn()
            print(f"round {r} {name:18s} {B * steps / s:9.1f} reads/s  {s / steps * 1e3:.3f} ms/step   clocks/power after: {sm_clock()}", flush=True)
    eng.close()


if __name__ == "__main__":
    main()
