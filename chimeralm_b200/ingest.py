"""Python face of the native BAM ingest (`clm_bam_*` in include/chimeralm_b200.h).

`read_bam_flat` is what the predict data module uses in place of the reference's per-read
generator (`parse_bam_file`, chimeralm/data/bam.py:26-38): the kept reads come back as ONE
uint8 array of ASCII bases plus int64 offsets and the query names, in file order, already cut
to `max_bases` — the layout `clm_encode_batch` consumes, so no per-read Python object is made
between the BAM and the GPU.
"""

from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import _lib
from ._lib import ChimeraLMNativeError

NAME_STRIDE = 256  # BAM l_read_name is one byte: names are at most 254 chars + NUL


class NativeBamReader:
    """Streaming reader: `next_block()` fills caller-sized buffers with the next kept reads."""

    def __init__(self, path: str | Path, n_threads: int = 0, chunk_bytes: int | None = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.clm_bam_open(str(path).encode(), int(n_threads), C.byref(h))
        if rc < 0:
            msg = self.lib.clm_bam_error(None)
            raise ValueError(f"clm_bam_open failed (status {rc}): {msg.decode() if msg else '?'}")
        self.h = h
        if chunk_bytes is not None and self.lib.clm_bam_set_chunk_bytes(self.h, int(chunk_bytes)) < 0:
            raise ValueError(f"bad chunk size {chunk_bytes}")

    def next_block(self, max_reads: int, max_bases: int, bases: np.ndarray, offsets: np.ndarray,
                   names: np.ndarray | None, chimeric_only: bool = True) -> int:
        """Decode up to `max_reads` reads into `bases` (uint8), `offsets` (int64[max_reads+1]) and
        `names` (uint8[max_reads, NAME_STRIDE] or None).  Returns the number of reads, 0 at EOF."""
        assert bases.dtype == np.uint8 and offsets.dtype == np.int64 and offsets.size >= max_reads + 1
        assert bases.flags.c_contiguous and offsets.flags.c_contiguous
        if names is not None:
            assert names.dtype == np.uint8 and names.flags.c_contiguous and names.shape[0] >= max_reads
        n = self.lib.clm_bam_next(self.h, max_reads, max_bases, int(chimeric_only), bases.ctypes.data, bases.size,
                                  offsets.ctypes.data, names.ctypes.data if names is not None else None,
                                  names.shape[1] if names is not None else 0)
        if n < 0:
            msg = self.lib.clm_bam_error(self.h)
            raise ChimeraLMNativeError(f"clm_bam_next failed (status {n}): {msg.decode() if msg else '?'}")
        return int(n)

    def set_shard(self, rank: int, world: int) -> None:
        """Only kept reads with running index i % world == rank are returned from now on."""
        if self.lib.clm_bam_set_shard(self.h, int(rank), int(world)) < 0:
            raise ValueError(f"bad shard {rank}/{world}")

    @property
    def records_seen(self) -> int:
        return int(self.lib.clm_bam_records_seen(self.h))

    def close(self) -> None:
        if self.h:
            self.lib.clm_bam_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _names_from_rows(rows: np.ndarray) -> list[str]:
    out = []
    for r in rows:
        b = r.tobytes()
        out.append(b[: b.index(0)].decode("ascii", "replace"))
    return out


def read_bam_flat(path: str | Path, max_bases: int, max_reads: int | None = None, chimeric_only: bool = True,
                  n_threads: int = 0, block_reads: int = 8192, block_bytes: int = 64 << 20, chunk_bytes: int | None = None):
    """All kept reads of a BAM as (names, flat uint8 bases, int64 offsets[n+1]), file order."""
    block_bytes = max(block_bytes, max_bases)
    names: list[str] = []
    chunks: list[np.ndarray] = []
    lens: list[np.ndarray] = []
    bases = np.empty(block_bytes, np.uint8)
    offs = np.empty(block_reads + 1, np.int64)
    nm = np.empty((block_reads, NAME_STRIDE), np.uint8)
    with NativeBamReader(path, n_threads, chunk_bytes) as rd:
        while max_reads is None or len(names) < max_reads:
            want = block_reads if max_reads is None else min(block_reads, max_reads - len(names))
            n = rd.next_block(want, max_bases, bases, offs, nm, chimeric_only)
            if n == 0:
                break
            names += _names_from_rows(nm[:n])
            chunks.append(bases[: offs[n]].copy())
            lens.append(np.diff(offs[: n + 1]))
    offsets = np.zeros(len(names) + 1, np.int64)
    if lens:
        np.cumsum(np.concatenate(lens), out=offsets[1:])
    flat = np.concatenate(chunks) if chunks else np.zeros(0, np.uint8)
    return names, flat, offsets
