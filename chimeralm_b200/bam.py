"""Self-contained BAM reading/writing for the predict and filter commands (no pysam/htslib).

Mirrors `chimeralm/data/bam.py:21-38` (reference): `is_chimeric` keeps records that are
mapped, carry an `SA` aux tag and are neither secondary nor supplementary; `parse_bam_file`
yields `{"id": query_name, "seq": query_sequence}` in file order (the sequence as stored in
the BAM, i.e. already reverse-complemented for reverse-strand records, like pysam's
`query_sequence`).

BGZF is a series of gzip members, which `gzip`/`zlib` read back to back; records follow the
SAM/BAM spec section 4.2.
"""

from __future__ import annotations

import gzip
import struct
import zlib
from collections.abc import Iterator
from dataclasses import dataclass
from pathlib import Path

import numpy as np

FLAG_UNMAPPED, FLAG_SECONDARY, FLAG_SUPPLEMENTARY = 0x4, 0x100, 0x800
_NIB = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)
_AUX_FIXED = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4}


@dataclass
class BamRecord:
    ref_id: int
    pos: int
    flag: int
    name: str
    l_seq: int
    raw: bytes          # the whole record after the block_size field
    seq_off: int        # offset of the 4-bit sequence inside raw
    aux_off: int        # offset of the aux area inside raw

    @property
    def is_unmapped(self) -> bool:
        return bool(self.flag & FLAG_UNMAPPED)

    @property
    def is_secondary(self) -> bool:
        return bool(self.flag & FLAG_SECONDARY)

    @property
    def is_supplementary(self) -> bool:
        return bool(self.flag & FLAG_SUPPLEMENTARY)

    def sequence_bytes(self) -> np.ndarray:
        """ASCII bases (uint8) decoded from the 4-bit packing."""
        packed = np.frombuffer(self.raw, dtype=np.uint8, count=(self.l_seq + 1) // 2, offset=self.seq_off)
        out = np.empty(packed.size * 2, dtype=np.uint8)
        out[0::2] = _NIB[packed >> 4]
        out[1::2] = _NIB[packed & 15]
        return out[: self.l_seq]

    @property
    def query_sequence(self) -> str:
        return self.sequence_bytes().tobytes().decode("ascii")

    def has_tag(self, tag: str) -> bool:
        want = tag.encode()
        raw, p, n = self.raw, self.aux_off, len(self.raw)
        while p + 3 <= n:
            t, ty = raw[p : p + 2], raw[p + 2 : p + 3]
            if t == want:
                return True
            p += 3
            if ty in _AUX_FIXED:
                p += _AUX_FIXED[ty]
            elif ty in (b"Z", b"H"):
                p = raw.index(b"\0", p) + 1
            elif ty == b"B":
                sub, cnt = raw[p : p + 1], struct.unpack_from("<i", raw, p + 1)[0]
                p += 5 + cnt * _AUX_FIXED[sub]
            else:
                raise ValueError(f"bad aux type {ty!r} in record {self.name}")
        return False


def is_chimeric(read: BamRecord) -> bool:
    """Reference `is_chimeric` (chimeralm/data/bam.py:21-23)."""
    return not read.is_unmapped and read.has_tag("SA") and not read.is_secondary and not read.is_supplementary


class BamReader:
    """Sequential BAM reader: header text, reference list, then records."""

    def __init__(self, path: str | Path):
        self.f = gzip.open(str(path), "rb")
        if self.f.read(4) != b"BAM\1":
            raise ValueError(f"{path}: not a BAM file")
        (l_text,) = struct.unpack("<i", self.f.read(4))
        self.header_text = self.f.read(l_text)
        (n_ref,) = struct.unpack("<i", self.f.read(4))
        self.references = []
        for _ in range(n_ref):
            (l_name,) = struct.unpack("<i", self.f.read(4))
            name = self.f.read(l_name)[:-1].decode()
            (l_ref,) = struct.unpack("<i", self.f.read(4))
            self.references.append((name, l_ref))

    def header_bytes(self) -> bytes:
        out = [b"BAM\1", struct.pack("<i", len(self.header_text)), self.header_text, struct.pack("<i", len(self.references))]
        for name, l_ref in self.references:
            nm = name.encode() + b"\0"
            out += [struct.pack("<i", len(nm)), nm, struct.pack("<i", l_ref)]
        return b"".join(out)

    def __iter__(self) -> Iterator[BamRecord]:
        f = self.f
        while True:
            head = f.read(4)
            if len(head) < 4:
                return
            (block_size,) = struct.unpack("<i", head)
            raw = f.read(block_size)
            if len(raw) < block_size:
                raise ValueError("truncated BAM record")
            yield record_from_raw(raw)

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def record_from_raw(raw: bytes) -> BamRecord:
    """Parse the fixed part of one BAM record (everything after its block_size field)."""
    ref_id, pos, l_read_name, _mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", raw, 0)
    name = raw[32 : 32 + l_read_name - 1].decode("ascii", "replace")
    seq_off = 32 + l_read_name + 4 * n_cigar
    aux_off = seq_off + (l_seq + 1) // 2 + l_seq
    return BamRecord(ref_id, pos, flag, name, l_seq, raw, seq_off, aux_off)


def samtools_sort_key(r: BamRecord):
    """`samtools sort` coordinate order: refID as unsigned (unplaced reads last), position, forward strand first."""
    return ((r.ref_id & 0xFFFFFFFF), r.pos + 1, (r.flag >> 4) & 1)


def sorted_records_external(records, run_bytes: int = 256 << 20, tmpdir=None) -> Iterator[BamRecord]:
    """Coordinate-sort a stream of records with bounded memory, the way `samtools sort` (which the reference calls
    through `pysam.sort`, chimeralm/__main__.py:148) does: sorted runs of at most `run_bytes` of record data are spilled
    to temporary files and merged k-way; ties keep input order (every record carries its arrival number).  A stream
    that fits one run never touches the disk."""
    import heapq
    import tempfile

    runs, cur, cur_bytes, seq = [], [], 0, 0

    def spill():
        nonlocal cur, cur_bytes
        cur.sort(key=lambda t: t[0])
        f = tempfile.TemporaryFile(dir=tmpdir, prefix="clm_sort_run_")
        for (_k, n, raw) in cur:
            f.write(struct.pack("<qi", n, len(raw)))
            f.write(raw)
        f.seek(0)
        runs.append(f)
        cur, cur_bytes = [], 0

    for r in records:
        cur.append((samtools_sort_key(r) + (seq,), seq, r.raw))
        cur_bytes += len(r.raw) + 64
        seq += 1
        if cur_bytes >= run_bytes:
            spill()
    if not runs:
        cur.sort(key=lambda t: t[0])
        for (_k, _n, raw) in cur:
            yield record_from_raw(raw)
        return
    if cur:
        spill()

    def read_run(f):
        while True:
            head = f.read(12)
            if len(head) < 12:
                f.close()
                return
            n, size = struct.unpack("<qi", head)
            rec = record_from_raw(f.read(size))
            yield samtools_sort_key(rec) + (n,), rec

    for _k, rec in heapq.merge(*(read_run(f) for f in runs), key=lambda t: t[0]):
        yield rec


def coordinate_sorted_header(bam: "BamReader") -> bytes:
    """Header of `bam` with `@HD ... SO:coordinate` (what `samtools sort` writes; an index needs a sorted file)."""
    lines = bam.header_text.rstrip(b"\0").split(b"\n")
    if lines and lines[0].startswith(b"@HD"):
        fields = [f for f in lines[0].split(b"\t") if not f.startswith((b"SO:", b"GO:"))]
        lines[0] = b"\t".join(fields + [b"SO:coordinate"])
    else:
        lines.insert(0, b"@HD\tVN:1.6\tSO:coordinate")
    text = b"\n".join(lines)
    if not text.endswith(b"\n"):
        text += b"\n"
    out = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", len(bam.references))]
    for name, l_ref in bam.references:
        nm = name.encode() + b"\0"
        out += [struct.pack("<i", len(nm)), nm, struct.pack("<i", l_ref)]
    return b"".join(out)


def parse_bam_file(file_path: str | Path) -> Iterator[dict]:
    """Reference `parse_bam_file` (chimeralm/data/bam.py:26-38)."""
    with BamReader(file_path) as bam:
        for read in bam:
            if is_chimeric(read):
                yield {"id": read.name, "seq": read.query_sequence}


def parse_bam_file_bytes(file_path: str | Path):
    """Same records as `parse_bam_file` but bases stay uint8 arrays (feeds the CUDA encoder)."""
    with BamReader(file_path) as bam:
        for read in bam:
            if is_chimeric(read):
                yield read.name, read.sequence_bytes()


def make_record(name: str, seq: str | bytes, flag: int = 0, sa_tag: bool = True, ref_id: int = 0, pos: int = 0,
                extra_aux: bytes = b"") -> BamRecord:
    """A synthetic aligned record (one `<len>M` CIGAR op, quality 0xff) for tests and benches."""
    sb = seq.encode() if isinstance(seq, str) else bytes(seq)
    n = len(sb)
    codes = np.zeros(256, np.uint8) + 15
    codes[_NIB] = np.arange(16, dtype=np.uint8)
    nib = codes[np.frombuffer(sb, np.uint8)]
    if n & 1:
        nib = np.append(nib, np.uint8(0))
    packed = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8).tobytes()
    nm = name.encode() + b"\0"
    aux = extra_aux + (b"SAZchr1,100,+,50M,60,0;\0" if sa_tag else b"") + b"NMi" + struct.pack("<i", 0)
    raw = (struct.pack("<iiBBHHHiiii", ref_id, pos, len(nm), 60, 4680, 1, flag, n, -1, -1, 0) + nm +
           struct.pack("<I", (n << 4) | 0) + packed + b"\xff" * n + aux)
    seq_off = 32 + len(nm) + 4
    return BamRecord(ref_id, pos, flag, name, n, raw, seq_off, seq_off + (n + 1) // 2 + n)


def minimal_header(references=(("chr1", 1_000_000),)) -> bytes:
    text = b"@HD\tVN:1.6\tSO:unsorted\n" + b"".join(b"@SQ\tSN:%s\tLN:%d\n" % (n.encode(), l) for n, l in references)
    out = [b"BAM\1", struct.pack("<i", len(text)), text, struct.pack("<i", len(references))]
    for n, l in references:
        nm = n.encode() + b"\0"
        out += [struct.pack("<i", len(nm)), nm, struct.pack("<i", l)]
    return b"".join(out)


# ---------------------------------------------------------------------------------- writing
_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_block(data: bytes, level: int = 6) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


class BamWriter:
    """Minimal BGZF BAM writer (records are copied byte-for-byte from the input)."""

    def __init__(self, path: str | Path, header_bytes: bytes):
        self.f = open(str(path), "wb")
        self.buf = bytearray(header_bytes)

    def write(self, rec: BamRecord) -> None:
        self.buf += struct.pack("<i", len(rec.raw)) + rec.raw
        while len(self.buf) >= 0xFF00:
            self.f.write(_bgzf_block(bytes(self.buf[:0xFF00])))
            del self.buf[:0xFF00]

    def close(self) -> None:
        if self.buf:
            self.f.write(_bgzf_block(bytes(self.buf)))
        self.f.write(_BGZF_EOF)
        self.f.close()
