"""BAI index for a coordinate-sorted BAM (no htslib): the last step of `chimeralm filter`.

Replaces `pysam.index(sorted_output_path)` (`chimeralm/__main__.py:150-151`, reference).  The index follows the SAM/BAM
specification section 5.2 and is built the way htslib's `hts_idx_push` / `hts_idx_finish` build it, so that the result can
be compared chunk for chunk with an index samtools wrote (`tests/golden/test_chimric_reads.bam.bai` is the reference's
own fixture for `tests/data/test_chimric_reads.bam`):

* a chunk = a run of consecutive records falling into the same bin, `[virtual offset of the first record, virtual offset
  after the last)`;
* after the last record, a bin whose chunks span less than one BGZF block distance (64 KiB of file offset) is merged
  into its parent when the parent exists, then adjacent chunks that touch the same BGZF block are joined;
* the linear index holds, per 16 kb window, the offset of the first record overlapping it; empty windows inherit the
  next window's offset;
* every reference with records gets the pseudo-bin 37450 (`[first, last)` offsets and mapped / unmapped counts), and
  the number of unplaced reads closes the file.

Virtual offsets (`compressed block start << 16 | offset inside the block`) need the BGZF block boundaries, so the file is
walked block by block here rather than through `gzip`.
"""

from __future__ import annotations

import struct
import zlib
from pathlib import Path

META_BIN = 37450
_MIN_MARKER_DIST = 0x10000
_CIGAR_CONSUMES_REF = (1, 0, 1, 1, 0, 0, 0, 1, 1)   # M I D N S H P = X


class BgzfVirtualReader:
    """Sequential BGZF reader that reports htslib-compatible virtual offsets (`bgzf_tell`)."""

    def __init__(self, path: str | Path):
        self.f = open(str(path), "rb")
        self.block_address = 0      # file offset of the current block
        self.data = b""
        self.pos = 0

    def _load_block(self) -> bool:
        """Loads the next NON-EMPTY block.  Empty blocks (the EOF marker) and the end of the file leave the position
        where the last data byte put it, which is what `bgzf_tell` reports to `hts_idx_finish` at the end of a BAM."""
        while True:
            address = self.f.tell()
            head = self.f.read(12)
            if len(head) < 12:
                return False
            if head[:4] != b"\x1f\x8b\x08\x04":
                raise ValueError("not a BGZF block")
            (xlen,) = struct.unpack_from("<H", head, 10)
            extra = self.f.read(xlen)
            bsize, p = None, 0
            while p + 4 <= xlen:
                si1, si2, slen = extra[p], extra[p + 1], struct.unpack_from("<H", extra, p + 2)[0]
                if si1 == 66 and si2 == 67 and slen == 2:
                    bsize = struct.unpack_from("<H", extra, p + 4)[0]
                p += 4 + slen
            if bsize is None:
                raise ValueError("BGZF block without a BC subfield")
            cdata = self.f.read(bsize - xlen - 19)
            crc, isize = struct.unpack("<II", self.f.read(8))
            data = zlib.decompress(cdata, -15) if isize else b""
            if len(data) != isize or (zlib.crc32(data) & 0xFFFFFFFF) != crc:
                raise ValueError("corrupt BGZF block")
            if data:
                self.block_address, self.data, self.pos = address, data, 0
                return True

    def read(self, n: int) -> bytes:
        out = []
        while n > 0:
            if self.pos >= len(self.data) and not self._load_block():
                break
            take = self.data[self.pos : self.pos + n]
            self.pos += len(take)
            n -= len(take)
            out.append(take)
        # htslib: a read that ends exactly at the end of a block leaves the position at the START of the next one
        if self.pos == len(self.data) and self.data:
            self.block_address = self.f.tell()
            self.data, self.pos = b"", 0
        return b"".join(out)

    def tell(self) -> int:
        return (self.block_address << 16) | self.pos

    def seek(self, voffset: int) -> None:
        self.f.seek(voffset >> 16)
        self.data, self.pos = b"", 0
        self.block_address = voffset >> 16
        if voffset & 0xFFFF:
            if not self._load_block():
                raise ValueError("virtual offset past the end of the file")
            self.pos = voffset & 0xFFFF

    def close(self) -> None:
        self.f.close()


def reg2bin(beg: int, end: int) -> int:
    """SAM spec 5.3 (`hts_reg2bin` with min_shift 14, 5 levels); `end` is exclusive."""
    end -= 1
    if beg >> 14 == end >> 14:
        return 4681 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return 585 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return 73 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return 9 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return 1 + (beg >> 26)
    return 0


def _first_bin_of_level(level: int) -> int:
    return ((1 << 3 * level) - 1) // 7


def _bin_parent(b: int) -> int:
    return (b - 1) >> 3


class _IndexBuilder:
    def __init__(self, n_ref: int, offset0: int):
        self.bins: list[dict[int, list[list[int]]] | None] = [None] * n_ref
        self.lidx: list[list[int]] = [[] for _ in range(n_ref)]
        self.n_no_coor = 0
        self.last_tid = self.last_bin = self.save_tid = self.save_bin = None
        self.last_coor = -1
        self.last_off = self.save_off = self.off_beg = offset0
        self.n_mapped = self.n_unmapped = 0

    def _insert(self, tid, b, beg, end):
        if tid is None or tid < 0:
            return
        self.bins[tid].setdefault(b, []).append([beg, end])

    def push(self, tid: int, beg: int, end: int, offset: int, is_mapped: bool) -> None:
        if tid < 0:
            beg, end = -1, 0
        if self.last_tid != tid:
            if tid >= 0 and self.n_no_coor:
                raise ValueError("unplaced reads are not in a single block at the end of the BAM")
            if tid >= 0 and self.bins[tid] is not None:
                raise ValueError("BAM is not coordinate-sorted: reference blocks are not contiguous")
            new_chrom = True
            self.last_tid = tid
            self.last_bin = None
        else:
            new_chrom = False
            if tid >= 0 and self.last_coor > beg:
                raise ValueError("BAM is not coordinate-sorted: positions out of order")
        if tid >= 0:
            if self.bins[tid] is None:
                self.bins[tid] = {}
            beg, end = max(beg, 0), (end if end > 0 else 1)
            lo, hi = beg >> 14, (end - 1) >> 14
            l = self.lidx[tid]
            if len(l) < hi + 1:
                l.extend([-1] * (hi + 1 - len(l)))
            for i in range(lo, hi + 1):
                if l[i] == -1:
                    l[i] = self.last_off
            b = reg2bin(beg, end)
        else:
            self.n_no_coor += 1
            b = 4680                         # hts_reg2bin(-1, 0)
        if self.last_bin != b:
            if self.save_bin is not None:
                self._insert(self.save_tid, self.save_bin, self.save_off, self.last_off)
            if new_chrom and self.save_bin is not None:
                self._insert(self.save_tid, META_BIN, self.off_beg, self.last_off)
                self._insert(self.save_tid, META_BIN, self.n_mapped, self.n_unmapped)
                self.n_mapped = self.n_unmapped = 0
                self.off_beg = self.last_off
            self.save_off = self.last_off
            self.save_bin = self.last_bin = b
            self.save_tid = tid
        if is_mapped:
            self.n_mapped += 1
        else:
            self.n_unmapped += 1
        self.last_off = offset
        self.last_coor = beg

    def finish(self, final_offset: int) -> None:
        if self.save_tid is not None and self.save_tid >= 0:
            self._insert(self.save_tid, self.save_bin, self.save_off, final_offset)
            self._insert(self.save_tid, META_BIN, self.off_beg, final_offset)
            self._insert(self.save_tid, META_BIN, self.n_mapped, self.n_unmapped)
        for tid, l in enumerate(self.lidx):
            for i in range(len(l) - 2, -1, -1):
                if l[i] == -1:
                    l[i] = l[i + 1]
            if self.bins[tid] is not None:
                self._compress(self.bins[tid])

    @staticmethod
    def _compress(bins: dict[int, list[list[int]]]) -> None:
        for level in range(5, 0, -1):
            start = _first_bin_of_level(level)
            for b in sorted(k for k in bins if start <= k < META_BIN):
                chunks = bins[b]
                if level < 5 and len(chunks) > 1:
                    chunks.sort()
                if (chunks[-1][1] >> 16) - (chunks[0][0] >> 16) < _MIN_MARKER_DIST:
                    parent = bins.get(_bin_parent(b))
                    if parent is None:
                        continue
                    parent.extend(chunks)
                    del bins[b]
        if 0 in bins:
            bins[0].sort()
        for b, chunks in bins.items():
            if b >= META_BIN:
                continue
            merged = [chunks[0]]
            for c in chunks[1:]:
                if merged[-1][1] >> 16 >= c[0] >> 16:
                    if merged[-1][1] < c[1]:
                        merged[-1][1] = c[1]
                else:
                    merged.append(c)
            bins[b] = merged


def _reference_end(raw: bytes, pos: int, flag: int, l_read_name: int, n_cigar: int) -> int:
    """`bam_endpos`: pos + reference bases consumed by the CIGAR; pos + 1 for unmapped / CIGAR-less records."""
    if flag & 0x4 or n_cigar == 0:
        return pos + 1
    rlen = 0
    for (op,) in struct.iter_unpack("<I", raw[32 + l_read_name : 32 + l_read_name + 4 * n_cigar]):
        o = op & 15
        if o < 9 and _CIGAR_CONSUMES_REF[o]:
            rlen += op >> 4
    return pos + (rlen if rlen else 1)


def build_index(bam_path: str | Path) -> dict:
    """Index of a coordinate-sorted BAM as `{"refs": [{"bins": {bin: [[beg, end], ...]}, "linear": [...]}, ...],
    "n_no_coor": int}` (bins include the 37450 pseudo-bin)."""
    r = BgzfVirtualReader(bam_path)
    try:
        if r.read(4) != b"BAM\1":
            raise ValueError(f"{bam_path}: not a BAM file")
        (l_text,) = struct.unpack("<i", r.read(4))
        r.read(l_text)
        (n_ref,) = struct.unpack("<i", r.read(4))
        for _ in range(n_ref):
            (l_name,) = struct.unpack("<i", r.read(4))
            r.read(l_name + 4)
        ib = _IndexBuilder(n_ref, r.tell())
        while True:
            head = r.read(4)
            if len(head) < 4:
                break
            (block_size,) = struct.unpack("<i", head)
            raw = r.read(block_size)
            if len(raw) < block_size:
                raise ValueError("truncated BAM record")
            ref_id, pos, l_read_name, _mapq, _bin, n_cigar, flag, _l_seq = struct.unpack_from("<iiBBHHHi", raw, 0)
            ib.push(ref_id, pos, _reference_end(raw, pos, flag, l_read_name, n_cigar), r.tell(), not (flag & 0x4))
        ib.finish(r.tell())
    finally:
        r.close()
    return {"refs": [{"bins": ib.bins[i] or {}, "linear": ib.lidx[i]} for i in range(n_ref)], "n_no_coor": ib.n_no_coor}


def serialize_index(idx: dict) -> bytes:
    out = [b"BAI\1", struct.pack("<i", len(idx["refs"]))]
    for ref in idx["refs"]:
        out.append(struct.pack("<i", len(ref["bins"])))
        for b in sorted(ref["bins"]):
            chunks = ref["bins"][b]
            out.append(struct.pack("<Ii", b, len(chunks)))
            out += [struct.pack("<QQ", c[0], c[1]) for c in chunks]
        out.append(struct.pack("<i", len(ref["linear"])))
        out += [struct.pack("<Q", o) for o in ref["linear"]]
    out.append(struct.pack("<Q", idx["n_no_coor"]))
    return b"".join(out)


def parse_index(data: bytes) -> dict:
    """Inverse of `serialize_index`; also reads indexes written by samtools (bins in any order)."""
    if data[:4] != b"BAI\1":
        raise ValueError("not a BAI index")
    (n_ref,) = struct.unpack_from("<i", data, 4)
    p, refs = 8, []
    for _ in range(n_ref):
        (n_bin,) = struct.unpack_from("<i", data, p)
        p += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", data, p)
            p += 8
            bins[b] = [list(struct.unpack_from("<QQ", data, p + 16 * i)) for i in range(n_chunk)]
            p += 16 * n_chunk
        (n_intv,) = struct.unpack_from("<i", data, p)
        p += 4
        linear = list(struct.unpack_from(f"<{n_intv}Q", data, p))
        p += 8 * n_intv
        refs.append({"bins": bins, "linear": linear})
    n_no_coor = struct.unpack_from("<Q", data, p)[0] if p + 8 <= len(data) else 0
    return {"refs": refs, "n_no_coor": n_no_coor}


def index_bam(bam_path: str | Path, bai_path: str | Path | None = None) -> Path:
    """`pysam.index(path)`: writes `<path>.bai` next to the BAM."""
    bai_path = Path(str(bam_path) + ".bai") if bai_path is None else Path(bai_path)
    bai_path.write_bytes(serialize_index(build_index(bam_path)))
    return bai_path


def reg2bins(beg: int, end: int) -> list[int]:
    """Bins that may hold records overlapping `[beg, end)` (SAM spec 5.3)."""
    end -= 1
    out = [0]
    for shift, first in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        out.extend(range(first + (beg >> shift), first + (end >> shift) + 1))
    return out


def query_offsets(idx: dict, tid: int, beg: int, end: int) -> list[list[int]]:
    """Chunks a reader has to scan for records of reference `tid` overlapping `[beg, end)` (bins + linear-index cut)."""
    ref = idx["refs"][tid]
    lin = ref["linear"]
    min_off = lin[min(beg >> 14, len(lin) - 1)] if lin else 0
    chunks = [c for b in reg2bins(beg, end) for c in ref["bins"].get(b, []) if c[1] > min_off]
    return sorted(chunks)


def fetch(bam_path: str | Path, idx: dict, tid: int, beg: int, end: int) -> list[tuple[str, int, int]]:
    """`(name, pos, end)` of the records of reference `tid` overlapping `[beg, end)`, found THROUGH the index (what
    `AlignmentFile.fetch(contig, beg, end)` does): only the chunks `query_offsets` returns are read."""
    out, seen = [], set()
    r = BgzfVirtualReader(bam_path)
    try:
        for c_beg, c_end in query_offsets(idx, tid, beg, end):
            r.seek(c_beg)
            while r.tell() < c_end:
                start = r.tell()
                head = r.read(4)
                if len(head) < 4:
                    break
                raw = r.read(struct.unpack("<i", head)[0])
                ref_id, pos, l_read_name, _mapq, _bin, n_cigar, flag, _l_seq = struct.unpack_from("<iiBBHHHi", raw, 0)
                if ref_id != tid or pos >= end:
                    break
                rec_end = _reference_end(raw, pos, flag, l_read_name, n_cigar)
                if rec_end > beg and start not in seen:
                    seen.add(start)
                    out.append((raw[32 : 32 + l_read_name - 1].decode("ascii", "replace"), pos, rec_end))
    finally:
        r.close()
    return out
