// Native BAM ingest for the predict path: BGZF inflate on a thread pool, record parse, the
// reference's `is_chimeric` filter and 4-bit -> ASCII base decode straight into caller-owned
// (pinned) buffers.  Replaces, for `chimeralm predict`, pysam.AlignmentFile iteration +
// `is_chimeric` + `parse_bam_file` (reference chimeralm/data/bam.py:21-38) and the per-read
// Python objects HF `Dataset.from_generator` builds from them (chimeralm/data/bam.py:129-174).
//
// File format: SAM/BAM specification sections 4.1 (BGZF) and 4.2 (records).  zlib does the
// raw-deflate work; blocks are independent, so a chunk of compressed blocks is inflated by
// n_threads workers while the caller parses the previous chunk.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <new>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <future>
#include <string>
#include <thread>
#include <vector>

#include "../../include/chimeralm_b200.h"

namespace {

inline uint32_t le16(const uint8_t* p) { return uint32_t(p[0]) | uint32_t(p[1]) << 8; }
inline uint32_t le32(const uint8_t* p) { return le16(p) | le16(p + 2) << 16; }

struct Chunk {
  std::vector<uint8_t> data;  // inflated bytes of a run of whole BGZF blocks
  bool eof = false;
  std::string err;
};

struct Block {
  size_t in_off, in_len, out_off;
  uint32_t isize, crc;
};

thread_local std::string g_open_err;

}  // namespace

struct clm_bam {
  FILE* f = nullptr;
  int n_threads = 1;
  size_t chunk_bytes = size_t(16) << 20;
  size_t ramp_bytes = size_t(1) << 20;   // the first loads are short (1, 4, 16 MiB): the first reads come out after a 1 MiB
                                         // inflate instead of a 16 MiB one (~0.1 s of start-up latency for the predict loop)
  std::string err;
  std::vector<uint8_t> cbuf;  // compressed bytes read but not yet inflated (partial block carry)
  bool file_eof = false;
  std::future<Chunk> next;    // the chunk being inflated in the background
  std::vector<uint8_t> buf;   // inflated window the parser walks
  size_t pos = 0;
  bool eof = false;
  long long n_records = 0;
  long long n_kept = 0;       // records that passed the filter (the global read index)
  int shard_rank = 0, shard_world = 1;
  std::string header_text;
  int n_ref = 0;

  Chunk load_chunk();
  bool refill();
  bool ensure(size_t need);
};

Chunk clm_bam::load_chunk() {
  Chunk out;
  if (!file_eof) {
    size_t have = cbuf.size();
    const size_t want = std::min(chunk_bytes, ramp_bytes);
    ramp_bytes = std::min(chunk_bytes, ramp_bytes * 4);
    cbuf.resize(have + want);
    size_t got = fread(cbuf.data() + have, 1, want, f);
    cbuf.resize(have + got);
    if (got < want) {
      if (ferror(f)) {
        out.err = "read error";
        return out;
      }
      file_eof = true;
    }
  }
  const uint8_t* c = cbuf.data();
  const size_t n = cbuf.size();
  std::vector<Block> blks;
  size_t p = 0, out_off = 0;
  while (p + 18 <= n) {
    if (c[p] != 0x1f || c[p + 1] != 0x8b || c[p + 2] != 8 || !(c[p + 3] & 4)) {
      out.err = "not a BGZF block (bad gzip member header)";
      return out;
    }
    const size_t xlen = le16(c + p + 10);
    if (p + 12 + xlen > n) break;
    long bsize = -1;
    for (size_t q = p + 12; q + 4 <= p + 12 + xlen;) {
      const size_t slen = le16(c + q + 2);
      if (c[q] == 'B' && c[q + 1] == 'C' && slen == 2 && q + 6 <= p + 12 + xlen) bsize = long(le16(c + q + 4)) + 1;
      q += 4 + slen;
    }
    if (bsize < long(12 + xlen + 8)) {
      out.err = "BGZF block without a valid BC subfield";
      return out;
    }
    if (p + size_t(bsize) > n) break;
    Block b;
    b.in_off = p + 12 + xlen;
    b.in_len = size_t(bsize) - 12 - xlen - 8;
    b.crc = le32(c + p + bsize - 8);
    b.isize = le32(c + p + bsize - 4);
    b.out_off = out_off;
    if (b.isize > 65536) {
      out.err = "BGZF block larger than 64 KiB";
      return out;
    }
    out_off += b.isize;
    blks.push_back(b);
    p += size_t(bsize);
  }
  if (blks.empty()) {
    if (file_eof && n == 0) {
      out.eof = true;
      return out;
    }
    if (file_eof) {
      out.err = "truncated BGZF block at end of file";
      return out;
    }
    return load_chunk();   // the read so far ends inside the first block (tiny chunk size): read more, cbuf keeps the carry
  }
  out.data.resize(out_off);
  std::atomic<size_t> cursor{0};
  std::atomic<int> bad{0};
  auto work = [&]() {
    z_stream zs;
    std::memset(&zs, 0, sizeof zs);
    if (inflateInit2(&zs, -15) != Z_OK) {
      bad = 1;
      return;
    }
    for (;;) {
      const size_t i = cursor.fetch_add(1);
      if (i >= blks.size() || bad.load()) break;
      const Block& b = blks[i];
      if (b.isize == 0) continue;
      inflateReset(&zs);
      zs.next_in = const_cast<Bytef*>(c + b.in_off);
      zs.avail_in = uInt(b.in_len);
      zs.next_out = out.data.data() + b.out_off;
      zs.avail_out = b.isize;
      const int rc = inflate(&zs, Z_FINISH);
      if (rc != Z_STREAM_END || zs.avail_out != 0) {
        bad = 2;
        break;
      }
      if (uint32_t(crc32(crc32(0L, Z_NULL, 0), out.data.data() + b.out_off, b.isize)) != b.crc) {
        bad = 3;
        break;
      }
    }
    inflateEnd(&zs);
  };
  const int nt = int(std::min<size_t>(size_t(n_threads), blks.size()));
  std::vector<std::thread> pool;
  for (int t = 1; t < nt; ++t) pool.emplace_back(work);
  work();
  for (auto& th : pool) th.join();
  if (bad.load()) {
    out.err = bad.load() == 3 ? "BGZF CRC mismatch" : "inflate failed (corrupt BGZF block)";
    out.data.clear();
    return out;
  }
  cbuf.erase(cbuf.begin(), cbuf.begin() + long(p));
  return out;
}

bool clm_bam::refill() {
  Chunk ch = next.get();
  if (!ch.err.empty()) {
    err = ch.err;
    eof = true;
    return false;
  }
  if (ch.eof) {
    eof = true;
    return false;
  }
  next = std::async(std::launch::async, [this] { return load_chunk(); });
  if (pos) {
    buf.erase(buf.begin(), buf.begin() + long(pos));
    pos = 0;
  }
  buf.insert(buf.end(), ch.data.begin(), ch.data.end());
  return true;
}

bool clm_bam::ensure(size_t need) {
  while (buf.size() - pos < need) {
    if (eof || !refill()) return false;
  }
  return true;
}

namespace {

// True when the aux area holds an `SA` tag (BAM spec 4.2.4 aux encoding).  `ok` goes false on
// a malformed aux area.
bool has_sa_tag(const uint8_t* a, size_t n, bool* ok) {
  size_t p = 0;
  while (p + 3 <= n) {
    if (a[p] == 'S' && a[p + 1] == 'A') return true;
    const uint8_t ty = a[p + 2];
    p += 3;
    switch (ty) {
      case 'A': case 'c': case 'C': p += 1; break;
      case 's': case 'S': p += 2; break;
      case 'i': case 'I': case 'f': p += 4; break;
      case 'Z': case 'H': {
        const void* z = std::memchr(a + p, 0, n > p ? n - p : 0);
        if (!z) { *ok = false; return false; }
        p = size_t(static_cast<const uint8_t*>(z) - a) + 1;
        break;
      }
      case 'B': {
        if (p + 5 > n) { *ok = false; return false; }
        const uint8_t sub = a[p];
        const size_t cnt = le32(a + p + 1);
        const size_t w = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 :
                         (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : 0;
        if (!w) { *ok = false; return false; }
        p += 5 + cnt * w;
        break;
      }
      default: *ok = false; return false;
    }
  }
  return false;
}

struct NibbleLut {
  uint16_t v[256];
  NibbleLut() {
    const char* s = "=ACMGRSVTWYHKDBN";
    for (int i = 0; i < 256; ++i) v[i] = uint16_t(uint8_t(s[i >> 4])) | uint16_t(uint8_t(s[i & 15])) << 8;
  }
};
const NibbleLut kNib;

}  // namespace

extern "C" {

int clm_bam_open(const char* path, int n_threads, clm_bam** out) {
  if (!path || !out) return CLM_ERR_INVALID;
  *out = nullptr;
  FILE* f = fopen(path, "rb");
  if (!f) {
    g_open_err = std::string("cannot open ") + path;
    return CLM_ERR_INVALID;
  }
  clm_bam* r = new (std::nothrow) clm_bam;
  if (!r) {
    fclose(f);
    return CLM_ERR_NOMEM;
  }
  r->f = f;
  if (n_threads < 1) n_threads = int(std::max(1u, std::thread::hardware_concurrency()));
  r->n_threads = n_threads;
  r->next = std::async(std::launch::async, [r] { return r->load_chunk(); });
  auto fail = [&](const std::string& m) {
    g_open_err = std::string(path) + ": " + (r->err.empty() ? m : r->err);
    clm_bam_close(r);
    return CLM_ERR_INVALID;
  };
  if (!r->ensure(12)) return fail("not a BAM file (too short)");
  if (std::memcmp(r->buf.data() + r->pos, "BAM\1", 4) != 0) return fail("not a BAM file (bad magic)");
  const size_t l_text = le32(r->buf.data() + r->pos + 4);
  if (!r->ensure(12 + l_text)) return fail("truncated BAM header");
  r->header_text.assign(reinterpret_cast<const char*>(r->buf.data() + r->pos + 8), l_text);
  r->n_ref = int(le32(r->buf.data() + r->pos + 8 + l_text));
  r->pos += 12 + l_text;
  for (int i = 0; i < r->n_ref; ++i) {
    if (!r->ensure(4)) return fail("truncated BAM reference list");
    const size_t l_name = le32(r->buf.data() + r->pos);
    if (!r->ensure(8 + l_name)) return fail("truncated BAM reference list");
    r->pos += 8 + l_name;
  }
  *out = r;
  return CLM_OK;
}

void clm_bam_close(clm_bam* r) {
  if (!r) return;
  if (r->next.valid()) r->next.wait();
  if (r->f) fclose(r->f);
  delete r;
}

const char* clm_bam_error(const clm_bam* r) { return r ? r->err.c_str() : g_open_err.c_str(); }

long long clm_bam_records_seen(const clm_bam* r) { return r ? r->n_records : 0; }

int clm_bam_set_chunk_bytes(clm_bam* r, long long bytes) {
  if (!r || bytes < 64) return CLM_ERR_INVALID;
  if (r->next.valid()) r->next.wait();   // the background load reads chunk_bytes: change it only between loads
  r->chunk_bytes = (size_t)bytes;
  return CLM_OK;
}

int clm_bam_set_shard(clm_bam* r, int rank, int world) {
  if (!r || world < 1 || rank < 0 || rank >= world) return CLM_ERR_INVALID;
  r->shard_rank = rank;
  r->shard_world = world;
  return CLM_OK;
}

long long clm_bam_next(clm_bam* r, long long max_reads, long long max_bases, int chimeric_only, uint8_t* bases,
                       long long bases_cap, int64_t* offsets, char* names, int name_stride) {
  if (!r || !bases || !offsets || max_reads < 0 || max_bases < 0 || bases_cap < 0 || (names && name_stride < 2))
    return CLM_ERR_INVALID;
  long long n_out = 0, used = 0;
  offsets[0] = 0;
  while (n_out < max_reads) {
    if (!r->ensure(4)) {
      if (!r->err.empty()) return CLM_ERR_INVALID;
      if (r->buf.size() != r->pos) {
        r->err = "truncated BAM record";
        return CLM_ERR_INVALID;
      }
      break;  // clean end of file
    }
    const size_t block_size = le32(r->buf.data() + r->pos);
    if (block_size < 32) {
      r->err = "bad BAM record size";
      return CLM_ERR_INVALID;
    }
    if (!r->ensure(4 + block_size)) {
      if (r->err.empty()) r->err = "truncated BAM record";
      return CLM_ERR_INVALID;
    }
    const uint8_t* rec = r->buf.data() + r->pos + 4;
    const size_t l_read_name = rec[8];
    const size_t n_cigar = le16(rec + 12);
    const uint32_t flag = le16(rec + 14);
    const size_t l_seq = le32(rec + 16);
    const size_t seq_off = 32 + l_read_name + 4 * n_cigar;
    const size_t aux_off = seq_off + (l_seq + 1) / 2 + l_seq;
    if (aux_off > block_size || l_read_name == 0) {
      r->err = "malformed BAM record";
      return CLM_ERR_INVALID;
    }
    bool keep = true;
    if (chimeric_only) {
      // is_chimeric (reference chimeralm/data/bam.py:21-23): mapped, has SA, primary line.
      keep = !(flag & 0x4) && !(flag & 0x100) && !(flag & 0x800);
      if (keep) {
        bool ok = true;
        keep = has_sa_tag(rec + aux_off, block_size - aux_off, &ok);
        if (!ok) {
          r->err = "malformed aux area in BAM record";
          return CLM_ERR_INVALID;
        }
      }
    }
    if (keep && r->shard_world > 1 && r->n_kept % r->shard_world != r->shard_rank) {
      ++r->n_kept;  // another rank's read: counted, not decoded
      keep = false;
    }
    if (keep) {
      const long long nb = std::min<long long>((long long)l_seq, max_bases);
      if (used + nb > bases_cap) {
        if (n_out == 0) {
          r->err = "bases_cap smaller than one read";
          return CLM_ERR_INVALID;
        }
        break;  // leave the record for the next call
      }
      const uint8_t* sq = rec + seq_off;
      uint8_t* dst = bases + used;
      const long long pairs = nb / 2;
      for (long long i = 0; i < pairs; ++i) {
        const uint16_t two = kNib.v[sq[i]];
        dst[2 * i] = uint8_t(two);
        dst[2 * i + 1] = uint8_t(two >> 8);
      }
      if (nb & 1) dst[nb - 1] = uint8_t(kNib.v[sq[pairs]]);
      if (names) {
        char* nm = names + n_out * (long long)name_stride;
        const size_t ln = std::min<size_t>(l_read_name - 1, size_t(name_stride - 1));
        std::memcpy(nm, rec + 32, ln);
        nm[ln] = 0;
      }
      used += nb;
      ++r->n_kept;
      ++n_out;
      offsets[n_out] = used;
    }
    r->pos += 4 + block_size;
    ++r->n_records;
  }
  return n_out;
}

}  // extern "C"
