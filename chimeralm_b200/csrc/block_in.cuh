// Fused first half of a HyenaDNA block, one persistent kernel (SURVEY.md A.6 head + A.3):
//
//   xn      = LayerNorm1(res)              (normalised by the PRODUCER of res - block_mlp's output epilogue or
//                                           the embedding kernel; affine folded into W'/b', see below)
//   u       = xn * W_in^T + b_in                               (HyenaOperator.in_proj, 256 -> 768)
//   uc[c,t] = w[c,0] u[c,t-2] + w[c,1] u[c,t-1] + w[c,2] u[c,t] + cb[c]      (short_filter, causal k=3)
//   x0, x1, v = uc[0:256], uc[256:512], uc[512:768];   vx = v * x1            (first gate)
//   outputs : vx, x0 as channel-major bf16 [B][256][Tp]  (what the long convolution consumes)
//
// Replaces nn.LayerNorm + nn.Linear + nn.Conv1d(groups=768) + split + mul of the reference and
// the transpose between them: per token the kernel reads 512 B (xn) and writes 1 KB (vx, x0);
// the 768-wide in_proj output never leaves the SM.
//
// The GEMM is computed TRANSPOSED: D[channel, token] = W'[channel,:] . xn[token,:], so that in
// TMEM a lane is a channel and a column is a token.  The epilogue thread that owns a channel
// then runs the 3-tap causal filter along its own registers and emits time-contiguous rows.
// A tile is 128 new tokens of one read plus a 16-token halo on the left (N = 144 columns; only
// the last two halo columns are used, 16 keeps N a legal UMMA shape), so tiles are independent.
// Per tile two passes (channel halves); each pass accumulates the x0/x1/v blocks of the same
// 128 channels in TMEM columns [0,144) [144,288) [288,432).
//
// A read's tail of 1..16 tokens (T = 128 f + L) rides on its last full tile: N = 160 columns, TMEM groups at a stride of 160
// (BlockInParams::ext_L).  Block 0 does not run this kernel at all: its input is one of 16 embedding rows, so
// in_proj is a table (embed_in.cuh).
//
// LayerNorm's affine is folded offline (clm_finalize): W' = W_in * diag(gamma),
// b' = b_in + W_in beta, so the kernel only normalises.
//
// Warp roles: warp 0 = TMA producer (weight ring, 2 x 32 KB), warp 1 = MMA issuer,
// warps 2..9 = epilogue (conv + gate + TMA store).  The B operand tile arrives by TMA: an earlier version
// normalised it in-kernel from the fp32 residual and spent 44% of each tile waiting on that burst of loads
// (profiles/r1_trace_block_in.txt).
// Weight slots are 32 KB single-TMA boxes (a single thread issues only ~1 TMA per 240 cycles,
// profiles/probes/tma_stream_probe.cu).
#pragma once
#include <cuda_bf16.h>

#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace clm {

struct BlockInParams {
  const float* b_in;    // [768] folded bias
  const float* cw;      // [768][3] short filter taps
  const float* cb;      // [768]   short filter bias
  int B, T;
  int tiles_per_seq, num_tiles;
  int vx_f16;           // write v * x1 as fp16 instead of bf16 (operand of the tensor-core long convolution)
  const float* vx_scale;  // [256] power-of-two factor a[ch] on v * x1 when vx_f16 (nullptr = 1): keeps the fp16 rows and every
                          // fp16 intermediate of the tensor-core FFT in range for weights of any magnitude (longconv_tc.cuh)
  long long* trace;     // optional [2][64] clock64 stamps of CTA 0 (row 0 = MMA issuer, row 1 = epilogue warp 2)
  int prefetch_xn;      // 1: the producer prefetches the next token tile into L2 (option in_prefetch)
  // ext_L > 0 (reads of 128 f + L tokens, 1 <= L <= 16): tiles_per_seq = f and the LAST tile of every read also computes the
  // L tail tokens - N = 160 token columns instead of 144, 16 more output columns for the second token half's warps, one more
  // store round that also zero-fills the rest of the read's last 128-token row.  The tail would otherwise be a tile of its
  // own that streams the same 384 KB of weights for L tokens (K2's 8 193-token reads: 2 080 tiles = 15 waves for 14 of work).
  int ext_L;
};

namespace bi {
constexpr int D = 256, BT = 128, HALO = 16, NCOL = BT + HALO;   // 144 token columns per tile
constexpr int EXT = 16, NCOL_EXT = NCOL + EXT;                  // 160 columns in a read's last tile when it carries the tail (ext_L)
constexpr int KB_ROWS_BYTES = NCOL * 128;                       // one k-block of xn: 144 rows x 128 B
constexpr int KB_ROWS_BYTES_EXT = NCOL_EXT * 128;               // ... 160 rows
constexpr int XN_BYTES = 4 * KB_ROWS_BYTES_EXT;                 // 81920 (room for the extended tile)
constexpr int SLOT_BYTES = 2 * 128 * 64 * 2;                    // 32 KB: two k-blocks of [128 channels x 64 k], ONE TMA box
constexpr int NSLOT = 3;                                        // ring depth x 32 KB is what hides the TMA latency: with 2 slots
                                                                // (one of look-ahead) a pass took 4.4-5.9 K cycles to issue 3.5 K of MMAs
constexpr int STAGE_BOX = 128 * 128;                            // 16 KB: [128 channels x 64 tokens] bf16
constexpr int OFF_XN = 0;
constexpr int OFF_W = OFF_XN + XN_BYTES;                        // 73728 (multiple of 1024)
constexpr int OFF_STAGE = OFF_W + NSLOT * SLOT_BYTES;           // 139264
constexpr int OFF_BAR = OFF_STAGE + 2 * STAGE_BOX;              // staging: ONE tensor at a time (x0, then v * x1), two token halves
constexpr int SMEM_TOTAL = OFF_BAR + 256;
constexpr int THREADS = 320, EPI_THREADS = 256;
constexpr int GCOLS = NCOL_EXT;                                 // TMEM columns per channel group (3 x 160 <= 512)
}  // namespace bi

__device__ __forceinline__ void tmem_ld_32x32b_x2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(bi::THREADS, 1)
block_in_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmVX,
                const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmXN,
                const __grid_constant__ CUtensorMap tmXNE, BlockInParams p) {
  using namespace bi;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  ptx::griddep_launch();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;          // [NSLOT]
  uint64_t* w_empty = bars + 4;     // [NSLOT] (NSLOT <= 4)
  uint64_t* xn_full = bars + 8;     // LN warps wrote the B operand
  uint64_t* xn_free = bars + 9;     // both passes' MMAs finished reading it
  // Two accumulator sets per pass so that the tensor pipe never waits for the epilogue to drain TMEM: A = the x0 group
  // (columns [0,144)), B = the x1 and v groups ([144,432)).  While the epilogue drains one set the MMAs fill the other.
  uint64_t* acc_full = bars + 10;   // [2] set A / B accumulated
  uint64_t* acc_free = bars + 12;   // [2] epilogue drained set A / B
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* trace = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  int trace_n = 0;
  auto stamp = [&](int role) {
    if (trace && (threadIdx.x & 31) == 0 && trace_n < 64) trace[role * 64 + trace_n++] = clock64();
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmVX); ptx::prefetch_tmap(&tmX0); ptx::prefetch_tmap(&tmXN);
    ptx::prefetch_tmap(&tmXNE);
    for (int i = 0; i < NSLOT; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    ptx::mbar_init(xn_full, 1); ptx::mbar_init(xn_free, 1);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_free[i], 8); }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::griddep_wait();   // xn comes from the kernel before this one

  if (warp == 0) {
    // =========================== TMA producer: 12 weight slots (2 k-blocks each) per tile ===========
    // W' is pre-tiled at finalize as [n/128][kb][128][64]: k-blocks (2kp, 2kp+1) of one 128-channel block
    // are 256 consecutive rows of a [.. x 64] matrix, i.e. one 32 KB TMA box per slot.
    if (lane == 0) {
      uint32_t wi = 0, it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        {  // B operand: normalised tokens [t0 - 16, t0 + 128) of read b, 4 k-blocks of [144 rows x 64]; rows
           // outside [0, T) are zero-filled by TMA (their products are discarded by the epilogue)
          const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * BT;
          const bool ext = p.ext_L > 0 && (tile % p.tiles_per_seq) == p.tiles_per_seq - 1;   // 160-row box (tmXNE)
          const int kbb = ext ? KB_ROWS_BYTES_EXT : KB_ROWS_BYTES;
          ptx::mbar_wait(xn_free, (it & 1) ^ 1);
          ptx::mbar_expect_tx(xn_full, 4 * kbb);
          for (int kb = 0; kb < 4; ++kb) ptx::tma_load_3d(smem + OFF_XN + kb * kbb, ext ? &tmXNE : &tmXN, xn_full, kb * 64, t0 - HALO, b);
          // The token tile is single-buffered (72 KB next to the weight ring), so this load is issued only when the previous
          // tile's MMAs are done and its 3.5-4 K cycles were fully exposed (profiles/r2_trace_block_in.txt) - most of it HBM
          // time for activations the previous kernel wrote.  Ask for the NEXT tile now: by the time it is loaded it sits in L2.
          if (p.prefetch_xn) {
            const int nxt = tile + gridDim.x;
            if (nxt < p.num_tiles) {
              const int nb = nxt / p.tiles_per_seq, nt0 = (nxt % p.tiles_per_seq) * BT;
              for (int kb = 0; kb < 4; ++kb) ptx::tma_prefetch_3d(&tmXN, kb * 64, nt0 - HALO, nb);
            }
          }
        }
        for (int h = 0; h < 2; ++h)
          for (int g = 0; g < 3; ++g)
            for (int kp = 0; kp < 2; ++kp) {
              const uint32_t s = wi % NSLOT, ph = (wi / NSLOT) & 1;
              ptx::mbar_wait(&w_empty[s], ph ^ 1);
              ptx::mbar_expect_tx(&w_full[s], SLOT_BYTES);
              ptx::tma_load_2d(smem + OFF_W + s * SLOT_BYTES, &tmW, &w_full[s], 0, ((g * 2 + h) * 4 + 2 * kp) * 128);
              ++wi;
            }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    {   // whole warp, uniform control flow; one elected lane issues (ptx::umma_f16_e)
      constexpr uint32_t idesc_n = ptx::idesc_bf16_f32(128, NCOL), idesc_e = ptx::idesc_bf16_f32(128, NCOL_EXT);
      const uint32_t sXN = ptx::smem_u32(smem + OFF_XN), sW = ptx::smem_u32(smem + OFF_W);
      uint32_t wi = 0, it = 0, pass = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        stamp(0);
        ptx::mbar_wait(xn_full, it & 1);
        stamp(0);
        const bool ext = p.ext_L > 0 && (tile % p.tiles_per_seq) == p.tiles_per_seq - 1;
        const uint32_t idesc = ext ? idesc_e : idesc_n;
        const uint32_t kbb = ext ? KB_ROWS_BYTES_EXT : KB_ROWS_BYTES;
        for (int h = 0; h < 2; ++h, ++pass) {
          for (int g = 0; g < 3; ++g) {
            if (g < 2) {   // g == 0 starts set A, g == 1 starts set B
              ptx::mbar_wait(&acc_free[g], (pass & 1) ^ 1);
              ptx::tc_fence_after_sync();
              if (g == 0) stamp(0);
            }
            for (int kp = 0; kp < 2; ++kp) {
              const uint32_t s = wi % NSLOT, ph = (wi / NSLOT) & 1;
              ptx::mbar_wait(&w_full[s], ph);
              ptx::tc_fence_after_sync();
#pragma unroll
              for (int q2 = 0; q2 < 2; ++q2) {
                const int kb = 2 * kp + q2;
                const uint64_t da = ptx::smem_desc_k_sw128(sW + s * SLOT_BYTES + q2 * (SLOT_BYTES / 2));
                const uint64_t db = ptx::smem_desc_k_sw128(sXN + kb * kbb);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma_f16_e(tmem_base + g * GCOLS, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              }
              ptx::umma_commit_e(&w_empty[s]);
              ++wi;
            }
            if (g == 0) ptx::umma_commit_e(&acc_full[0]);
            if (g == 2) ptx::umma_commit_e(&acc_full[1]);
          }
          if (h == 1) ptx::umma_commit_e(xn_free);
          stamp(0);
        }
      }
    }
  } else {
    // =========================== LayerNorm producers + epilogue ===========================
    const int e = warp - 2;
    const int q = warp & 3;            // TMEM lane quarter
    const int hf = e >> 2;             // LN: column half of the row; epilogue: token half of the tile
    const int r = q * 32 + lane;       // LN: row 0..127 of the tile; epilogue: channel inside the pass
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t sST = ptx::smem_u32(smem + OFF_STAGE);
    const bool issuer = (threadIdx.x == 64);
    // per-thread channel constants for both passes (thread = channel h*128 + r of each group): loaded once
    float c_bia[2][3], c_w0[2][3], c_w1[2][3], c_w2[2][3], c_cb[2][3];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        const int ch = g * 256 + h * 128 + r;
        c_bia[h][g] = __ldg(p.b_in + ch);
        c_w0[h][g] = __ldg(p.cw + ch * 3);
        c_w1[h][g] = __ldg(p.cw + ch * 3 + 1);
        c_w2[h][g] = __ldg(p.cw + ch * 3 + 2);
        c_cb[h][g] = __ldg(p.cb + ch);
        if (g == 2 && p.vx_scale) {   // fold a[ch] into the v group's short-filter taps and bias: a * v exactly, no extra work
          const float a = __ldg(p.vx_scale + h * 128 + r);
          c_w0[h][g] *= a; c_w1[h][g] *= a; c_w2[h][g] *= a; c_cb[h][g] *= a;
        }
      }
    uint32_t it = 0, pass = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_seq;
      const int t0 = (tile % p.tiles_per_seq) * BT;
      const bool ext = p.ext_L > 0 && (tile % p.tiles_per_seq) == p.tiles_per_seq - 1;   // this tile also carries the read's tail
      const bool tr = trace && warp == 2 && lane == 0;
      if (tr) stamp(1);
      // ------------------------------------------------ two passes of conv + gate epilogue
#pragma unroll 1
      for (int h = 0; h < 2; ++h, ++pass) {
        float bia[3], w0[3], w1[3], w2[3], cbv[3];
#pragma unroll
        for (int g = 0; g < 3; ++g) {
          bia[g] = h ? c_bia[1][g] : c_bia[0][g];
          w0[g] = h ? c_w0[1][g] : c_w0[0][g];
          w1[g] = h ? c_w1[1][g] : c_w1[0][g];
          w2[g] = h ? c_w2[1][g] : c_w2[0][g];
          cbv[g] = h ? c_cb[1][g] : c_cb[0][g];
        }
        // staging buffers may still be read by the previous pass's TMA stores
        if (issuer) ptx::tma_store_wait_read<0>();
        ptx::bar_sync(1, EPI_THREADS);
        if (tr) stamp(1);
        const int cbase = HALO + hf * 64;   // first output column of this thread
        const uint32_t swz = uint32_t(r & 7);
        const uint32_t rowoff = uint32_t(r) * 128;
        float hm2[3], hm1[3];               // u[j-2], u[j-1] carried along the columns
        auto halo = [&](int g) {
          uint32_t a, c2;
          tmem_ld_32x32b_x2(lane_addr + g * GCOLS + cbase - 2, a, c2);
          ptx::tmem_ld_wait();
          const int tm2 = t0 - HALO + cbase - 2;
          hm2[g] = (tm2 >= 0) ? __uint_as_float(a) + bia[g] : 0.f;
          hm1[g] = (tm2 + 1 >= 0) ? __uint_as_float(c2) + bia[g] : 0.f;
        };
        // group g, sub-block s (32 token columns): 3-tap causal filter along this thread's registers
        auto conv_sub = [&](int g, int s, float (&out)[32]) {
          uint32_t a[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + g * GCOLS + cbase + s * 32, a);
          ptx::tmem_ld_wait();
          float um2 = hm2[g], um1 = hm1[g];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float u = __uint_as_float(a[j]) + bia[g];   // cbase + s*32 + j >= HALO -> t >= 0
            out[j] = fmaf(w0[g], um2, fmaf(w1[g], um1, fmaf(w2[g], u, cbv[g])));
            um2 = um1;
            um1 = u;
          }
          hm2[g] = um2;
          hm1[g] = um1;
        };
        // the 16 extra columns [144, 160) of an extended tile: they follow the second token half's last sub-block, so its
        // warps carry on with their filter state; tokens past the end of the read come out as ZERO
        auto conv_ext = [&](int g, float (&out)[EXT]) {
          uint32_t a[8], c8[8];
          ptx::tmem_ld_32x32b_x8(lane_addr + g * GCOLS + NCOL, a);
          ptx::tmem_ld_32x32b_x8(lane_addr + g * GCOLS + NCOL + 8, c8);
          ptx::tmem_ld_wait();
          float um2 = hm2[g], um1 = hm1[g];
#pragma unroll
          for (int j = 0; j < EXT; ++j) {
            const float u = __uint_as_float(j < 8 ? a[j] : c8[j - 8]) + bia[g];
            out[j] = fmaf(w0[g], um2, fmaf(w1[g], um1, fmaf(w2[g], u, cbv[g])));
            um2 = um1;
            um1 = u;
          }
        };
        // store round of an extended tile: tokens [t0 + 128, t0 + 256) = the read's last 128-token row.  The second half's
        // warps filter and stage the 16 tail columns (zeros from ext_L on) and zeros up to 64, the first half's warps the
        // all-zero second box.  The tail columns are read from TMEM HERE, after the tile's main store was issued - the values
        // live in registers only inside this round; the accumulator set is released to the next pass afterwards (one short
        // tensor-pipe stall per pass of a read's last tile, 32 of K2's 2 048 tiles).
        auto store_ext = [&](int which) {   // 0: x0 (set A), 1: v * x1 (set B)
          float val[EXT];
          if (hf == 1) {
            if (which == 0) {
              conv_ext(0, val);
            } else {
              float x1e[EXT];
              conv_ext(1, x1e);
              conv_ext(2, val);
#pragma unroll
              for (int j = 0; j < EXT; ++j) val[j] *= x1e[j];
            }
          }
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_free[which]);
          const bool as_f16 = which == 1 && p.vx_f16;
          if (issuer) ptx::tma_store_wait_read<0>();
          ptx::bar_sync(1, EPI_THREADS);
          const uint32_t rowaddr = sST + hf * STAGE_BOX + rowoff;   // hf 1 -> box 1 (tail + zeros), hf 0 -> box 0 (zeros)
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0;
            if (hf == 1 && k < 2) {
              float m[8];
#pragma unroll
              for (int e8 = 0; e8 < 8; ++e8) m[e8] = (k * 8 + e8 < p.ext_L) ? val[k * 8 + e8] : 0.f;
              if (as_f16) { q0 = pack_f16(m[0], m[1]); q1 = pack_f16(m[2], m[3]); q2 = pack_f16(m[4], m[5]); q3 = pack_f16(m[6], m[7]); }
              else { q0 = pack_bf16(m[0], m[1]); q1 = pack_bf16(m[2], m[3]); q2 = pack_bf16(m[4], m[5]); q3 = pack_bf16(m[6], m[7]); }
            }
            ptx::st_shared_v4(rowaddr + ((uint32_t(k) ^ swz) << 4), q0, q1, q2, q3);
          }
          ptx::fence_proxy_async_smem();
          ptx::bar_sync(2, EPI_THREADS);
          if (issuer) {
            const CUtensorMap* tm = which == 0 ? &tmX0 : &tmVX;
            ptx::tma_store_3d(tm, smem + OFF_STAGE + STAGE_BOX, t0 + BT, h * 128, b);   // staged by the hf == 1 warps
            ptx::tma_store_3d(tm, smem + OFF_STAGE, t0 + BT + 64, h * 128, b);         // zeros
            ptx::tma_store_commit();
          }
        };
        // ---- set A: x0
        ptx::mbar_wait(&acc_full[0], pass & 1);
        ptx::tc_fence_after_sync();
        if (tr) stamp(1);
        halo(0);
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
          float x0v[32];
          conv_sub(0, s, x0v);
          if (s == 1 && !ext) {              // set A is in registers: the next pass may overwrite it
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc_free[0]);
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {     // 4 chunks of 8 tokens
            const float* x0p = &x0v[k * 8];
            const uint32_t chunk = (uint32_t(s * 4 + k) ^ swz) << 4;
            ptx::st_shared_v4(sST + hf * STAGE_BOX + rowoff + chunk, pack_bf16(x0p[0], x0p[1]),
                              pack_bf16(x0p[2], x0p[3]), pack_bf16(x0p[4], x0p[5]), pack_bf16(x0p[6], x0p[7]));
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::bar_sync(2, EPI_THREADS);
        if (tr) stamp(1);   // x0 staged
        if (issuer) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) ptx::tma_store_3d(&tmX0, smem + OFF_STAGE + hh * STAGE_BOX, t0 + hh * 64, h * 128, b);
          ptx::tma_store_commit();
        }
        if (ext) store_ext(0);
        // ---- set B: v * x1
        ptx::mbar_wait(&acc_full[1], pass & 1);
        ptx::tc_fence_after_sync();
        if (tr) stamp(1);   // set B accumulated
        halo(1);
        halo(2);
#pragma unroll 1
        for (int s = 0; s < 2; ++s) {
          float x1v[32], vv[32];
          conv_sub(1, s, x1v);
          conv_sub(2, s, vv);
          if (s == 1 && !ext) {
            ptx::tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&acc_free[1]);
          }
          if (s == 0) {   // the staging buffers are being read by the x0 store of this pass (long since issued)
            if (tr) stamp(1);   // first half of set B convolved
            if (issuer) ptx::tma_store_wait_read<0>();
            ptx::bar_sync(1, EPI_THREADS);
            if (tr) stamp(1);   // staging free again
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float* x1p = &x1v[k * 8];
            const float* vp = &vv[k * 8];
            const uint32_t chunk = (uint32_t(s * 4 + k) ^ swz) << 4;
            if (p.vx_f16) {
              // the tensor-core conv reads whole 128-token rows: positions past the end of the read must be ZERO
              float m[8];
#pragma unroll
              for (int e8 = 0; e8 < 8; ++e8) m[e8] = (t0 + hf * 64 + s * 32 + k * 8 + e8 < p.T) ? vp[e8] * x1p[e8] : 0.f;
              ptx::st_shared_v4(sST + hf * STAGE_BOX + rowoff + chunk, pack_f16(m[0], m[1]), pack_f16(m[2], m[3]),
                                pack_f16(m[4], m[5]), pack_f16(m[6], m[7]));
            } else {
              ptx::st_shared_v4(sST + hf * STAGE_BOX + rowoff + chunk, pack_bf16(vp[0] * x1p[0], vp[1] * x1p[1]),
                                pack_bf16(vp[2] * x1p[2], vp[3] * x1p[3]), pack_bf16(vp[4] * x1p[4], vp[5] * x1p[5]),
                                pack_bf16(vp[6] * x1p[6], vp[7] * x1p[7]));
            }
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::bar_sync(2, EPI_THREADS);
        if (issuer) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) ptx::tma_store_3d(&tmVX, smem + OFF_STAGE + hh * STAGE_BOX, t0 + hh * 64, h * 128, b);
          ptx::tma_store_commit();
        }
        if (ext) store_ext(1);
        if (tr) stamp(1);
      }
    }
    if (issuer) ptx::tma_store_wait<0>();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// W'[n,k] = W[n,k] * gamma[k];  b'[n] = b[n] + sum_k W[n,k] * beta[k]   (one block per output row n)
__global__ void __launch_bounds__(256) fold_ln_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      float* __restrict__ Wf, float* __restrict__ bf, int K) {
  __shared__ float red[8];
  const int n = blockIdx.x;
  float acc = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float w = W[(long long)n * K + k];
    Wf[(long long)n * K + k] = w * gamma[k];
    acc += w * beta[k];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    bf[n] = bias[n] + t;
  }
}

}  // namespace clm
