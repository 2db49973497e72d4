// Fused second half of a HyenaDNA block, one persistent kernel (SURVEY.md A.3 tail + A.6):
//
//   r1  = y * Wout^T + b_out + res                 (HyenaOperator.out_proj + residual)
//   xn  = LayerNorm2(r1)
//   h   = gelu_tanh(xn * W1^T + b1)                (HyenaMlp.fc1)
//   out = h * W2^T + b2 + r1                       (HyenaMlp.fc2 + residual)       -> res (fp32, R32 layout)
//
// Replaces 4 reference ops (nn.Linear x3, nn.LayerNorm, F.gelu) and keeps r1, xn and the
// 1024-wide hidden activations on chip: per token the kernel reads 512 B (y, bf16) + 1 KB
// (res) and writes 1 KB (res), instead of ~9.2 KB for the unfused sequence.
//
// Round 2 additions: reads that end in a partial tile of <= 64 tokens share GATHERED tiles (rows are
// independent in all three GEMMs: BlockMlpParams::gather_L, gather_tails_kernel), and block 0 reads its residual input from the
// embedding table by token id (BlockMlpParams::res_tab) - there is no embedding kernel on the product path.
//
// One CTA per SM loops over tiles of 128 tokens.  All three GEMMs run on tcgen05 with the
// accumulators in TMEM:
//   (the two 256-column halves swap roles every tile)
//   cols [0,256)    R  : out_proj accumulator, rewritten in place by the epilogue as r1 and then used as
//                        the fc2 accumulator (so the fc2 result already carries the residual)
//   cols [256,384)  XN : LayerNorm2 output as packed bf16 - the A operand of fc1 read from TMEM
//                        (tcgen05.mma TS form), so fc1 streams only its weights from shared memory
//   cols [384,512)  H  : fc1 chunk accumulator (drained into registers before the GELU math)
// Shared memory (224 KB): HB (2 x 32 KB: gelu(h) chunks as the A operand of fc2) and a 5 x 32 KB ring
// through which TMA streams, in exactly the order the MMA thread consumes them, the y tile (two
// k-blocks per slot) and every weight tile (one 32 KB box per slot).  A slot is held until the MMAs
// reading it complete, so ring depth x slot size is what hides the L2 latency.
//
// Warp roles (three warpgroups, 384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (+TMEM alloc), warps 2..3 idle,
// warps 4..11 = epilogue (TMEM lane quarter = warp % 4, column half = (warp - 4) / 4).  The roles sit on warpgroup
// boundaries so that `setmaxnreg` can move registers from the first warpgroup (104 each) to the epilogue warps (200
// each): with 10 warps three of them share a scheduler's 16 K registers and ptxas caps the kernel at 168, too few to
// hold the next tile's residual half-row (128 registers) across the output epilogue without spilling loaded values.
#pragma once
#include <cuda_bf16.h>

#include <type_traits>

#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace clm {

struct BlockMlpParams {
  int M;                 // tokens
  float* res;            // [M,256] fp32 (R32 layout), read and overwritten
  int layer;             // index into bm::c_mlp (biases; LayerNorm2 affine is folded into W1/b1)
  float eps;
  int num_tiles;
  // y_cm == 0: y is token-major [M,256] (tmY 2-D, tiles of 128 consecutive rows).
  // y_cm == 1: y is channel-major [B][256][Tp] (tmY 3-D {t, c, b}); tiles are 128 tokens of ONE read,
  //            fed to out_proj as an MN-major A operand, so no transpose pass is needed.
  int y_cm, T, tiles_per_seq;
  // Tail gathering (y_cm only; gather_L == 0: off).  A read of T = 128 f + L tokens with 1 <= L <= 64 would end in a tile
  // with L valid rows that costs as much as a full one (B of them per launch: K2's 8 193-token reads pay 2 080 tiles =
  // 15 waves for 2 048 tiles of work).  Instead tiles_per_seq = f, tiles [0, n_full_tiles) are the full ones, and the tails
  // of gather_P = 128 / L reads share each of the remaining tiles: row r of gathered tile g is token 128 f + r % L of read
  // g P + r / L.  Their y operand comes from yg (tmYG: channel-major [256][gathered tiles x 128], written by
  // gather_tails_kernel), their normalised rows go to xn by plain stores.
  int gather_L, gather_P, n_full_tiles, B;
  __nv_bfloat16* xn;     // xn rows [B*T][256] (gathered tiles only; the full tiles use tmXN)
  // res_tab != nullptr (block 0): the residual INPUT of row i is the embedding row of token id ids[i], read from a 32-row table
  // in the R32 layout (L2-resident) - the embedding kernel and its 1 KB per token of HBM writes and reads disappear.  The
  // output still goes to res.
  const float* res_tab;
  const void* ids;       // [M] token ids, element type ids_dtype (2 = u8, 3 = i32, 4 = i64: clm_dtype)
  int ids_dtype, vocab_rows;
  // write_xn: the output epilogue also emits xn = (out - mean) * rstd as bf16 [B][T][256] (tmXN, 3-D {col, t, b}):
  // the next consumer's LayerNorm (affine folded into its weights) without another pass over the residual.
  int write_xn;
  // skip_res_store: do not write the fp32 residual back (the LAST block's residual has no reader: the scorer and the
  // pooling consume the normalised xn rows) - 128 KB of stores per tile less on the SM's ~30 B/clk store path
  int skip_res_store;
  int stagger_cycles;    // CTA b starts (b % 4) * stagger_cycles late (0 = off)
  int store_a;           // 0..4: column groups (of 4) whose residual stores are issued in E3's statistics sweep
  int helpers_high;      // 1: the helper warps (producer, MMA issuer) take the highest warp ids, the epilogue warps 0..7
  long long* trace;      // optional [3][64] clock64 stamps written by CTA 0 (null in production)
};

namespace bm {
constexpr int MAX_LAYERS = 8;
// Per-layer bias vectors, read with warp-uniform addresses in the epilogues -> constant cache
// (with 226 KB of shared memory per CTA only ~4 KB of L1 is left, so __ldg thrashed).
struct LayerConsts {
  float b_out[256];
  float b2[256];
  float b1[1024];   // (fc1 bias with LayerNorm2's beta folded in) / 2, see gelu_tanh_bf16x2
};
__constant__ LayerConsts c_mlp[MAX_LAYERS];

constexpr int D = 256, DI = 1024, BM = 128, BK = 64;
constexpr int KB_BYTES = BM * BK * 2;            // 16 KB: one [128 x 64] bf16 tile
constexpr int X_BYTES = 4 * KB_BYTES;            // 64 KB
constexpr int HB_BYTES = 2 * KB_BYTES;           // 32 KB per buffer
constexpr int SLOT_BYTES = 2 * KB_BYTES;         // 32 KB
constexpr int NSLOT = 5;                         // 160 KB ring: weights AND the y tile stream through it
constexpr int OFF_HB = 0;
constexpr int OFF_W = OFF_HB + 2 * HB_BYTES;
constexpr int OFF_BAR = OFF_W + NSLOT * SLOT_BYTES;   // 229376
constexpr int OFF_PART = OFF_BAR + 256;               // LayerNorm partial sums [2][2][128] fp32
constexpr int SMEM_TOTAL = OFF_PART + 2 * 2 * BM * 4; // 231680 <= 232448 (227 KB)
constexpr int THREADS = 320;            // the experimental variants (block_mlp2/16/pp): warps 0, 1, 2..9
constexpr int THREADS_WG = 384;         // block_mlp_kernel: three warpgroups
constexpr int EPI_WARP0 = 4;
constexpr int EPI_TID0 = EPI_WARP0 * 32;
constexpr int EPI_THREADS = 256;
constexpr int NCHUNK = DI / 128;                 // 8 fc1 column chunks

// The fused kernels' fc1 weights and bias are stored pre-multiplied by 1/2 (exact in bf16 / fp32), so the accumulator
// holds h = x / 2 and gelu_tanh(x) = h + h tanh(u), u = sqrt(2/pi) (x + 0.044715 x^3) = h (2 c0 + 8 c1 h^2): one multiply per
// element less than starting from x.
__device__ __forceinline__ float gelu_tanh_fast(float h) {
  // tanh.approx is ONE MUFU op (the exp + rcp formulation needs two and made the GELU epilogue MUFU-bound at the
  // MMA's pace).  |tanh.approx error| <= 2^-10.987, well below the bf16 rounding applied to the result.
  const float u = h * fmaf(0.285419265090401f, h * h, 1.5957691216057308f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  return fmaf(h, t, h);
}
// Two GELUs at once on packed fp32 pairs (FMUL2/FFMA2): 4 issue slots per element instead of 8.
__device__ __forceinline__ uint32_t gelu_tanh_bf16x2(f2t h) {
  const f2t C0 = f2_pack(1.5957691216057308f, 1.5957691216057308f), C1 = f2_pack(0.285419265090401f, 0.285419265090401f);
  const f2t u = f2_mul(h, f2_fma(f2_mul(h, h), C1, C0));
  float u0, u1, t0, t1;
  f2_unpack(u, u0, u1);
#if defined(CLM_E2_DIAG) && CLM_E2_DIAG == 1   // timing diagnostic only (wrong results): no MUFU
  t0 = u0; t1 = u1;
#else
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
#endif
#if defined(CLM_E2_DIAG) && CLM_E2_DIAG == 2   // timing diagnostic only (wrong results): MUFU only, no tail arithmetic
  return pack_bf16(t0, t1);
#else
  float r0, r1;
  f2_unpack(f2_fma(h, f2_pack(t0, t1), h), r0, r1);
  return pack_bf16(r0, r1);
#endif
}
}  // namespace bm

// y [B][D][Tp] channel-major -> yg [D][TG]: column 128 g + r = token t0 + r % L of read g P + r / L (zero where there is none)
__global__ void gather_tails_kernel(const __nv_bfloat16* __restrict__ y, __nv_bfloat16* __restrict__ yg, int B, int D, int Tp,
                                    int TG, int t0, int L, int P) {
  ptx::griddep_launch();
  ptx::griddep_wait();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= D * TG) return;
  const int c = idx / TG, col = idx - c * TG;
  const int g = col >> 7, r = col & 127, jj = r / L, l = r - jj * L, j = g * P + jj;
  yg[idx] = (jj < P && j < B) ? y[((size_t)j * D + c) * Tp + t0 + l] : __float2bfloat16(0.f);
}

// EARLY_RES: the next tile's residual half-row is loaded into registers BEFORE the output epilogue (E3) of this tile
// instead of after it, so the loads (per-SM outstanding-miss bound: ~6.5 K cycles for 128 KB even from L2) complete
// under E3's stores instead of in front of E1.
// LAG: fc2 of chunk j - LAG is issued right after fc1 of chunk j (1 in production).  tcgen05.mma issue blocks while the MMA
// queue is full, so the issuing thread runs in step with the tensor pipe: per chunk it spends ~1 440 cycles issuing fc1
// (16 TS-form MMAs, 90 cycles each instead of the probe's 64), ~1 070 issuing fc2 and ~880 in barrier waits that nothing
// overlaps - that sum is the 3.1 K chunk period.  LAG = 2 (wait for the drain first, the fc2 issued afterwards has had its
// operand ready for a whole chunk) moves time from the gelu(h) wait to the drain wait and is no faster; a second issuing
// thread for fc2 was slower (its MMAs queue in front of fc1 on the epilogue's critical path).  profiles/r1_trace_block_mlp.txt.
template <int EARLY_RES, int LAG = 1>   // EARLY_RES: number of the 32 float4 loaded early (0 = all after E3)
__global__ void __launch_bounds__(bm::THREADS_WG, 1)
block_mlp_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWout,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmXN, const __grid_constant__ CUtensorMap tmYG, BlockMlpParams p) {
  using namespace bm;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // swizzled UMMA/TMA tiles need 1 KB alignment
  ptx::griddep_launch();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;                 // [NSLOT]
  uint64_t* w_empty = bars + 5;            // [NSLOT]
  uint64_t* g1_done = bars + 10;           // out_proj accumulator complete
  uint64_t* xn_full = bars + 11;           // epilogue wrote xn into TMEM and r1 into R
  uint64_t* hacc_full = bars + 12;         // fc1 chunk accumulator complete
  uint64_t* hacc_free = bars + 13;         // epilogue drained the fc1 chunk accumulator into registers
  uint64_t* hbuf_full = bars + 14;         // [2] gelu(h) chunk written to HB
  uint64_t* hbuf_free = bars + 16;         // [2] fc2 finished reading HB
  uint64_t* out_full = bars + 18;          // fc2 accumulator complete
  uint64_t* r_free = bars + 19;            // epilogue drained R
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);
  float (*s_part)[2][BM] = reinterpret_cast<float (*)[2][BM]>(smem + OFF_PART);  // [half][sum|sumsq][row]

  // `warp` is the ROLE index (0 producer, 1 MMA issuer, 2-3 idle, 4..11 epilogue).  With helpers_high the roles 0..3 sit on the
  // physical warps 8..11: the warp scheduler favours higher warp ids, and the MMA issuer shares its scheduler with two
  // epilogue warps.  The TMEM lane quarter (physical warp % 4) is the same either way.
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = p.helpers_high ? (pwarp + 4) % 12 : pwarp;
  // Each CTA walks the 8 fc1/fc2 hidden chunks starting at a different one, so the 148 CTAs do
  // not all pull the same weight tile out of L2 at the same moment.
  const int rot = blockIdx.x & (NCHUNK - 1);
  // trace rows: 0 = producer, 1 = MMA issuer, 2 = epilogue warp 2; CTA 0 only
  long long* trace = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  int trace_n = 0;
  auto stamp = [&](int role) {
    if (trace && lane == 0 && trace_n < 64) trace[role * 64 + trace_n++] = clock64();
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmY); ptx::prefetch_tmap(&tmWout); ptx::prefetch_tmap(&tmW1); ptx::prefetch_tmap(&tmW2);
    ptx::prefetch_tmap(&tmXN); ptx::prefetch_tmap(&tmYG);
    for (int i = 0; i < NSLOT; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    ptx::mbar_init(g1_done, 1);
    ptx::mbar_init(xn_full, 8);
    ptx::mbar_init(hacc_full, 1); ptx::mbar_init(hacc_free, 8);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&hbuf_full[i], 8); ptx::mbar_init(&hbuf_free[i], 1); }
    ptx::mbar_init(out_full, 1); ptx::mbar_init(r_free, 8);
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::griddep_wait();   // y, the residual and (gathered tiles) yg come from the kernels before this one
  // All CTAs run the same tile schedule in lockstep, so their residual reads (E1) and writes (E3) would hit HBM as
  // chip-wide bursts while the tensor pipes idle.  Starting the CTAs in `stagger` phase groups spreads that traffic.
  if (p.stagger_cycles > 0) {
    const long long t_end = clock64() + (long long)(blockIdx.x & 3) * p.stagger_cycles;
    while (clock64() < t_end) {}
  }

  // setmaxnreg sits INSIDE each role's branch: ptxas applies the smallest count to everything after a point where the
  // branches have merged again
  if (warp == 0) {
    // =========================== TMA producer ===========================
    ptx::setmaxnreg_dec<104>();
    if (lane == 0) {
      uint32_t wi = 0;  // running weight-slot counter
      auto slot_acquire = [&]() -> uint8_t* {
        const uint32_t s = wi % NSLOT, ph = (wi / NSLOT) & 1;
        ptx::mbar_wait(&w_empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&w_full[s], SLOT_BYTES);
        return smem + OFF_W + s * SLOT_BYTES;
      };
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        // y tile: 4 k-blocks of 16 KB, two per ring slot
        for (int kp = 0; kp < 2; ++kp) {
          uint8_t* s = slot_acquire();
          uint64_t* fb = &w_full[wi % NSLOT];
          for (int q = 0; q < 2; ++q) {
            const int kb = 2 * kp + q;
            if (p.y_cm) {
              const bool gt = tile >= p.n_full_tiles;   // gathered tails: same box, from yg
              const CUtensorMap* tm = gt ? &tmYG : &tmY;
              const int b = gt ? 0 : tile / p.tiles_per_seq;
              const int t0 = gt ? (tile - p.n_full_tiles) * BM : (tile % p.tiles_per_seq) * BM;
              for (int hh = 0; hh < 2; ++hh)   // k-block = 64 channels; two 64-token halves of 8 KB each
                ptx::tma_load_3d(s + q * KB_BYTES + hh * (KB_BYTES / 2), tm, fb, t0 + hh * 64, kb * BK, b);
            } else {
              ptx::tma_load_2d(s + q * KB_BYTES, &tmY, fb, kb * BK, tile * BM);
            }
          }
          ++wi;
          if (kp == 0) stamp(0);
        }
        // Weights are pre-tiled at finalize as [N/rt][K/64][rt][64] (rt = 256 for Wout/W2, 128 for W1), so
        // every 32 KB slot is ONE TMA instruction (a single thread issues ~1 TMA per 240 cycles).
        for (int kb = 0; kb < 4; ++kb) {               // out_proj: k-block kb = rows [256 kb, +256)
          uint8_t* s = slot_acquire();
          ptx::tma_load_2d(s, &tmWout, &w_full[wi % NSLOT], 0, kb * 256);
          ++wi;
        }
        for (int j = 0; j < NCHUNK + LAG; ++j) {
          if (j < NCHUNK) {  // fc1 chunk jc: k-blocks (2 h2, 2 h2 + 1) = rows [(4 jc + 2 h2) 128, +256)
            const int jc = (j + rot) & (NCHUNK - 1);
            for (int h2 = 0; h2 < 2; ++h2) {
              uint8_t* s = slot_acquire();
              ptx::tma_load_2d(s, &tmW1, &w_full[wi % NSLOT], 0, (jc * 4 + 2 * h2) * 128);
              ++wi;
            }
          }
          if (j >= LAG) {  // fc2 K-chunk jj: k-blocks 2 jj + kb = rows [(2 jj + kb) 256, +256)
            const int jj = (j - LAG + rot) & (NCHUNK - 1);
            for (int kb = 0; kb < 2; ++kb) {
              uint8_t* s = slot_acquire();
              ptx::tma_load_2d(s, &tmW2, &w_full[wi % NSLOT], 0, (jj * 2 + kb) * 256);
              ++wi;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    ptx::setmaxnreg_dec<104>();
    {   // the WHOLE warp runs this loop (uniform control flow and descriptors); one elected lane issues, see ptx::umma_f16_e
      constexpr uint32_t idesc256 = ptx::idesc_bf16_f32(BM, 256);
      constexpr uint32_t idesc256_amn = ptx::idesc_bf16_f32_amn(BM, 256);
      constexpr uint32_t idesc128 = ptx::idesc_bf16_f32(BM, 128);
      const uint32_t sHB = ptx::smem_u32(smem + OFF_HB);
      const uint32_t sW = ptx::smem_u32(smem + OFF_W);
      uint32_t wi = 0;
      long long wt_slot = 0, wt_hbuf = 0, wt_hacc = 0, wt_tile = 0, wt_i1 = 0, wt_i2 = 0, t_all = clock64();
      // The barrier of the NEXT ring slot is probed right after the current slot is handed out, so the
      // ~100-cycle try_wait round trip overlaps the MMA issue instead of preceding every slot.
      bool next_ready = false;
      uint32_t probed_wi = 0xffffffffu;
      auto probe_next = [&](uint32_t w) {
        next_ready = ptx::mbar_try_wait(&w_full[w % NSLOT], (w / NSLOT) & 1);
        probed_wi = w;
      };
      auto slot_wait = [&]() -> uint32_t {
        const uint32_t s = wi % NSLOT, ph = (wi / NSLOT) & 1;
        if (!(probed_wi == wi && next_ready)) {
          const long long t_ = trace ? clock64() : 0;
          ptx::mbar_wait(&w_full[s], ph);
          if (trace) wt_slot += clock64() - t_;
        }
        ptx::tc_fence_after_sync();
        probe_next(wi + 1);
        return sW + s * SLOT_BYTES;
      };
      auto slot_release = [&]() {
        ptx::umma_commit_e(&w_empty[wi % NSLOT]);
        ++wi;
      };
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t tph = it & 1;
        // TMEM halves alternate per tile: R takes the half that held XN+H in the previous tile, so out_proj
        // of this tile runs while the epilogue is still draining the previous tile's R.
        const uint32_t TM_R = tph ? 256u : 0u, TM_XN = tph ? 0u : 256u, TM_H = tph ? 128u : 384u;
        // ---- G1: R = y * Wout^T
        stamp(1);
        { const long long t_ = trace ? clock64() : 0;
          ptx::mbar_wait(hacc_free, ((it * NCHUNK) & 1) ^ 1);   // previous tile's last fc1 chunk drained (long ago)
          if (trace) wt_tile += clock64() - t_; }
        ptx::tc_fence_after_sync();
        stamp(1);
        // ring order: y(kb 0,1), y(kb 2,3), Wout kb 0..3 - the y slots are released only after the last k-block
        const uint32_t sy0 = slot_wait(); const uint32_t wi_y0 = wi; ++wi;
        const uint32_t sy1 = slot_wait(); const uint32_t wi_y1 = wi; ++wi;
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t sw = slot_wait();
          const uint64_t db = ptx::smem_desc_k_sw128(sw);
          const uint32_t sx = ((kb < 2) ? sy0 : sy1) + (kb & 1) * KB_BYTES;
          if (p.y_cm) {
            // A = y^T tile: MN(token)-major, 2 atoms of 64 tokens 8 KB apart, K rows of 128 B; 16 K-rows per step
            const uint64_t da = ptx::smem_desc_mn_sw128(sx, KB_BYTES / 2, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16_e(tmem_base + TM_R, da + (2048 >> 4) * k, db + 2 * k, idesc256_amn, (kb | k) != 0);
          } else {
            const uint64_t da = ptx::smem_desc_k_sw128(sx);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_f16_e(tmem_base + TM_R, da + 2 * k, db + 2 * k, idesc256, (kb | k) != 0);
          }
          slot_release();
          if (kb == 1) ptx::umma_commit_e(&w_empty[wi_y0 % NSLOT]);
          if (kb == 3) ptx::umma_commit_e(&w_empty[wi_y1 % NSLOT]);
        }
        ptx::umma_commit_e(g1_done);
        stamp(1);
        // ---- fc1 / fc2 software pipeline
        for (int j = 0; j < NCHUNK + LAG; ++j) {
          if (j < NCHUNK) {
            const uint32_t u = it * NCHUNK + j;
            { const long long t_ = trace ? clock64() : 0;
              if (j == 0) {
                ptx::mbar_wait(xn_full, tph);
                ptx::mbar_wait(r_free, tph ^ 1);   // previous tile's R (this tile's XN/H half) fully drained
              }
              if (trace && j == 0) wt_tile += clock64() - t_; }
            { const long long t_ = trace ? clock64() : 0;
              ptx::mbar_wait(hacc_free, (u & 1) ^ 1);
              if (trace) wt_hacc += clock64() - t_; }
            ptx::tc_fence_after_sync();
            stamp(1);
            const long long ti1_ = trace ? clock64() : 0, ts1_ = wt_slot;   // issue time of the group, slot waits excluded
            for (int h2 = 0; h2 < 2; ++h2) {
              const uint32_t sw = slot_wait();
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int kb = 2 * h2 + q;
                const uint64_t db = ptx::smem_desc_k_sw128(sw + q * KB_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)   // A = xn from TMEM: 8 columns (16 bf16) per K step
                  ptx::umma_f16_ts_e(tmem_base + TM_H, tmem_base + TM_XN + (kb * 4 + k) * 8, db + 2 * k, idesc128, (kb | k) != 0);
              }
              slot_release();
            }
            ptx::umma_commit_e(hacc_full);
            if (trace) wt_i1 += (clock64() - ti1_) - (wt_slot - ts1_);
          }
          if (j >= LAG) {
            const int jj = j - LAG;
            const uint32_t b = jj & 1, u = it * 4 + (jj >> 1);
            { const long long t_ = trace ? clock64() : 0;
              ptx::mbar_wait(&hbuf_full[b], u & 1);
              if (trace) wt_hbuf += clock64() - t_; }
            ptx::tc_fence_after_sync();
            stamp(1);
            const long long ti2_ = trace ? clock64() : 0, ts2_ = wt_slot;
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t sw = slot_wait();
              const uint64_t da = ptx::smem_desc_k_sw128(sHB + b * HB_BYTES + kb * KB_BYTES);
              const uint64_t db = ptx::smem_desc_k_sw128(sw);
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_f16_e(tmem_base + TM_R, da + 2 * k, db + 2 * k, idesc256, 1u);
              slot_release();
            }
            ptx::umma_commit_e(&hbuf_free[b]);
            if (trace) wt_i2 += (clock64() - ti2_) - (wt_slot - ts2_);
          }
        }
        ptx::umma_commit_e(out_full);
        stamp(1);
      }
      if (trace && lane == 0) {   // where the issuing thread waited (row 0, slots 32..36): weights, gelu(h), H drain, tile-level, total
        trace[32] = wt_slot; trace[33] = wt_hbuf; trace[34] = wt_hacc; trace[35] = wt_tile; trace[36] = clock64() - t_all;
        trace[37] = wt_i1; trace[38] = wt_i2;   // cycles spent issuing the fc1 / fc2 groups (slot waits excluded)
      }
    }
  } else if (warp < EPI_WARP0) {
    ptx::setmaxnreg_dec<104>();   // idle warps of the first warpgroup
  } else {
    // =========================== epilogue warps ===========================
    ptx::setmaxnreg_inc<200>();
    const int e = warp - EPI_WARP0;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int hf = e >> 2;           // column half
    const int r = q * 32 + lane;     // row inside the tile
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t sHB = ptx::smem_u32(smem + OFF_HB);
    const uint32_t swz = uint32_t(r & 7);
    const LayerConsts& lc = c_mlp[p.layer];
    const bool tr = trace && warp == EPI_WARP0 && lane == 0;
    uint32_t it = 0;
    float4 rs[32];
    const float* const res_in = p.res_tab ? p.res_tab : p.res;
    // row of the residual INPUT: the row itself, or (block 0) the token id's row of the embedding table
    auto src_row = [&](long long row, bool ok) -> long long {
      if (!p.res_tab || !ok) return row;
      long long id;
      if (p.ids_dtype == 2) id = reinterpret_cast<const uint8_t*>(p.ids)[row];
      else if (p.ids_dtype == 3) id = reinterpret_cast<const int32_t*>(p.ids)[row];
      else id = reinterpret_cast<const long long*>(p.ids)[row];
      return (id < 0 || id >= p.vocab_rows) ? 0 : id;   // out-of-range ids are flagged by embed_in_kernel; same substitution
    };
    auto load_res = [&](long long lrow, bool ok, auto j0_, auto j1_) {
#pragma unroll
      for (int j = decltype(j0_)::value; j < decltype(j1_)::value; ++j)
        rs[j] = ok ? *reinterpret_cast<const float4*>(res_in + ptx::r32_off(lrow, hf * 128 + 4 * j))
                   : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // (lo, n) become constants once the calling loop is fully unrolled, so rs[] stays in registers
    auto load_res_dyn = [&](long long lrow, bool ok, int lo, int n) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j >= lo && j < lo + n)
          rs[j] = ok ? *reinterpret_cast<const float4*>(res_in + ptx::r32_off(lrow, hf * 128 + 4 * j))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    // residual / xn row of this thread in `tile`
    auto tile_row = [&](int tile, long long& row, bool& ok) {
      if (p.y_cm) {
        if (tile < p.n_full_tiles) {
          const int t = (tile % p.tiles_per_seq) * BM + r;
          row = (long long)(tile / p.tiles_per_seq) * p.T + t;
          ok = t < p.T;
        } else {
          const int jj = r / p.gather_L, j = (tile - p.n_full_tiles) * p.gather_P + jj;
          row = (long long)j * p.T + (p.T - p.gather_L) + (r - jj * p.gather_L);
          ok = jj < p.gather_P && j < p.B;
        }
      } else {
        row = (long long)tile * BM + r;
        ok = row < p.M;
      }
    };
    using I0 = std::integral_constant<int, 0>;
    using IE = std::integral_constant<int, (EARLY_RES > 32 ? 32 : EARLY_RES)>;
    using I32 = std::integral_constant<int, 32>;
    if (EARLY_RES && (int)blockIdx.x < p.num_tiles) {
      long long row0;
      bool ok0;
      tile_row((int)blockIdx.x, row0, ok0);
      load_res(src_row(row0, ok0), ok0, I0{}, IE{});
    }
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t tph = it & 1;
      const uint32_t TM_R = tph ? 256u : 0u, TM_XN = tph ? 0u : 256u, TM_H = tph ? 128u : 384u;
      long long row;
      bool row_ok;
      tile_row(tile, row, row_ok);
      // row of this thread in the NEXT tile of this CTA (for the residual L2 prefetch)
      long long pf_row = 0;
      bool pf_ok = false;
      if (tile + (int)gridDim.x < p.num_tiles) tile_row(tile + (int)gridDim.x, pf_row, pf_ok);
      // from here on pf_row is the SOURCE row of the next tile's residual (the id load is issued a whole tile before its use)
      pf_row = src_row(pf_row, pf_ok);
      // ------------------------------------------------ E1: r1, LayerNorm2 -> xn (TMEM)
      // The residual half-row (128 fp32) is fetched into registers BEFORE waiting for the
      // out_proj accumulator, so its DRAM latency hides behind the y-tile load and G1.
      if (IE::value < 32) load_res(src_row(row, row_ok), row_ok, IE{}, I32{});
      ptx::mbar_wait(g1_done, tph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(2);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int col = hf * 128 + ci * 32;
        uint32_t a[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + TM_R + col, a);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4& rr = rs[ci * 8 + j];
          const int cc = col + 4 * j;
          rr.x += __uint_as_float(a[4 * j + 0]) + lc.b_out[cc + 0];   // rs now holds r1
          rr.y += __uint_as_float(a[4 * j + 1]) + lc.b_out[cc + 1];
          rr.z += __uint_as_float(a[4 * j + 2]) + lc.b_out[cc + 2];
          rr.w += __uint_as_float(a[4 * j + 3]) + lc.b_out[cc + 3];
          s1 += (rr.x + rr.y) + (rr.z + rr.w);
          s2 += (rr.x * rr.x + rr.y * rr.y) + (rr.z * rr.z + rr.w * rr.w);
          a[4 * j + 0] = __float_as_uint(rr.x);   // r1 stays in TMEM as the fc2 accumulator's initial value
          a[4 * j + 1] = __float_as_uint(rr.y);
          a[4 * j + 2] = __float_as_uint(rr.z);
          a[4 * j + 3] = __float_as_uint(rr.w);
        }
        ptx::tmem_st_32x32b_x32(lane_addr + TM_R + col, a);
      }
      s_part[hf][0][r] = s1;
      s_part[hf][1][r] = s2;
      if (warp == EPI_WARP0 && lane == 0) ptx::tma_store_wait_read<0>();   // previous tile's xn store has finished reading HB
      ptx::bar_sync(1, EPI_THREADS);
      const float ts1 = s_part[0][0][r] + s_part[1][0][r];
      const float ts2 = s_part[0][1][r] + s_part[1][1][r];
      const float mean = ts1 * (1.0f / D);
      const float var = fmaxf(ts2 * (1.0f / D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      // xn = (r1 - mean) * rstd (gamma/beta folded into W1/b1), packed bf16 pairs: this thread's 128 columns
      // are 64 TMEM columns of the A operand
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t w[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 v = rs[hh * 16 + j];
          w[2 * j] = pack_bf16((v.x - mean) * rstd, (v.y - mean) * rstd);
          w[2 * j + 1] = pack_bf16((v.z - mean) * rstd, (v.w - mean) * rstd);
        }
        ptx::tmem_st_32x32b_x32(lane_addr + TM_XN + hf * 64 + hh * 32, w);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(xn_full);
      if (tr) stamp(2);
      // ------------------------------------------------ E2: gelu(fc1 chunk) -> HB
#pragma unroll 1
      for (int j = 0; j < NCHUNK; ++j) {
        const uint32_t b = j & 1, u = it * 4 + (j >> 1), uh = it * NCHUNK + j;
        ptx::mbar_wait(hacc_full, uh & 1);
        ptx::tc_fence_after_sync();
        if (tr) stamp(2);
        uint32_t a0[32], a1[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + TM_H + hf * 64, a0);
        ptx::tmem_ld_32x32b_x32(lane_addr + TM_H + hf * 64 + 32, a1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(hacc_free);   // the accumulator is in registers: fc1 of the next chunk may start
        if (j < 4 && pf_ok && !p.res_tab) {
          // next tile's residual half-row (512 B per thread, 32 float4 at 512 B stride): pull 8 of them towards L2 per
          // chunk so that E1 of the next tile does not start with a 19 MB chip-wide HBM burst
#pragma unroll
          for (int i = 0; i < 8; ++i)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + ptx::r32_off(pf_row, hf * 128 + 4 * (8 * j + i))));
        }
        const uint32_t rowaddr = sHB + b * HB_BYTES + hf * KB_BYTES + r * 128;
        // all the GELU math first (results in registers), THEN wait for fc2 of chunk j - 2 to release the HB buffer:
        // that MMA group sits behind fc1 of chunk j in the tensor pipe, ~1000 cycles after this epilogue may start
        uint32_t o[32];
        // (CLM_E2_DIAG = 1 / 2 / 3 compile timing-only variants without MUFU / without the tail arithmetic / without the bias
        // loads: they shorten this ~2.3 K-cycle epilogue by 250 / 0 / 350 cycles - no single piece dominates it.)
        const float* b1p = lc.b1 + ((j + rot) & (NCHUNK - 1)) * 128 + hf * 64;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t* src = (g < 4) ? &a0[g * 8] : &a1[(g - 4) * 8];
#pragma unroll
          for (int jj = 0; jj < 4; ++jj)
#if defined(CLM_E2_DIAG) && CLM_E2_DIAG == 3   // timing diagnostic only (wrong results): no bias loads
            o[4 * g + jj] = gelu_tanh_bf16x2(f2_add(f2_packu(src[2 * jj], src[2 * jj + 1]), f2_pack(0.25f, 0.125f)));
#else
            o[4 * g + jj] = gelu_tanh_bf16x2(f2_add(f2_packu(src[2 * jj], src[2 * jj + 1]), f2_pack(b1p[g * 8 + 2 * jj], b1p[g * 8 + 2 * jj + 1])));
#endif
        }
        ptx::mbar_wait(&hbuf_free[b], (u & 1) ^ 1);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t chunk = uint32_t(g) ^ swz;
          ptx::st_shared_v4(rowaddr + chunk * 16, o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&hbuf_full[b]);
        if (tr) stamp(2);
      }
      // ------------------------------------------------ E3: out = R + b2 -> res (+ normalised xn for the next consumer)
      // EARLY_RES == 33: 8 float4 here, 16 between the steps of the statistics sweep, 8 between the steps of the store sweep -
      // the loads are throttled at issue by the SM's in-flight limit (~17 KB), so one burst of 32 stalls ~4 K cycles here
      if constexpr (EARLY_RES == 33) load_res_dyn(pf_row, pf_ok, 0, 8);
      else if (EARLY_RES) load_res(pf_row, pf_ok, I0{}, IE{});   // rs is dead since E1; another CTA's rows, so E3's stores do not alias them
      ptx::mbar_wait(out_full, tph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(2);
      if (p.write_xn) {
        // Sweep A (TMEM reads only, no stores): row statistics of out = R + b2.  The row-statistics barrier then comes after
        // a cheap sweep instead of after the store-heavy one, whose slowest warp used to hold everybody.
        float o1 = 0.f, o2 = 0.f;
#pragma unroll(EARLY_RES == 33 ? 4 : 1)
        for (int ci = 0; ci < 4; ++ci) {
          const int col = hf * 128 + ci * 32;
          uint32_t a[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + TM_R + col, a);
          ptx::tmem_ld_wait();
          // store_a: the residual stores of the first `store_a` column groups are issued HERE, in the statistics sweep, where
          // the store path would otherwise idle (the values are final; only the normalised copy needs the row statistics):
          // 128 KB per tile at ~30 B/clk/SM is the floor of this epilogue, and sweep B alone carried all of it
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(a[j]) + lc.b2[col + j];
            o1 += v;
            o2 = fmaf(v, v, o2);
          }
          if (ci < p.store_a && row_ok && !p.skip_res_store) {
            float* dst = p.res + ptx::r32_off(row, col);   // column chunks of a row are 128 floats apart in the R32 layout
#pragma unroll
            for (int g = 0; g < 8; ++g)
              *reinterpret_cast<float4*>(dst + g * 128) =
                  make_float4(__uint_as_float(a[4 * g]) + lc.b2[col + 4 * g], __uint_as_float(a[4 * g + 1]) + lc.b2[col + 4 * g + 1],
                              __uint_as_float(a[4 * g + 2]) + lc.b2[col + 4 * g + 2], __uint_as_float(a[4 * g + 3]) + lc.b2[col + 4 * g + 3]);
          }
          if constexpr (EARLY_RES == 33) load_res_dyn(pf_row, pf_ok, 8 + 4 * ci, 4);
        }
        s_part[hf][0][r] = o1;
        s_part[hf][1][r] = o2;
        ptx::bar_sync(2, EPI_THREADS);
        const float m_ = (s_part[0][0][r] + s_part[1][0][r]) * (1.0f / D);
        const float v_ = fmaxf((s_part[0][1][r] + s_part[1][1][r]) * (1.0f / D) - m_ * m_, 0.f);
        const float rs_ = rsqrtf(v_ + p.eps);
        // Sweep B: residual stores (fp32, R32 layout) and the normalised bf16 row staged into HB (idle until the next tile's
        // first GELU chunk) as 4 k-blocks of [128 rows x 128 B], 128B-swizzled, for the TMA store - one pass.
#pragma unroll(EARLY_RES == 33 ? 4 : 1)
        for (int ci = 0; ci < 4; ++ci) {
          const int col = hf * 128 + ci * 32;
          uint32_t a[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + TM_R + col, a);
          ptx::tmem_ld_wait();
          const uint32_t rowaddr = sHB + (col >> 6) * KB_BYTES + r * 128;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(a[g * 8 + j]) + lc.b2[col + g * 8 + j];
            if (ci >= p.store_a && row_ok && !p.skip_res_store) {
              *reinterpret_cast<float4*>(p.res + ptx::r32_off(row, col + g * 8)) = make_float4(x[0], x[1], x[2], x[3]);
              *reinterpret_cast<float4*>(p.res + ptx::r32_off(row, col + g * 8 + 4)) = make_float4(x[4], x[5], x[6], x[7]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = (x[j] - m_) * rs_;
            const uint32_t chunk = uint32_t(((col & 63) >> 3) + g) ^ swz;
            ptx::st_shared_v4(rowaddr + chunk * 16, pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                              pack_bf16(x[6], x[7]));
          }
          if constexpr (EARLY_RES == 33) load_res_dyn(pf_row, pf_ok, 24 + 2 * ci, 2);
        }
        ptx::fence_proxy_async_smem();
      } else {
        if constexpr (EARLY_RES == 33) load_res_dyn(pf_row, pf_ok, 8, 24);   // (no xn output: unit-test / debug form)
#pragma unroll 1
        for (int ci = 0; ci < 4; ++ci) {
          const int col = hf * 128 + ci * 32;
          uint32_t a[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + TM_R + col, a);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 v = make_float4(__uint_as_float(a[4 * j]) + lc.b2[col + 4 * j], __uint_as_float(a[4 * j + 1]) + lc.b2[col + 4 * j + 1],
                                         __uint_as_float(a[4 * j + 2]) + lc.b2[col + 4 * j + 2], __uint_as_float(a[4 * j + 3]) + lc.b2[col + 4 * j + 3]);
            if (row_ok) *reinterpret_cast<float4*>(p.res + ptx::r32_off(row, col + 4 * j)) = v;
          }
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(r_free);
      if (p.write_xn) {
        ptx::bar_sync(1, EPI_THREADS);
        if (tile >= p.n_full_tiles) {   // (n_full_tiles == num_tiles unless tails are gathered)
          // rows of different reads: no TMA box.  Every thread copies its own staged half-row (the bytes it wrote above) to
          // the xn row, 16 B at a time; the generic-proxy reads are done before the next tile's first HB write (program order).
          if (row_ok) {
            uint4* dst = reinterpret_cast<uint4*>(p.xn + row * D + hf * 128);
#pragma unroll 1
            for (int k = 0; k < 16; ++k) {   // 16-byte chunk k of the half-row: k-block 2 hf + k / 8, chunk k % 8 (swizzled)
              const uint32_t a_ = sHB + (2 * hf + (k >> 3)) * KB_BYTES + r * 128 + ((uint32_t(k & 7) ^ swz) << 4);
              uint4 v;
              asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a_));
              dst[k] = v;
            }
          }
        } else if (warp == EPI_WARP0 && lane == 0) {
          int xb, xt0;
          if (p.y_cm) { xb = tile / p.tiles_per_seq; xt0 = (tile % p.tiles_per_seq) * BM; }
          else { xb = 0; xt0 = tile * BM; }
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) ptx::tma_store_3d(&tmXN, smem + OFF_HB + kb * KB_BYTES, kb * BK, xt0, xb);
          ptx::tma_store_commit();
        }
      }
      if (tr) stamp(2);
    }
    if (warp == EPI_WARP0 && lane == 0) ptx::tma_store_wait<0>();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace clm
