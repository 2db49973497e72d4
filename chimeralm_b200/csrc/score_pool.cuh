// Attention scorer + pooling partials in one persistent kernel (replaces gemm_bf16_tn_kernel<256,2,EPI_SCORE> +
// pool_partial_kernel on the fused path):
//   score[t] = w2 . gelu_erf(W0' xn[t] + b0') + b2            (BinarySequenceClassifier attention branch,
//   per 128-token tile: m = max score, p[t] = exp(score[t] - m),       components/hyena.py:79-95,117-132; ln_f's affine is
//   l = sum p, v[c] = sum_t p[t] xn[t][c]                               folded into W0', b0' and applied to v at the end)
// Each token tile is read ONCE from HBM (the unfused pair read the tokens twice), the tcgen05 accumulator is drained by
// 8 epilogue warps (the exact-erf GELU is the expensive part) and the softmax-weighted sum is taken from the same
// shared-memory tile the MMA consumed.  pool_merge_kernel combines the per-tile partials of a read.
// Pipelining (round 2): the first version kept the scorer weights (128 KB) resident, which left room for ONE token tile, so
// load -> MMA -> GELU -> softmax -> pooling of a tile ran strictly in series (17 K cycles per tile, 16 % of the HBM roofline).
// Now the weights stream through a two-slot ring (32 KB k-blocks, L2 hits: 128 KB per tile and SM), the token tile and the
// accumulator are double-buffered (2 x 64 KB, 2 x 256 TMEM columns), and tile i + 1 is loaded and multiplied while the
// epilogue warps score and pool tile i.
#pragma once
#include <cuda_bf16.h>

#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace clm {

struct ScorePoolParams {
  const float* b0;     // [256] folded scorer bias
  const float* w2;     // [256]
  float b2;
  const float* g;      // [256] ln_f gamma  (applied to the pooled sum: sum p (xn g + b) = g sum p xn + b sum p)
  const float* beta;   // [256] ln_f beta
  float* score;        // [B*T] (kept for clm_attention_weights and debugging)
  float* part;         // [B][tiles_per_seq][2 + 256]: (m, l, v)
  int B, T, tiles_per_seq, num_tiles;
  // Pooling of BinarySequenceClassifier (chimeralm/models/components/hyena.py:97-136, mask == None): 0 attention (what
  // ChimeraLM uses, lm.py:46-55), 1 mean = uniform weights, 2 max over the sequence per channel, 3 cls = position 0 only.
  // Mean and cls are the attention machinery with fixed scores; max keeps (max, min) per channel because ln_f's gamma may
  // be negative.
  int pool_mode;
};

namespace sp {
constexpr int D = 256, BM = 128, BK = 64;
constexpr int W_KB = D * BK * 2;            // 32 KB: [256 n x 64 k]
constexpr int A_KB = BM * BK * 2;           // 16 KB: [128 tokens x 64 k]
constexpr int NW = 2;                       // weight ring slots (one 64-wide k-block of all 256 outputs each)
constexpr int A_TILE = 4 * A_KB;            // 64 KB
constexpr int OFF_W = 0;
constexpr int OFF_A = OFF_W + NW * W_KB;    // 65536; two tile buffers
constexpr int OFF_BAR = OFF_A + 2 * A_TILE; // 196608
constexpr int OFF_F = OFF_BAR + 128;        // s_part[2][128], p[128], red[16]
constexpr int OFF_C = OFF_F + (2 * 128 + 128 + 16) * 4;   // b0[256], w2[256]: with ~200 KB of shared memory carved out the L1 is a
constexpr int SMEM_TOTAL = OFF_C + 512 * 4;               // few KB and warp-uniform __ldg loads go to L2 every time
constexpr int THREADS = 320;

// Exact-erf GELU on a packed pair, erf by Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, i.e. fp32 round-off level):
//   erf|z| = 1 - t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1 / (1 + 0.3275911 |z|),  z = x / sqrt(2)
//   gelu(x) = 0.5 x (1 + erf z) = 0.5 (x + |x|) - 0.5 |x| poly exp(-z^2)
// ~11 issue slots per element (2 of them MUFU) instead of ~30 for erff().
__device__ __forceinline__ f2t gelu_erf2(f2t x) {
  float x0, x1;
  f2_unpack(x, x0, x1);
  const f2t ax = f2_pack(fabsf(x0), fabsf(x1));
  const f2t z = f2_mul(ax, f2_pack(0.70710678118654752f, 0.70710678118654752f));
  float d0, d1;
  f2_unpack(f2_fma(z, f2_pack(0.3275911f, 0.3275911f), f2_pack(1.0f, 1.0f)), d0, d1);
  float t0, t1;   // single-MUFU forms: 1 ulp is far below the 1.5e-7 of the approximation itself
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
  const f2t t = f2_pack(t0, t1);
  f2t pl = f2_fma(t, f2_pack(1.061405429f, 1.061405429f), f2_pack(-1.453152027f, -1.453152027f));
  pl = f2_fma(pl, t, f2_pack(1.421413741f, 1.421413741f));
  pl = f2_fma(pl, t, f2_pack(-0.284496736f, -0.284496736f));
  pl = f2_fma(pl, t, f2_pack(0.254829592f, 0.254829592f));
  pl = f2_mul(pl, t);
  float q0, q1;
  f2_unpack(f2_mul(f2_mul(z, z), f2_pack(-1.4426950408889634f, -1.4426950408889634f)), q0, q1);
  float e0, e1;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(q0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(q1));
  const f2t pe = f2_mul(pl, f2_pack(e0, e1));
  const f2t half = f2_pack(0.5f, 0.5f);
  const f2t hax = f2_mul(ax, half);
  return f2_sub(f2_fma(x, half, hax), f2_mul(hax, pe));
}
}  // namespace sp

__global__ void __launch_bounds__(sp::THREADS, 1)
score_pool_kernel(const __grid_constant__ CUtensorMap tmXN, const __grid_constant__ CUtensorMap tmW, ScorePoolParams p) {
  using namespace sp;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  ptx::griddep_launch();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = bars;        // [NW] weight k-block landed
  uint64_t* w_empty = bars + 2;   // [NW] its MMAs are done
  uint64_t* a_full = bars + 4;    // [2] token tile landed
  uint64_t* a_empty = bars + 6;   // [2] pooling finished reading the tile (8 warp arrivals)
  uint64_t* acc_full = bars + 8;  // [2] accumulator complete
  uint64_t* acc_free = bars + 10; // [2] accumulator drained (8 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
  float* s_part = reinterpret_cast<float*>(smem + OFF_F);   // [2][128]
  float* s_p = s_part + 256;                                // [128]
  float* s_c = reinterpret_cast<float*>(smem + OFF_C);      // b0 | w2
  for (int i = threadIdx.x; i < 512; i += THREADS) s_c[i] = i < 256 ? __ldg(p.b0 + i) : __ldg(p.w2 + i - 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmXN); ptx::prefetch_tmap(&tmW);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1);
      ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 8);
      ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_free[i], 8);
    }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  ptx::griddep_wait();   // the normalised rows come from the last block's kernel

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0, wi = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * BM;
        const uint32_t buf = it & 1;
        ptx::mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);
        ptx::mbar_expect_tx(&a_full[buf], A_TILE);
        for (int kb = 0; kb < 4; ++kb)   // rows >= T: zeros
          ptx::tma_load_3d(smem + OFF_A + buf * A_TILE + kb * A_KB, &tmXN, &a_full[buf], kb * BK, t0, b);
        for (int kb = 0; kb < 4; ++kb, ++wi) {   // the scorer weights again, k-block by k-block (L2 hits)
          const uint32_t sl = wi % NW;
          ptx::mbar_wait(&w_empty[sl], ((wi / NW) & 1) ^ 1);
          ptx::mbar_expect_tx(&w_full[sl], W_KB);
          ptx::tma_load_2d(smem + OFF_W + sl * W_KB, &tmW, &w_full[sl], kb * BK, 0);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(BM, D);
      const uint32_t sA = ptx::smem_u32(smem + OFF_A), sW = ptx::smem_u32(smem + OFF_W);
      uint32_t it = 0, wi = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t buf = it & 1, bph = (it >> 1) & 1;
        ptx::mbar_wait(&a_full[buf], bph);
        ptx::mbar_wait(&acc_free[buf], bph ^ 1);
        ptx::tc_fence_after_sync();
        for (int kb = 0; kb < 4; ++kb, ++wi) {
          const uint32_t sl = wi % NW;
          ptx::mbar_wait(&w_full[sl], (wi / NW) & 1);
          ptx::tc_fence_after_sync();
          const uint64_t da = ptx::smem_desc_k_sw128(sA + buf * A_TILE + kb * A_KB), db = ptx::smem_desc_k_sw128(sW + sl * W_KB);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + buf * 256, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          ptx::umma_commit(&w_empty[sl]);
        }
        ptx::umma_commit(&acc_full[buf]);
      }
    }
  } else {
    const int e = warp - 2, q = warp & 3, hf = e >> 2;
    const int r = q * 32 + lane;            // token row inside the tile (scores); also used as lane index below
    const int tid = threadIdx.x - 64;       // 0..255: channel for the pooling sum
    const uint32_t lane_addr0 = tmem_base + (uint32_t(q * 32) << 16);
    const float gam = __ldg(p.g + tid), bet = __ldg(p.beta + tid);
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int b = tile / p.tiles_per_seq, ts = tile % p.tiles_per_seq, t0 = ts * BM;
      const int valid = min(BM, p.T - t0);
      const uint32_t buf = it & 1, bph = (it >> 1) & 1;
      const uint32_t lane_addr = lane_addr0 + buf * 256;
      ptx::mbar_wait(&acc_full[buf], bph);
      ptx::tc_fence_after_sync();
      // ---- scores: this thread's 128 of the 256 scorer outputs of row r
      f2t sc2 = f2_pack(0.f, 0.f);
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        const int col = hf * 128 + ci * 32;
        uint32_t a[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + col, a);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b4 = *reinterpret_cast<const float4*>(s_c + col + j);
          const float4 w4 = *reinterpret_cast<const float4*>(s_c + 256 + col + j);
          const f2t g0 = gelu_erf2(f2_add(f2_packu(a[j], a[j + 1]), f2_pack(b4.x, b4.y)));
          const f2t g1 = gelu_erf2(f2_add(f2_packu(a[j + 2], a[j + 3]), f2_pack(b4.z, b4.w)));
          sc2 = f2_fma(g0, f2_pack(w4.x, w4.y), sc2);
          sc2 = f2_fma(g1, f2_pack(w4.z, w4.w), sc2);
        }
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&acc_free[buf]);   // the accumulator may be overwritten by the tile after next
      float sc_lo, sc_hi;
      f2_unpack(sc2, sc_lo, sc_hi);
      s_part[hf * 128 + r] = sc_lo + sc_hi;
      ptx::bar_sync(1, 256);
      // ---- tile softmax partial (every warp computes the same m, l; warp 0 of the group publishes p[])
      float s0 = -INFINITY, s1 = -INFINITY, s2 = -INFINITY, s3 = -INFINITY;
      {
        const float x0 = s_part[lane] + s_part[128 + lane] + p.b2, x1 = s_part[32 + lane] + s_part[160 + lane] + p.b2;
        const float x2 = s_part[64 + lane] + s_part[192 + lane] + p.b2, x3 = s_part[96 + lane] + s_part[224 + lane] + p.b2;
        s0 = lane < valid ? x0 : -INFINITY;
        s1 = 32 + lane < valid ? x1 : -INFINITY;
        s2 = 64 + lane < valid ? x2 : -INFINITY;
        s3 = 96 + lane < valid ? x3 : -INFINITY;
        if (p.pool_mode == 1 || p.pool_mode == 2) {          // mean (and the bookkeeping of max): every position weighs the same
          s0 = lane < valid ? 0.f : -INFINITY;
          s1 = 32 + lane < valid ? 0.f : -INFINITY;
          s2 = 64 + lane < valid ? 0.f : -INFINITY;
          s3 = 96 + lane < valid ? 0.f : -INFINITY;
        } else if (p.pool_mode == 3) {                        // cls: only position 0 of the read
          s0 = (t0 == 0 && lane == 0) ? 0.f : -INFINITY;
          s1 = s2 = s3 = -INFINITY;
        }
      }
      float m = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
      for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      const float mz = (m == -INFINITY) ? 0.f : m;   // a tile with no weight at all (cls, tiles after the first): p = 0, l = 0
      const float p0 = __expf(s0 - mz), p1 = __expf(s1 - mz), p2 = __expf(s2 - mz), p3 = __expf(s3 - mz);   // exp(-inf) = 0
      float l = (p0 + p1) + (p2 + p3);
      for (int o = 16; o; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
      if (e == 0) {
        s_p[lane] = p0; s_p[32 + lane] = p1; s_p[64 + lane] = p2; s_p[96 + lane] = p3;
        const long long row = (long long)b * p.T + t0;
        if (lane < valid) p.score[row + lane] = s0;
        if (32 + lane < valid) p.score[row + 32 + lane] = s1;
        if (64 + lane < valid) p.score[row + 64 + lane] = s2;
        if (96 + lane < valid) p.score[row + 96 + lane] = s3;
      }
      ptx::bar_sync(2, 256);
      // ---- pooling: v[c] = sum_t p[t] xn[t][c] from the swizzled K-major tile the MMA just consumed
      {
        const int c = tid, kb = c >> 6, cc = c & 63;
        const uint8_t* base = smem + OFF_A + buf * A_TILE + kb * A_KB + (cc & 7) * 2;
        const uint32_t chunk = uint32_t(cc >> 3);
        float* out = p.part + ((long long)b * p.tiles_per_seq + ts) * (2 + D);
        if (p.pool_mode == 2) {   // max over the tile's valid positions of ln_f(x)[c] = gamma xn + beta
          float hi = -INFINITY, lo = INFINITY;
          for (int t = 0; t < valid; ++t) {
            const unsigned short h = *reinterpret_cast<const unsigned short*>(base + t * 128 + ((chunk ^ uint32_t(t & 7)) << 4));
            const float x = __uint_as_float(uint32_t(h) << 16);
            hi = fmaxf(hi, x);
            lo = fminf(lo, x);
          }
          out[2 + c] = (gam >= 0.f ? hi : lo) * gam + bet;
          if (c == 0) { out[0] = 0.f; out[1] = 1.f; }
        } else {
          float v = 0.f;
#pragma unroll 8
          for (int t = 0; t < BM; ++t) {
            const unsigned short h = *reinterpret_cast<const unsigned short*>(base + t * 128 + ((chunk ^ uint32_t(t & 7)) << 4));
            v = fmaf(s_p[t], __uint_as_float(uint32_t(h) << 16), v);
          }
          out[2 + c] = v * gam + l * bet;
          if (c == 0) { out[0] = m; out[1] = l; }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&a_empty[buf]);   // the tile buffer may be overwritten (s_part / s_p: the epilogue's own)
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace clm
