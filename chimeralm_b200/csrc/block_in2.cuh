// CTA-pair (cta_group::2) form of the fused first half of a block (see block_in.cuh for the math and the layouts).
//
// Why a pair.  In the single-CTA kernel every MMA (M 128 x N 144 x K 16) fetches BOTH operands from shared memory: 4 KB of
// weights (A) and 4.6 KB of tokens (B) per 72 cycles of tensor work = 120 of the SM's 128 B/clk, before TMA fills and the
// epilogue's staging traffic - the clock trace (profiles/r2_trace_block_in.txt) shows a pass taking 5.2-6.2 K cycles to
// issue 3.5 K cycles of MMAs, plus 3.5-4 K cycles per tile waiting for the single-buffered 72 KB token tile, while the
// epilogue itself needs only ~4 K cycles per pass.  With M = 256 across two SMs each CTA supplies its own 128 weight rows
// and HALF of the token rows (72 of 144; the hardware exchanges the halves): 6.3 KB per MMA per SM, and the token tile
// shrinks to 36 KB per CTA, which leaves room to double-buffer it.
//
// Work split: the pair takes one 128-token tile at a time; CTA rank r computes channels [128 r, 128 r + 128) of each of
// the three groups (x0, x1, v) - exactly "pass h = r" of the single-CTA kernel, so the epilogue is unchanged.
//
// Pair protocol (rank 0 = leader), as in block_mlp2.cuh:
//   * both CTAs: the TMA producer loads the CTA's half of the token tile (72 rows) and its own weight slots; completion
//     bytes are credited to the LEADER's barriers (cp.async.bulk.tensor ... cta_group::2),
//   * leader only: the MMA warp issues every tcgen05.mma.cta_group::2 and multicasts the commits (slot release, token
//     tile release, accumulator ready) to the barriers at the same shared-memory offset in both CTAs,
//   * both CTAs: 8 epilogue warps drain the CTA's own TMEM rows; "accumulator set drained" arrives on the leader's
//     barrier (locally or through mapa + mbarrier.arrive.release.cluster), which therefore counts 16 warps.
#pragma once
#include "block_in.cuh"

namespace clm {
namespace bi2 {
using namespace bi;
constexpr int HALF = NCOL / 2;                                  // 72 token rows per CTA
constexpr int KB_HALF_BYTES = HALF * 128;                       // 9216: one k-block of this CTA's rows
constexpr int XN_HALF_BYTES = 4 * KB_HALF_BYTES;                // 36864
constexpr int NXN = 2;                                          // token tile double-buffered
// Weight ring: SEVEN slots of ONE k-block (16 KB: this CTA's 128 channels x 64 k).  The pair's slot round trip (both CTAs'
// TMA -> leader's barrier -> MMA -> multicast commit -> both producers) is longer than a single CTA's: with 3 x 32 KB the
// tile's 48 MMAs took 7-8 K cycles to issue, with 4 x 32 KB 3.9 K (profiles/r2_trace_block_in.txt); 4 x 32 KB does not fit
// next to two token buffers, 7 x 16 KB does.
constexpr int SLOT2_BYTES = 128 * 64 * 2;                       // 16 KB
constexpr int NSLOT2 = 7;
constexpr int OFF_XN2 = 0;
constexpr int OFF_W2 = OFF_XN2 + NXN * XN_HALF_BYTES;           // 73728
constexpr int OFF_STAGE2 = OFF_W2 + NSLOT2 * SLOT2_BYTES;       // 188416
constexpr int OFF_BAR2 = OFF_STAGE2 + 2 * STAGE_BOX;            // 221184
constexpr int SMEM_TOTAL2 = OFF_BAR2 + 256;
static_assert(SMEM_TOTAL2 <= 232448, "shared memory budget");
static_assert(KB_HALF_BYTES % 1024 == 0 && OFF_W2 % 1024 == 0 && OFF_STAGE2 % 1024 == 0, "swizzled tiles need 1 KB alignment");
}  // namespace bi2

namespace ptx {
// warp-uniform forms of the pair MMA / commit: the whole warp executes them, one elected lane issues (see umma_f16_e)
__device__ __forceinline__ void umma_f16_2cta_e(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta_e(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
}  // namespace ptx

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(bi::THREADS, 1)
block_in2_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmVX,
                 const __grid_constant__ CUtensorMap tmX0, const __grid_constant__ CUtensorMap tmXN, BlockInParams p) {
  using namespace bi2;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR2);
  uint64_t* w_full = bars;          // [NSLOT2] leader's copy is the live one (tx bytes from both CTAs)
  uint64_t* w_empty = bars + 8;     // [NSLOT2] per CTA, released by the leader's multicast commit
  uint64_t* xn_full = bars + 16;    // [NXN]   leader's copy (both halves' bytes)
  uint64_t* xn_free = bars + 18;    // [NXN]   per CTA (multicast commit)
  uint64_t* acc_full = bars + 20;   // [2]     per CTA (multicast commit): set A (x0) / set B (x1, v) accumulated
  uint64_t* acc_free = bars + 22;   // [2]     leader's copy, 16 warp arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cr = ptx::cluster_ctarank();
  const bool leader = (cr == 0);
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  long long* trace = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  int trace_n = 0;
  auto stamp = [&](int role) {
    if (trace && (threadIdx.x & 31) == 0 && trace_n < 64) trace[role * 64 + trace_n++] = clock64();
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmW); ptx::prefetch_tmap(&tmVX); ptx::prefetch_tmap(&tmX0); ptx::prefetch_tmap(&tmXN);
    for (int i = 0; i < NSLOT2; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    for (int i = 0; i < NXN; ++i) { ptx::mbar_init(&xn_full[i], 1); ptx::mbar_init(&xn_free[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_free[i], 16); }
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc_2cta<512>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();                     // the peer's barriers exist before anything is signalled to them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (lane == 0) {
      uint32_t wi = 0, it = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs, ++it) {
        {  // this CTA's half of the token tile: normalised tokens [t0 - 16 + 72 r, + 72) of read b, 4 k-blocks of [72 x 64];
           // rows outside [0, T) are zero-filled by TMA (their products are discarded by the epilogue)
          const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * BT;
          const uint32_t buf = it % NXN, ph = (it / NXN) & 1;
          ptx::mbar_wait(&xn_free[buf], ph ^ 1);
          if (leader) ptx::mbar_expect_tx(&xn_full[buf], 2 * XN_HALF_BYTES);
          const uint32_t ba = ptx::mapa(ptx::smem_u32(&xn_full[buf]), 0);
          for (int kb = 0; kb < 4; ++kb)
            ptx::tma_load_3d_2cta(smem + OFF_XN2 + buf * XN_HALF_BYTES + kb * KB_HALF_BYTES, &tmXN, ba, kb * 64,
                                  t0 - HALO + HALF * (int)cr, b);
        }
        for (int g = 0; g < 3; ++g)
          for (int kb = 0; kb < 4; ++kb) {   // this CTA's 128 channels of group g, k-block kb: one 16 KB box
            const uint32_t s = wi % NSLOT2, ph = (wi / NSLOT2) & 1;
            ptx::mbar_wait(&w_empty[s], ph ^ 1);
            if (leader) ptx::mbar_expect_tx(&w_full[s], 2 * SLOT2_BYTES);
            ptx::tma_load_2d_2cta(smem + OFF_W2 + s * SLOT2_BYTES, &tmW, ptx::mapa(ptx::smem_u32(&w_full[s]), 0), 0,
                                  ((g * 2 + (int)cr) * 4 + kb) * 128);
            ++wi;
          }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only; whole warp, one elected lane issues) =====================
    if (leader) {
      constexpr uint32_t idesc = ptx::idesc_bf16_f32(256, NCOL);
      const uint32_t sXN = ptx::smem_u32(smem + OFF_XN2), sW = ptx::smem_u32(smem + OFF_W2);
      uint32_t wi = 0, it = 0;
      for (int tile = pair; tile < p.num_tiles; tile += npairs, ++it) {
        const uint32_t buf = it % NXN;
        stamp(0);
        ptx::mbar_wait_cluster(&xn_full[buf], (it / NXN) & 1);
        stamp(0);
        for (int g = 0; g < 3; ++g) {
          if (g < 2) {   // g == 0 starts set A, g == 1 starts set B
            ptx::mbar_wait_cluster(&acc_free[g], (it & 1) ^ 1);
            ptx::tc_fence_after_sync();
            if (g == 0) stamp(0);
          }
          for (int kb = 0; kb < 4; ++kb) {
            const uint32_t s = wi % NSLOT2, ph = (wi / NSLOT2) & 1;
            ptx::mbar_wait_cluster(&w_full[s], ph);
            ptx::tc_fence_after_sync();
            const uint64_t da = ptx::smem_desc_k_sw128(sW + s * SLOT2_BYTES);
            const uint64_t db = ptx::smem_desc_k_sw128(sXN + buf * XN_HALF_BYTES + kb * KB_HALF_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16_2cta_e(tmem_base + g * GCOLS, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
            ptx::umma_commit_2cta_e(&w_empty[s]);
            ++wi;
          }
          if (g == 0) ptx::umma_commit_2cta_e(&acc_full[0]);
          if (g == 2) ptx::umma_commit_2cta_e(&acc_full[1]);
        }
        ptx::umma_commit_2cta_e(&xn_free[buf]);
        stamp(0);
      }
    }
  } else {
    // =========================== epilogue (both CTAs, own channels) ===========================
    const int e = warp - 2;
    const int q = warp & 3;            // TMEM lane quarter
    const int hf = e >> 2;             // token half of the tile
    const int r = q * 32 + lane;       // channel inside this CTA's 128
    const int h = (int)cr;             // which 128-channel half of each group this CTA owns
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t sST = ptx::smem_u32(smem + OFF_STAGE2);
    const bool issuer = (threadIdx.x == 64);
    auto arrive_leader = [&](uint64_t* bar) {   // the MMA thread's barriers live in the leader CTA
      if (leader) ptx::mbar_arrive(bar);
      else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), 0));
    };
    float bia[3], w0[3], w1[3], w2[3], cbv[3];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int ch = g * 256 + h * 128 + r;
      bia[g] = __ldg(p.b_in + ch);
      w0[g] = __ldg(p.cw + ch * 3);
      w1[g] = __ldg(p.cw + ch * 3 + 1);
      w2[g] = __ldg(p.cw + ch * 3 + 2);
      cbv[g] = __ldg(p.cb + ch);
    }
    if (p.vx_scale) {   // fold a[ch] into the v group's short-filter taps and bias: a * v exactly, no extra work
      const float a = __ldg(p.vx_scale + h * 128 + r);
      w0[2] *= a; w1[2] *= a; w2[2] *= a; cbv[2] *= a;
    }
    uint32_t it = 0;
    for (int tile = pair; tile < p.num_tiles; tile += npairs, ++it) {
      const int b = tile / p.tiles_per_seq;
      const int t0 = (tile % p.tiles_per_seq) * BT;
      const bool tr = trace && warp == 2 && lane == 0;
      if (tr) stamp(1);
      // staging buffers may still be read by the previous tile's TMA stores
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
      // ---- set A: x0
      ptx::mbar_wait(&acc_full[0], it & 1);
      ptx::tc_fence_after_sync();
      if (tr) stamp(1);
      halo(0);
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        float x0v[32];
        conv_sub(0, s, x0v);
        if (s == 1) {                      // set A is in registers: the next tile may overwrite it
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) arrive_leader(&acc_free[0]);
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
      if (tr) stamp(1);
      if (issuer) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) ptx::tma_store_3d(&tmX0, smem + OFF_STAGE2 + hh * STAGE_BOX, t0 + hh * 64, h * 128, b);
        ptx::tma_store_commit();
      }
      // ---- set B: v * x1
      ptx::mbar_wait(&acc_full[1], it & 1);
      ptx::tc_fence_after_sync();
      if (tr) stamp(1);
      halo(1);
      halo(2);
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        float x1v[32], vv[32];
        conv_sub(1, s, x1v);
        conv_sub(2, s, vv);
        if (s == 1) {
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) arrive_leader(&acc_free[1]);
        }
        if (s == 0) {   // the staging buffers are being read by the x0 store of this tile (long since issued)
          if (issuer) ptx::tma_store_wait_read<0>();
          ptx::bar_sync(1, EPI_THREADS);
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
        for (int hh = 0; hh < 2; ++hh) ptx::tma_store_3d(&tmVX, smem + OFF_STAGE2 + hh * STAGE_BOX, t0 + hh * 64, h * 128, b);
        ptx::tma_store_commit();
      }
      if (tr) stamp(1);
    }
    if (issuer) ptx::tma_store_wait<0>();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();        // both CTAs are done with each other's shared memory, barriers and TMEM
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_2cta<512>(tmem_base);
  }
}

}  // namespace clm
