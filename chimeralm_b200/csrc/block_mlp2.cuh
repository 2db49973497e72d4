// CTA-pair (cta_group::2) version of the fused block tail (see block_mlp.cuh for the math and the
// single-CTA structure it grew out of).
//
// Why: in the single-CTA kernel the tensor core is starved by shared-memory bandwidth - an SS-mode
// 128x256x16 MMA reads 12 KB of operands per 128 cycles (96 B/clk of the SM's 128 B/clk) while TMA is
// writing the next weight tiles and the epilogue is writing GELU outputs into the same memory
// (profiles/r1_trace_block_mlp.txt: every MMA group takes ~2x its nominal time although
// profiles/probes/umma_rate_probe.cu shows the pipe at nominal rate in isolation).  With a CTA pair
// each SM holds its own 128 token rows (A operand, accumulators) but only HALF of every weight tile
// (B operand, split along N); the hardware exchanges the halves.  Weight bytes written by TMA and read
// by the MMA per SM are halved.
//
// Pair protocol (rank 0 = leader):
//   * both CTAs: TMA producer warp loads the CTA's own y tile and its half of each weight tile; all
//     completion bytes are credited to the LEADER's slot barrier (cp.async.bulk.tensor ... cta_group::2),
//   * leader only: one thread issues every tcgen05.mma.cta_group::2 and multicasts the commits to the
//     barriers at the same shared-memory offset in both CTAs (slot release, accumulator ready, ...),
//   * both CTAs: 8 epilogue warps work on the CTA's own TMEM rows; what they signal to the MMA thread
//     (xn written, H drained, GELU chunk written, R drained) arrives on the leader's barriers, remotely
//     from the peer (mapa + mbarrier.arrive.release.cluster), so those barriers count 16 warps.
#pragma once
#include "block_mlp.cuh"

namespace clm {
namespace bm2 {
using namespace bm;
constexpr int SLOT2_BYTES = KB_BYTES;            // 16 KB per CTA per slot (the pair's slot is 32 KB)
constexpr int NSLOT2 = 10;                       // 160 KB ring per CTA
constexpr int OFF_W2 = OFF_HB + 2 * HB_BYTES;    // 65536
constexpr int OFF_BAR2 = OFF_W2 + NSLOT2 * SLOT2_BYTES;   // 229376
constexpr int OFF_PART2 = OFF_BAR2 + 512;
constexpr int SMEM_TOTAL2 = OFF_PART2 + 2 * 2 * BM * 4;   // 231936 <= 232448
}  // namespace bm2

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(bm::THREADS, 1)
block_mlp2_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWout,
                  const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                  const __grid_constant__ CUtensorMap tmXN, BlockMlpParams p) {
  using namespace bm2;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR2);
  uint64_t* w_full = bars;                 // [10] leader's copy is the live one (tx bytes from both CTAs)
  uint64_t* w_empty = bars + 10;           // [10] per CTA, released by the leader's multicast commit
  uint64_t* g1_done = bars + 20;           // per CTA (multicast)
  uint64_t* xn_full = bars + 21;           // leader's copy, 16 arrivals
  uint64_t* hacc_full = bars + 22;         // per CTA (multicast)
  uint64_t* hacc_free = bars + 23;         // leader's copy, 16 arrivals
  uint64_t* hbuf_full = bars + 24;         // [2] leader's copy, 16 arrivals
  uint64_t* hbuf_free = bars + 26;         // [2] per CTA (multicast)
  uint64_t* out_full = bars + 28;          // per CTA (multicast)
  uint64_t* r_free = bars + 29;            // leader's copy, 16 arrivals
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 32);
  float (*s_part)[2][BM] = reinterpret_cast<float (*)[2][BM]>(smem + OFF_PART2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cr = ptx::cluster_ctarank();
  const bool leader = (cr == 0);
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_pair_tiles = (p.num_tiles + 1) >> 1;
  const int rot = pair & (NCHUNK - 1);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmY); ptx::prefetch_tmap(&tmWout); ptx::prefetch_tmap(&tmW1); ptx::prefetch_tmap(&tmW2);
    ptx::prefetch_tmap(&tmXN);
    for (int i = 0; i < NSLOT2; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    ptx::mbar_init(g1_done, 1);
    ptx::mbar_init(xn_full, 16);
    ptx::mbar_init(hacc_full, 1); ptx::mbar_init(hacc_free, 16);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&hbuf_full[i], 16); ptx::mbar_init(&hbuf_free[i], 1); }
    ptx::mbar_init(out_full, 1); ptx::mbar_init(r_free, 16);
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc_2cta<512>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();                     // peer's barriers are initialised before anything is signalled to them
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    if (lane == 0) {
      uint32_t wi = 0;
      // acquire ring slot: wait for the local slot to be free; the leader arms its barrier for the pair's 32 KB
      auto slot_acquire = [&](uint32_t& bar_addr) -> uint8_t* {
        const uint32_t s = wi % NSLOT2, ph = (wi / NSLOT2) & 1;
        ptx::mbar_wait(&w_empty[s], ph ^ 1);
        if (leader) ptx::mbar_expect_tx(&w_full[s], 2 * SLOT2_BYTES);
        bar_addr = ptx::mapa(ptx::smem_u32(&w_full[s]), 0);
        ++wi;
        return smem + OFF_W2 + s * SLOT2_BYTES;
      };
      for (int pt = pair; pt < n_pair_tiles; pt += npairs) {
        const int tile = 2 * pt + (int)cr;   // may be == num_tiles for the peer of the last pair: loads go out of bounds -> zeros
        const int yb = p.y_cm ? tile / p.tiles_per_seq : 0;
        const int yt0 = p.y_cm ? (tile % p.tiles_per_seq) * BM : tile * BM;
        for (int kb = 0; kb < 4; ++kb) {
          uint32_t ba;
          uint8_t* s = slot_acquire(ba);       // this CTA's y k-block [128 tokens x 64 channels]
          if (p.y_cm) {
            for (int hh = 0; hh < 2; ++hh) ptx::tma_load_3d_2cta(s + hh * (KB_BYTES / 2), &tmY, ba, yt0 + hh * 64, kb * BK, yb);
          } else {
            ptx::tma_load_2d_2cta(s, &tmY, ba, kb * BK, yt0);
          }
          s = slot_acquire(ba);                // this CTA's half (128 rows) of out_proj k-block kb
          ptx::tma_load_2d_2cta(s, &tmWout, ba, 0, kb * 256 + (int)cr * 128);
        }
        for (int j = 0; j <= NCHUNK; ++j) {
          if (j < NCHUNK) {   // fc1 chunk jc: this CTA's 64 rows of k-blocks (2 h2, 2 h2 + 1), 8 KB each
            const int jc = (j + rot) & (NCHUNK - 1);
            for (int h2 = 0; h2 < 2; ++h2) {
              uint32_t ba;
              uint8_t* s = slot_acquire(ba);
              for (int q = 0; q < 2; ++q)
                ptx::tma_load_2d_2cta(s + q * (KB_BYTES / 2), &tmW1, ba, 0, (jc * 4 + 2 * h2 + q) * 128 + (int)cr * 64);
            }
          }
          if (j >= 1) {       // fc2 K-chunk jj: this CTA's 128 rows of k-block 2 jj + kb
            const int jj = (j - 1 + rot) & (NCHUNK - 1);
            for (int kb = 0; kb < 2; ++kb) {
              uint32_t ba;
              uint8_t* s = slot_acquire(ba);
              ptx::tma_load_2d_2cta(s, &tmW2, ba, 0, (jj * 2 + kb) * 256 + (int)cr * 128);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc256 = ptx::idesc_bf16_f32(256, 256);
      constexpr uint32_t idesc256_amn = ptx::idesc_bf16_f32_amn(256, 256);
      constexpr uint32_t idesc128 = ptx::idesc_bf16_f32(256, 128);
      const uint32_t sHB = ptx::smem_u32(smem + OFF_HB);
      const uint32_t sW = ptx::smem_u32(smem + OFF_W2);
      uint32_t wi = 0;
      bool next_ready = false;
      uint32_t probed_wi = 0xffffffffu;
      auto slot_wait = [&]() -> uint32_t {
        const uint32_t s = wi % NSLOT2, ph = (wi / NSLOT2) & 1;
        if (!(probed_wi == wi && next_ready)) ptx::mbar_wait_cluster(&w_full[s], ph);
        ptx::tc_fence_after_sync();
        const uint32_t w1 = wi + 1;
        next_ready = ptx::mbar_try_wait(&w_full[w1 % NSLOT2], (w1 / NSLOT2) & 1);
        probed_wi = w1;
        return sW + s * SLOT2_BYTES;
      };
      auto slot_release = [&](uint32_t w) { ptx::umma_commit_2cta(&w_empty[w % NSLOT2]); };
      uint32_t it = 0;
      for (int pt = pair; pt < n_pair_tiles; pt += npairs, ++it) {
        const uint32_t tph = it & 1;
        const uint32_t TM_R = tph ? 256u : 0u, TM_XN = tph ? 0u : 256u, TM_H = tph ? 128u : 384u;
        // ---- G1: R = y * Wout^T   (ring order: y kb, Wout kb, ...)
        ptx::mbar_wait_cluster(hacc_free, ((it * NCHUNK) & 1) ^ 1);
        ptx::tc_fence_after_sync();
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t sy = slot_wait(); const uint32_t wy = wi++;
          const uint32_t sw = slot_wait(); const uint32_t ww = wi++;
          const uint64_t db = ptx::smem_desc_k_sw128(sw);
          if (p.y_cm) {
            const uint64_t da = ptx::smem_desc_mn_sw128(sy, KB_BYTES / 2, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16_2cta(tmem_base + TM_R, da + (2048 >> 4) * k, db + 2 * k, idesc256_amn, (kb | k) != 0);
          } else {
            const uint64_t da = ptx::smem_desc_k_sw128(sy);
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_f16_2cta(tmem_base + TM_R, da + 2 * k, db + 2 * k, idesc256, (kb | k) != 0);
          }
          slot_release(wy);
          slot_release(ww);
        }
        ptx::umma_commit_2cta(g1_done);
        // ---- fc1 / fc2 software pipeline
        for (int j = 0; j <= NCHUNK; ++j) {
          if (j < NCHUNK) {
            const uint32_t u = it * NCHUNK + j;
            if (j == 0) {
              ptx::mbar_wait_cluster(xn_full, tph);
              ptx::mbar_wait_cluster(r_free, tph ^ 1);
            }
            ptx::mbar_wait_cluster(hacc_free, (u & 1) ^ 1);
            ptx::tc_fence_after_sync();
            for (int h2 = 0; h2 < 2; ++h2) {
              const uint32_t sw = slot_wait(); const uint32_t ww = wi++;
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int kb = 2 * h2 + q;
                const uint64_t db = ptx::smem_desc_k_sw128(sw + q * (KB_BYTES / 2));   // this CTA's 64 rows of the k-block
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  ptx::umma_f16_ts_2cta(tmem_base + TM_H, tmem_base + TM_XN + (kb * 4 + k) * 8, db + 2 * k, idesc128, (kb | k) != 0);
              }
              slot_release(ww);
            }
            ptx::umma_commit_2cta(hacc_full);
          }
          if (j >= 1) {
            const int jj = j - 1;
            const uint32_t b = jj & 1, u = it * 4 + (jj >> 1);
            ptx::mbar_wait_cluster(&hbuf_full[b], u & 1);
            ptx::tc_fence_after_sync();
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t sw = slot_wait(); const uint32_t ww = wi++;
              const uint64_t da = ptx::smem_desc_k_sw128(sHB + b * HB_BYTES + kb * KB_BYTES);
              const uint64_t db = ptx::smem_desc_k_sw128(sw);
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_f16_2cta(tmem_base + TM_R, da + 2 * k, db + 2 * k, idesc256, 1u);
              slot_release(ww);
            }
            ptx::umma_commit_2cta(&hbuf_free[b]);
          }
        }
        ptx::umma_commit_2cta(out_full);
      }
    }
  } else {
    // =========================== epilogue warps (both CTAs, own rows) ===========================
    const int e = warp - 2;
    const int q = warp & 3;
    const int hf = e >> 2;
    const int r = q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t sHB = ptx::smem_u32(smem + OFF_HB);
    const uint32_t swz = uint32_t(r & 7);
    const LayerConsts& lc = c_mlp[p.layer];
    // signal the MMA thread: the barrier lives in the leader CTA
    auto arrive_leader = [&](uint64_t* bar) {
      if (leader) ptx::mbar_arrive(bar);
      else ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(bar), 0));
    };
    uint32_t it = 0;
    for (int pt = pair; pt < n_pair_tiles; pt += npairs, ++it) {
      const int tile = 2 * pt + (int)cr;
      const bool tile_ok = tile < p.num_tiles;
      const uint32_t tph = it & 1;
      const uint32_t TM_R = tph ? 256u : 0u, TM_XN = tph ? 0u : 256u, TM_H = tph ? 128u : 384u;
      long long row;
      bool row_ok;
      if (p.y_cm) {
        const int b = tile / p.tiles_per_seq, t = (tile % p.tiles_per_seq) * BM + r;
        row = (long long)b * p.T + t;
        row_ok = tile_ok && t < p.T;
      } else {
        row = (long long)tile * BM + r;
        row_ok = tile_ok && row < p.M;
      }
      // ------------------------------------------------ E1: r1, LayerNorm2 -> xn (TMEM)
      float4 rs[32];
#pragma unroll
      for (int j = 0; j < 32; ++j)
        rs[j] = row_ok ? *reinterpret_cast<const float4*>(p.res + ptx::r32_off(row, hf * 128 + 4 * j))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
      ptx::mbar_wait(g1_done, tph);
      ptx::tc_fence_after_sync();
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
          rr.x += __uint_as_float(a[4 * j + 0]) + lc.b_out[cc + 0];
          rr.y += __uint_as_float(a[4 * j + 1]) + lc.b_out[cc + 1];
          rr.z += __uint_as_float(a[4 * j + 2]) + lc.b_out[cc + 2];
          rr.w += __uint_as_float(a[4 * j + 3]) + lc.b_out[cc + 3];
          s1 += (rr.x + rr.y) + (rr.z + rr.w);
          s2 += (rr.x * rr.x + rr.y * rr.y) + (rr.z * rr.z + rr.w * rr.w);
          a[4 * j + 0] = __float_as_uint(rr.x);
          a[4 * j + 1] = __float_as_uint(rr.y);
          a[4 * j + 2] = __float_as_uint(rr.z);
          a[4 * j + 3] = __float_as_uint(rr.w);
        }
        ptx::tmem_st_32x32b_x32(lane_addr + TM_R + col, a);
      }
      s_part[hf][0][r] = s1;
      s_part[hf][1][r] = s2;
      if (threadIdx.x == 64) ptx::tma_store_wait_read<0>();
      ptx::bar_sync(1, EPI_THREADS);
      const float ts1 = s_part[0][0][r] + s_part[1][0][r];
      const float ts2 = s_part[0][1][r] + s_part[1][1][r];
      const float mean = ts1 * (1.0f / D);
      const float var = fmaxf(ts2 * (1.0f / D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
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
      if (lane == 0) arrive_leader(xn_full);
      // ------------------------------------------------ E2: gelu(fc1 chunk) -> HB
#pragma unroll 1
      for (int j = 0; j < NCHUNK; ++j) {
        const uint32_t b = j & 1, u = it * 4 + (j >> 1), uh = it * NCHUNK + j;
        ptx::mbar_wait(hacc_full, uh & 1);
        ptx::tc_fence_after_sync();
        uint32_t a0[32], a1[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + TM_H + hf * 64, a0);
        ptx::tmem_ld_32x32b_x32(lane_addr + TM_H + hf * 64 + 32, a1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) arrive_leader(hacc_free);
        ptx::mbar_wait(&hbuf_free[b], (u & 1) ^ 1);
        const float* b1p = lc.b1 + ((j + rot) & (NCHUNK - 1)) * 128 + hf * 64;
        const uint32_t rowaddr = sHB + b * HB_BYTES + hf * KB_BYTES + r * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const uint32_t* src = (g < 4) ? &a0[g * 8] : &a1[(g - 4) * 8];
          float x[8];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) x[jj] = gelu_tanh_fast(__uint_as_float(src[jj]) + b1p[g * 8 + jj]);
          const uint32_t chunk = uint32_t(g) ^ swz;
          ptx::st_shared_v4(rowaddr + chunk * 16, pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                            pack_bf16(x[6], x[7]));
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) arrive_leader(&hbuf_full[b]);
      }
      // ------------------------------------------------ E3: out = R + b2 -> res (+ normalised xn)
      ptx::mbar_wait(out_full, tph);
      ptx::tc_fence_after_sync();
      float o1 = 0.f, o2 = 0.f;
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
          o1 += (v.x + v.y) + (v.z + v.w);
          o2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
      }
      if (p.write_xn) {
        s_part[hf][0][r] = o1;
        s_part[hf][1][r] = o2;
        ptx::bar_sync(2, EPI_THREADS);
        const float m_ = (s_part[0][0][r] + s_part[1][0][r]) * (1.0f / D);
        const float v_ = fmaxf((s_part[0][1][r] + s_part[1][1][r]) * (1.0f / D) - m_ * m_, 0.f);
        const float rs_ = rsqrtf(v_ + p.eps);
#pragma unroll 1
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
            for (int j = 0; j < 8; ++j) x[j] = (__uint_as_float(a[g * 8 + j]) + lc.b2[col + g * 8 + j] - m_) * rs_;
            const uint32_t chunk = uint32_t(((col & 63) >> 3) + g) ^ swz;
            ptx::st_shared_v4(rowaddr + chunk * 16, pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                              pack_bf16(x[6], x[7]));
          }
        }
        ptx::fence_proxy_async_smem();
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) arrive_leader(r_free);
      if (p.write_xn) {
        ptx::bar_sync(1, EPI_THREADS);
        if (threadIdx.x == 64 && tile_ok) {
          int xb, xt0;
          if (p.y_cm) { xb = tile / p.tiles_per_seq; xt0 = (tile % p.tiles_per_seq) * BM; }
          else { xb = 0; xt0 = tile * BM; }
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) ptx::tma_store_3d(&tmXN, smem + OFF_HB + kb * KB_BYTES, kb * BK, xt0, xb);
          ptx::tma_store_commit();
        }
      }
    }
    if (threadIdx.x == 64) ptx::tma_store_wait<0>();
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
