// block_mlp_kernel with SIXTEEN epilogue warps (four per scheduler) instead of eight: the clock trace of the 8-warp version
// (profiles/r1_trace_block_mlp.txt) shows its chunk phase bound by the GELU epilogue, whose two warps per scheduler stall
// on MUFU results together (1750 cycles for a 768-cycle fma / 1024-cycle MUFU floor), and ~21 K cycles per tile of
// load/store-latency-bound LayerNorm / output phases.  With 576 threads the register budget is 96 per thread, so the
// LayerNorm2 epilogue no longer holds its residual row across the row-statistics barrier (r1 is re-read from TMEM) and every
// thread owns a column QUARTER.  Producer, MMA issue order, TMEM plan and shared-memory layout are those of block_mlp.cuh,
// with a 4-slot weight ring (a 5-slot one measured no faster) to make room for the 4-way partial sums.
#pragma once
#include "block_mlp.cuh"
#include "longconv_tc.cuh"   // tc::tmem_ld16 / tmem_st8

namespace clm {
namespace tc {
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
}  // namespace tc

namespace bm16 {
constexpr int NSLOT = 4;
constexpr int OFF_W = bm::OFF_HB + 2 * bm::HB_BYTES;
constexpr int OFF_BAR = OFF_W + NSLOT * bm::SLOT_BYTES;
constexpr int OFF_PART = OFF_BAR + 256;                        // LayerNorm partial sums [4][2][128] fp32
constexpr int SMEM_TOTAL = OFF_PART + 4 * 2 * bm::BM * 4;
constexpr int THREADS = 576, EPI_THREADS = 512;
}  // namespace bm16

__global__ void __launch_bounds__(bm16::THREADS, 1)
block_mlp16_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmWout,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2,
                 const __grid_constant__ CUtensorMap tmXN, BlockMlpParams p) {
  using namespace bm;
  using bm16::NSLOT; using bm16::OFF_W; using bm16::OFF_BAR; using bm16::OFF_PART; using bm16::EPI_THREADS;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // swizzled UMMA/TMA tiles need 1 KB alignment
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
  float (*s_part)[2][BM] = reinterpret_cast<float (*)[2][BM]>(smem + OFF_PART);  // [column quarter][sum|sumsq][row]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Each CTA walks the 8 fc1/fc2 hidden chunks starting at a different one, so the 148 CTAs do
  // not all pull the same weight tile out of L2 at the same moment.
  const int rot = blockIdx.x & (NCHUNK - 1);
  // trace rows: 0 = producer, 1 = MMA issuer, 2 = epilogue warp 2; CTA 0 only
  long long* trace = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  int trace_n = 0;
  auto stamp = [&](int role) {
    if (trace && trace_n < 64) trace[role * 64 + trace_n++] = clock64();
  };

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmY); ptx::prefetch_tmap(&tmWout); ptx::prefetch_tmap(&tmW1); ptx::prefetch_tmap(&tmW2);
    ptx::prefetch_tmap(&tmXN);
    for (int i = 0; i < NSLOT; ++i) { ptx::mbar_init(&w_full[i], 1); ptx::mbar_init(&w_empty[i], 1); }
    ptx::mbar_init(g1_done, 1);
    ptx::mbar_init(xn_full, 16);
    ptx::mbar_init(hacc_full, 1); ptx::mbar_init(hacc_free, 16);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&hbuf_full[i], 16); ptx::mbar_init(&hbuf_free[i], 1); }
    ptx::mbar_init(out_full, 1); ptx::mbar_init(r_free, 16);
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  // All CTAs run the same tile schedule in lockstep, so their residual reads (E1) and writes (E3) would hit HBM as
  // chip-wide bursts while the tensor pipes idle.  Starting the CTAs in `stagger` phase groups spreads that traffic.
  if (p.stagger_cycles > 0) {
    const long long t_end = clock64() + (long long)(blockIdx.x & 3) * p.stagger_cycles;
    while (clock64() < t_end) {}
  }

  if (warp == 0) {
    // =========================== TMA producer ===========================
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
        // Ring order per tile: y(kb 0,1), Wout kb 0, Wout kb 1, y(kb 2,3), Wout kb 2, Wout kb 3, then the fc1/fc2 stream.  (With
        // both y slots first, Wout kb 3 would land on the slot that y(kb 2,3) only frees after the kb 3 MMAs: dead-lock
        // with a 4-slot ring.)  Weights are pre-tiled at finalize so that every 32 KB slot is ONE TMA instruction.
        for (int kp = 0; kp < 2; ++kp) {
          {
            uint8_t* s = slot_acquire();
            uint64_t* fb = &w_full[wi % NSLOT];
            for (int q = 0; q < 2; ++q) {
              const int kb = 2 * kp + q;
              if (p.y_cm) {
                const int b = tile / p.tiles_per_seq, t0 = (tile % p.tiles_per_seq) * BM;
                for (int hh = 0; hh < 2; ++hh)   // k-block = 64 channels; two 64-token halves of 8 KB each
                  ptx::tma_load_3d(s + q * KB_BYTES + hh * (KB_BYTES / 2), &tmY, fb, t0 + hh * 64, kb * BK, b);
              } else {
                ptx::tma_load_2d(s + q * KB_BYTES, &tmY, fb, kb * BK, tile * BM);
              }
            }
            ++wi;
            if (kp == 0) stamp(0);
          }
          for (int q = 0; q < 2; ++q) {                  // out_proj: k-block kb = rows [256 kb, +256)
            uint8_t* s = slot_acquire();
            ptx::tma_load_2d(s, &tmWout, &w_full[wi % NSLOT], 0, (2 * kp + q) * 256);
            ++wi;
          }
        }
        for (int j = 0; j <= NCHUNK; ++j) {
          if (j < NCHUNK) {  // fc1 chunk jc: k-blocks (2 h2, 2 h2 + 1) = rows [(4 jc + 2 h2) 128, +256)
            const int jc = (j + rot) & (NCHUNK - 1);
            for (int h2 = 0; h2 < 2; ++h2) {
              uint8_t* s = slot_acquire();
              ptx::tma_load_2d(s, &tmW1, &w_full[wi % NSLOT], 0, (jc * 4 + 2 * h2) * 128);
              ++wi;
            }
          }
          if (j >= 1) {  // fc2 K-chunk jj: k-blocks 2 jj + kb = rows [(2 jj + kb) 256, +256)
            const int jj = (j - 1 + rot) & (NCHUNK - 1);
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
    if (lane == 0) {
      constexpr uint32_t idesc256 = ptx::idesc_bf16_f32(BM, 256);
      constexpr uint32_t idesc256_amn = ptx::idesc_bf16_f32_amn(BM, 256);
      constexpr uint32_t idesc128 = ptx::idesc_bf16_f32(BM, 128);
      const uint32_t sHB = ptx::smem_u32(smem + OFF_HB);
      const uint32_t sW = ptx::smem_u32(smem + OFF_W);
      uint32_t wi = 0;
      long long wt_slot = 0, wt_hbuf = 0, wt_hacc = 0, wt_tile = 0, t_all = clock64();
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
        ptx::umma_commit(&w_empty[wi % NSLOT]);
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
        // ring order: y(kb 0,1), Wout kb 0, 1, y(kb 2,3), Wout kb 2, 3 - a y slot is released after its second k-block
        for (int kp = 0; kp < 2; ++kp) {
          const uint32_t sy = slot_wait();
          const uint32_t wi_y = wi;
          ++wi;
          for (int q2 = 0; q2 < 2; ++q2) {
            const int kb = 2 * kp + q2;
            const uint32_t sw = slot_wait();
            const uint64_t db = ptx::smem_desc_k_sw128(sw);
            const uint32_t sx = sy + q2 * KB_BYTES;
            if (p.y_cm) {
              // A = y^T tile: MN(token)-major, 2 atoms of 64 tokens 8 KB apart, K rows of 128 B; 16 K-rows per step
              const uint64_t da = ptx::smem_desc_mn_sw128(sx, KB_BYTES / 2, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_f16(tmem_base + TM_R, da + (2048 >> 4) * k, db + 2 * k, idesc256_amn, (kb | k) != 0);
            } else {
              const uint64_t da = ptx::smem_desc_k_sw128(sx);
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + TM_R, da + 2 * k, db + 2 * k, idesc256, (kb | k) != 0);
            }
            slot_release();
          }
          ptx::umma_commit(&w_empty[wi_y % NSLOT]);
        }
        ptx::umma_commit(g1_done);
        stamp(1);
        // ---- fc1 / fc2 software pipeline
        for (int j = 0; j <= NCHUNK; ++j) {
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
            for (int h2 = 0; h2 < 2; ++h2) {
              const uint32_t sw = slot_wait();
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int kb = 2 * h2 + q;
                const uint64_t db = ptx::smem_desc_k_sw128(sw + q * KB_BYTES);
#pragma unroll
                for (int k = 0; k < 4; ++k)   // A = xn from TMEM: 8 columns (16 bf16) per K step
                  ptx::umma_f16_ts(tmem_base + TM_H, tmem_base + TM_XN + (kb * 4 + k) * 8, db + 2 * k, idesc128, (kb | k) != 0);
              }
              slot_release();
            }
            ptx::umma_commit(hacc_full);
          }
          if (j >= 1) {
            const int jj = j - 1;
            const uint32_t b = jj & 1, u = it * 4 + (jj >> 1);
            { const long long t_ = trace ? clock64() : 0;
              ptx::mbar_wait(&hbuf_full[b], u & 1);
              if (trace) wt_hbuf += clock64() - t_; }
            ptx::tc_fence_after_sync();
            stamp(1);
            for (int kb = 0; kb < 2; ++kb) {
              const uint32_t sw = slot_wait();
              const uint64_t da = ptx::smem_desc_k_sw128(sHB + b * HB_BYTES + kb * KB_BYTES);
              const uint64_t db = ptx::smem_desc_k_sw128(sw);
#pragma unroll
              for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + TM_R, da + 2 * k, db + 2 * k, idesc256, 1u);
              slot_release();
            }
            ptx::umma_commit(&hbuf_free[b]);
          }
        }
        ptx::umma_commit(out_full);
        stamp(1);
      }
      if (trace) {   // where the issuing thread waited (row 0, slots 32..36): weights, gelu(h), H drain, tile-level, total
        trace[32] = wt_slot; trace[33] = wt_hbuf; trace[34] = wt_hacc; trace[35] = wt_tile; trace[36] = clock64() - t_all;
      }
    }
  } else {
    // =========================== epilogue warps (16: four per scheduler) ===========================
    const int e = warp - 2;
    const int q = warp & 3;          // TMEM lane quarter this warp may access
    const int cq = e >> 2;           // column quarter: 64 of the 256 model columns, 32 of the 128 columns of a hidden chunk
    const int r = q * 32 + lane;     // row inside the tile
    const uint32_t lane_addr = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t sHB = ptx::smem_u32(smem + OFF_HB);
    const uint32_t swz = uint32_t(r & 7);
    const LayerConsts& lc = c_mlp[p.layer];
    const bool tr = trace && warp == 2 && lane == 0;
    uint32_t it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t tph = it & 1;
      const uint32_t TM_R = tph ? 256u : 0u, TM_XN = tph ? 0u : 256u, TM_H = tph ? 128u : 384u;
      long long row;
      bool row_ok;
      if (p.y_cm) {
        const int b = tile / p.tiles_per_seq, t = (tile % p.tiles_per_seq) * BM + r;
        row = (long long)b * p.T + t;
        row_ok = t < p.T;
      } else {
        row = (long long)tile * BM + r;
        row_ok = row < p.M;
      }
      long long pf_row = 0;   // this thread's row in the NEXT tile of this CTA (residual L2 prefetch)
      bool pf_ok = false;
      {
        const int nt_ = tile + gridDim.x;
        if (nt_ < p.num_tiles) {
          if (p.y_cm) {
            const int t = (nt_ % p.tiles_per_seq) * BM + r;
            pf_row = (long long)(nt_ / p.tiles_per_seq) * p.T + t;
            pf_ok = t < p.T;
          } else {
            pf_row = (long long)nt_ * BM + r;
            pf_ok = pf_row < p.M;
          }
        }
      }
      // ------------------------------------------------ E1: r1 = acc + b_out + res -> TMEM, LayerNorm2 -> xn (TMEM)
      // the residual quarter-row (64 fp32) is fetched BEFORE waiting for the out_proj accumulator
      float4 rs[16];
#pragma unroll
      for (int j = 0; j < 16; ++j)
        rs[j] = row_ok ? *reinterpret_cast<const float4*>(p.res + ptx::r32_off(row, cq * 64 + 4 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
      ptx::mbar_wait(g1_done, tph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(2);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int col = cq * 64 + ci * 16;
        uint32_t a[16];
        tc::tmem_ld16(lane_addr + TM_R + col, a);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float4 rr = rs[ci * 4 + j];
          const int cc = col + 4 * j;
          rr.x += __uint_as_float(a[4 * j + 0]) + lc.b_out[cc + 0];
          rr.y += __uint_as_float(a[4 * j + 1]) + lc.b_out[cc + 1];
          rr.z += __uint_as_float(a[4 * j + 2]) + lc.b_out[cc + 2];
          rr.w += __uint_as_float(a[4 * j + 3]) + lc.b_out[cc + 3];
          s1 += (rr.x + rr.y) + (rr.z + rr.w);
          s2 += (rr.x * rr.x + rr.y * rr.y) + (rr.z * rr.z + rr.w * rr.w);
          a[4 * j + 0] = __float_as_uint(rr.x);   // r1 stays in TMEM: fc2 accumulates on top of it
          a[4 * j + 1] = __float_as_uint(rr.y);
          a[4 * j + 2] = __float_as_uint(rr.z);
          a[4 * j + 3] = __float_as_uint(rr.w);
        }
        tc::tmem_st16(lane_addr + TM_R + col, a);
      }
      s_part[cq][0][r] = s1;
      s_part[cq][1][r] = s2;
      if (threadIdx.x == 64) ptx::tma_store_wait_read<0>();   // previous tile's xn store has finished reading HB
      ptx::tmem_st_wait();                                     // r1 is re-read from TMEM below
      ptx::bar_sync(1, EPI_THREADS);
      const float ts1 = (s_part[0][0][r] + s_part[1][0][r]) + (s_part[2][0][r] + s_part[3][0][r]);
      const float ts2 = (s_part[0][1][r] + s_part[1][1][r]) + (s_part[2][1][r] + s_part[3][1][r]);
      const float mean = ts1 * (1.0f / D);
      const float var = fmaxf(ts2 * (1.0f / D) - mean * mean, 0.f);
      const float rstd = rsqrtf(var + p.eps);
      // xn = (r1 - mean) * rstd, packed bf16 pairs: this thread's 64 columns = 32 TMEM columns of the fc1 A operand
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        uint32_t a[16], w[8];
        tc::tmem_ld16(lane_addr + TM_R + cq * 64 + ci * 16, a);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          w[j] = pack_bf16((__uint_as_float(a[2 * j]) - mean) * rstd, (__uint_as_float(a[2 * j + 1]) - mean) * rstd);
        tc::tmem_st8(lane_addr + TM_XN + cq * 32 + ci * 8, w);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(xn_full);
      if (tr) stamp(2);
      // ------------------------------------------------ E2: gelu(fc1 chunk) -> HB   (32 of the chunk's 128 columns per thread)
#pragma unroll 1
      for (int j = 0; j < NCHUNK; ++j) {
        const uint32_t b = j & 1, u = it * 4 + (j >> 1), uh = it * NCHUNK + j;
        ptx::mbar_wait(hacc_full, uh & 1);
        ptx::tc_fence_after_sync();
        if (tr) stamp(2);
        uint32_t a0[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + TM_H + cq * 32, a0);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(hacc_free);   // the accumulator is in registers: fc1 of the next chunk may start
        if (j < 4 && pf_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.res + ptx::r32_off(pf_row, cq * 64 + 4 * (4 * j + i))));
        }
        const float* b1p = lc.b1 + ((j + rot) & (NCHUNK - 1)) * 128 + cq * 32;
        // hidden columns [32 cq, +32) of the chunk = k-block cq / 2, 16-byte chunks 4 (cq & 1) .. + 3 of the 128-byte row
        const uint32_t rowaddr = sHB + b * HB_BYTES + (cq >> 1) * KB_BYTES + r * 128;
        uint32_t o[16];
#pragma unroll
        for (int m = 0; m < 16; ++m)
          o[m] = gelu_tanh_bf16x2(f2_add(f2_packu(a0[2 * m], a0[2 * m + 1]), f2_pack(b1p[2 * m], b1p[2 * m + 1])));
        ptx::mbar_wait(&hbuf_free[b], (u & 1) ^ 1);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t chunk = uint32_t((cq & 1) * 4 + g) ^ swz;
          ptx::st_shared_v4(rowaddr + chunk * 16, o[4 * g], o[4 * g + 1], o[4 * g + 2], o[4 * g + 3]);
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&hbuf_full[b]);
        if (tr) stamp(2);
      }
      // ------------------------------------------------ E3: out = R + b2 -> res (+ normalised xn for the next consumer)
      ptx::mbar_wait(out_full, tph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(2);
      float o1 = 0.f, o2 = 0.f;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int col = cq * 64 + ci * 16;
        uint32_t a[16];
        tc::tmem_ld16(lane_addr + TM_R + col, a);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 v = make_float4(__uint_as_float(a[4 * j]) + lc.b2[col + 4 * j], __uint_as_float(a[4 * j + 1]) + lc.b2[col + 4 * j + 1],
                                       __uint_as_float(a[4 * j + 2]) + lc.b2[col + 4 * j + 2], __uint_as_float(a[4 * j + 3]) + lc.b2[col + 4 * j + 3]);
          if (row_ok) *reinterpret_cast<float4*>(p.res + ptx::r32_off(row, col + 4 * j)) = v;
          o1 += (v.x + v.y) + (v.z + v.w);
          o2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
      }
      if (p.write_xn) {
        s_part[cq][0][r] = o1;
        s_part[cq][1][r] = o2;
        ptx::bar_sync(2, EPI_THREADS);
        const float m_ = ((s_part[0][0][r] + s_part[1][0][r]) + (s_part[2][0][r] + s_part[3][0][r])) * (1.0f / D);
        const float v_ = fmaxf(((s_part[0][1][r] + s_part[1][1][r]) + (s_part[2][1][r] + s_part[3][1][r])) * (1.0f / D) - m_ * m_, 0.f);
        const float rs_ = rsqrtf(v_ + p.eps);
        // second sweep over R: normalise and stage into HB (idle until the next tile's first GELU chunk): this thread's 64
        // columns are exactly row r of k-block cq ([128 rows x 128 B], 128B-swizzled) for the TMA store
        const uint32_t rowaddr = sHB + cq * KB_BYTES + r * 128;
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int col = cq * 64 + ci * 16;
          uint32_t a[16];
          tc::tmem_ld16(lane_addr + TM_R + col, a);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = (__uint_as_float(a[g * 8 + j]) + lc.b2[col + g * 8 + j] - m_) * rs_;
            const uint32_t chunk = uint32_t(ci * 2 + g) ^ swz;
            ptx::st_shared_v4(rowaddr + chunk * 16, pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]),
                              pack_bf16(x[6], x[7]));
          }
        }
        ptx::fence_proxy_async_smem();
      }
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(r_free);
      if (p.write_xn) {
        ptx::bar_sync(1, EPI_THREADS);
        if (threadIdx.x == 64) {
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
    if (threadIdx.x == 64) ptx::tma_store_wait<0>();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace clm
