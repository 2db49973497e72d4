// Tensor-core FFT long convolution with TWO items in flight per SM (reads of 2057..8200 tokens, one transform per item).
// Same math, operand forms and accuracy as longconv_tc_kernel<false> (longconv_tc.cuh: steps 1/3/5/7 on tcgen05, epilogue
// phases E1..E4); what changes is the schedule.  There one item owns all 512 TMEM columns and its four matrix steps and
// four epilogue phases run in series (tensor pipe ~45 % active).  Here every matrix product is issued as two N = 128 halves
// (re | im, the sign of the one negative term of a complex product comes from the instruction descriptor's negate bit, so
// the constant stack shrinks to [Re F; Im F] = 64 KB and a third z buffer fits), which makes a 128-column UNIT the
// granule of TMEM, and two items, three (then five) slots apart, share the four units U0..U3:
//
//   slot         0        1        2        3        4        5        6        7
//   tensor    A.M1     B-.M7    A.M3     B.M1     A.M5     B.M3     A.M7     B.M5        (A = item 2k, B = item 2k+1,
//   epilogue  B-.E3    A.E1     B-.E4    A.E2     B.E1     A.E3     B.E2     A.E4         B- = item 2k-1)
//
//   U0: A.A_re -> A.P1 | B.A_re -> B.P1 | A.z'                 U2: B-.z' | B.A_im (E1 drains it first) | A.B_im | B.B_im
//   U1: A.A_im | A.S_re -> A.P2 | B.S_re -> B.P2               U3: A.S_im | A.B_re (E3 drains it first) | B.S_im | B.B_re
//
// Two hand-overs need a unit while the other item's epilogue still reads it: E1 therefore pulls its im unit (64 columns per
// thread) and E3 its B_re unit into registers FIRST and signals (`imd`, `bdr`); everything else is ordered by the in-order
// tensor pipe or by the barriers the data flow needs anyway.  The tail tokens (t >= 8192, at most LONGCONV_TAIL_MAX) no
// longer cost an extra output row in step 7: F[k1][64] = (-1)^k1, so their transform part is an alternating sum over BT,
// done together with the direct products by an otherwise idle warp, which also issues the gate-tile loads.
//
// Warps: 0..7 = epilogue (TMEM lane quarter = warp % 4, index half = warp / 4), 8 = TMA producer (z tiles, output stores),
// 9 = MMA issuer, 10 = gate loads, 11 = tail tokens; `setmaxnreg` moves registers to the epilogue warpgroups.
#pragma once
#include "longconv_tc.cuh"

namespace clm {
namespace tc2 {
using namespace tc;
constexpr int S2_PANEL = 256 * 128;               // bytes per 64-wide K panel of [Re F; Im F]
constexpr int S2_BYTES = 2 * S2_PANEL;            // 65536
constexpr int NZ = 3;                             // z buffers: z tile -> gate tile -> output tile of one item
constexpr int OFF2_S = 0;
constexpr int OFF2_Z = OFF2_S + S2_BYTES;
constexpr int OFF2_BT = OFF2_Z + NZ * Z_BYTES;
constexpr int OFF2_BAR = OFF2_BT + BT_BYTES;      // 229376
constexpr int SMEM2_TOTAL = OFF2_BAR + 1024;
constexpr int THREADS2 = 384;
constexpr int FRE = 0, FIM = 128;                 // row blocks of the constant stack
constexpr int EPI2_W0 = 0;                        // epilogue warps 0..7
// The helper warps take the HIGHEST warp ids: the warp scheduler favours higher ids, and the MMA issuer (a few instructions
// per 64-cycle MMA) must not queue behind the two epilogue warps of its scheduler.
constexpr int W_PROD = 8, W_MMA = 9, W_AUX = 10, W_TAIL = 11;

// kind::f16 instruction descriptor, fp16 x fp16 -> fp32, M = 128, optional negation of the A / B operand (bits 13 / 14)
__host__ __device__ constexpr uint32_t idesc2(uint32_t n, bool a_mn, bool b_mn, bool a_neg, bool b_neg) {
  return (1u << 4) | (a_neg ? (1u << 13) : 0u) | (b_neg ? (1u << 14) : 0u) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) |
         ((n >> 3) << 17) | ((128u >> 4) << 24);
}
}  // namespace tc2

// TR: record the clock trace of CTA 0 (clm_longconv_tc_trace); the product launch carries no stamps.
// P4: reads of at most 4096 tokens, FOUR reads per item.  A causal convolution of a T-token read needs taps 0..T-1 only,
// so with the filter truncated to 4096 taps (its own spectrum table) a 16384-point transform has room for two reads per
// real sequence: [read 0 | 4096 zeros | read 1 | 4096 zeros] convolves to [out 0 (8192) | out 1 (8192)] without overlap.
// z = (reads 4q, 4q+1) + i (reads 4q+2, 4q+3).  The nonzero input rows are n1 in [0,32) and [64,96) (still K = 64 in step 1:
// the constants' K slices jump to the second panel), the wanted output rows are the same two runs (step 7: two N = 32
// groups per half instead of one N = 64), and the gate / output tile holds 32 rows of each of the four reads.  Everything
// between (steps 3 and 5, E1-E3) is unchanged, so the conv cost per read halves for the short-read buckets.
// CH: reads longer than 8 200 tokens, overlap-add over chunks of 8 192 tokens in the frequency domain.  A work unit is
// (item, chunk); the two units in flight are consecutive chunks of one item (or the last / first of two).  Chunk c's
// spectrum S_c is parked (fp16 pairs, per-CTA scratch, each thread re-reads only what it wrote) and E2 forms
//   V_c = sum_{j <= c} S_{c-j} H_j,   H_j = FFT([k_j | k_{j-1}])   (V-form tables, tc::spectrum_kernel)
// so that the chunk's outputs are the FIRST half of one inverse transform: no second output half, no carry between chunks,
// and steps 1 / 7, E1, E3, E4 are exactly the single-transform ones.  Tail tokens (t >= NC C): see the tail warp.
template <bool TR, bool P4, bool CH>
__global__ void __launch_bounds__(tc2::THREADS2, 1)
longconv_tc2_kernel(const __grid_constant__ CUtensorMap tmVX, const __grid_constant__ CUtensorMap tmOut,
                    const __grid_constant__ CUtensorMap tmX0, LongConvTcParams p) {
  using namespace tc2;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  ptx::griddep_launch();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF2_BAR);
  uint64_t* z_full = bars;           // [3] z tile landed (TMA)
  uint64_t* g_full = bars + 3;       // [3] x0 gate tile landed in the z buffer
  uint64_t* out_ready = bars + 6;    // [3] E4 wrote the output tile (8 warp arrivals)
  uint64_t* x_full = bars + 9;       // [type] step 1 complete
  uint64_t* y_full = bars + 11;      // [type] step 3 complete
  uint64_t* x2_full = bars + 13;     // [type] step 5 complete
  uint64_t* o_full = bars + 15;      // [type] step 7 complete
  uint64_t* p1_full = bars + 17;     // [type][index half] E1 wrote P1
  uint64_t* p2_full = bars + 21;     // [type][index half] E2 wrote P2
  uint64_t* bt_full = bars + 25;     // [type] E3 wrote BT
  uint64_t* e4_done = bars + 27;     // [type] E4 has read z' out of TMEM
  uint64_t* imd = bars + 29;         // E1 of a B-type item has pulled A_im (unit 2) into registers
  uint64_t* bdr = bars + 30;         // E3 of an A-type item has pulled B_re (unit 3) into registers
  uint64_t* bt_read = bars + 31;     // [type] the tail warp has finished reading BT
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 40);

  // `warp` is the ROLE index (0..7 epilogue, 8..11 helpers); the TMEM lane quarter (physical warp % 4) equals role % 4 either way
  const int pwarp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = p.helpers_low ? (pwarp + 8) % 12 : pwarp;
  const int per = (p.n_items + gridDim.x - 1) / gridDim.x;
  const int item0 = blockIdx.x * per, item1 = min(p.n_items, item0 + per);
  const int NC = CH ? p.n_chunks : 1;
  const int n = max(0, item1 - item0) * NC;   // work units of this CTA: (local item, chunk), chunk fastest
  const int inv_nc = 65536 / NC + 1;          // u / NC = (u * inv_nc) >> 16 for u < 4096, NC <= 4
  auto u_item = [&](int u) { return CH ? (u * inv_nc) >> 16 : u; };
  auto u_chunk = [&](int u) { return CH ? u - ((u * inv_nc) >> 16) * NC : 0; };
  const int nt = p.nt;
  float* tailF = reinterpret_cast<float*>(bars + 48);   // [NZ][2 reads][8]: first-row outputs of a unit, for the tail warp (CH)
  long long* trace = (TR && p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  int trace_n = 0;
  auto stamp = [&](int role) {
    if (TR && trace && (threadIdx.x & 31) == 0 && trace_n < 64) trace[role * 64 + trace_n++] = clock64();
  };

  // constant stack [Re F; Im F] -> shared memory (rows 128..383 of each panel of the 96 KB image)
  {
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF2_S);
    for (int i = threadIdx.x; i < S2_BYTES / 16; i += THREADS2) {
      const int pnl = i / (S2_PANEL / 16), off = i % (S2_PANEL / 16);
      dst[i] = __ldg(p.S + pnl * (S_PANEL / 16) + 128 * 8 + off);
    }
  }
  if (warp == W_PROD && lane == 0) {
    ptx::prefetch_tmap(&tmVX); ptx::prefetch_tmap(&tmOut); ptx::prefetch_tmap(&tmX0);
    for (int i = 0; i < 3; ++i) { ptx::mbar_init(&z_full[i], 1); ptx::mbar_init(&g_full[i], 1); ptx::mbar_init(&out_ready[i], 8); }
    for (int t = 0; t < 2; ++t) {
      ptx::mbar_init(&x_full[t], 1); ptx::mbar_init(&y_full[t], 1); ptx::mbar_init(&x2_full[t], 1); ptx::mbar_init(&o_full[t], 1);
      ptx::mbar_init(&p1_full[2 * t], 8); ptx::mbar_init(&p1_full[2 * t + 1], 8);
      ptx::mbar_init(&p2_full[2 * t], 8); ptx::mbar_init(&p2_full[2 * t + 1], 8);
      ptx::mbar_init(&bt_full[t], 8); ptx::mbar_init(&e4_done[t], 8); ptx::mbar_init(&bt_read[t], 1);
    }
    ptx::mbar_init(imd, 8); ptx::mbar_init(bdr, 8);
    ptx::fence_mbar_init();
  } else if (warp == W_MMA) {
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t U0 = tmem_base, U1 = tmem_base + 128, U2 = tmem_base + 256, U3 = tmem_base + 384;
  ptx::griddep_wait();   // vx and x0 come from the kernel before this one (the 64 KB constant stack above does not)

  if (warp == W_PROD) {
    // =========================== TMA producer (z loads, output stores) ===========================
    ptx::setmaxnreg_dec<104>();
    if (lane == 0) {
      auto load_z = [&](int it) {
        const int buf = it % NZ, item = item0 + u_item(it), c = u_chunk(it);
        const int ch = item / p.n_pairs, pr = item % p.n_pairs;
        uint8_t* z = smem + OFF2_Z + buf * Z_BYTES;
        ptx::mbar_expect_tx(&z_full[buf], Z_BYTES);
        for (int part = 0; part < 2; ++part) {
          if constexpr (P4) {   // K rows 0..31: first read of the part (n1 = 0..31), K rows 32..63: second read (n1 = 64..95)
            for (int sl = 0; sl < 2; ++sl) {
              const int row = (4 * pr + 2 * part + sl) * p.D + ch;   // reads past B are out of bounds -> zero filled
              for (int a = 0; a < 2; ++a)
                ptx::tma_load_3d(z + part * 16384 + a * 8192 + sl * 4096, &tmVX, &z_full[buf], 64 * a, 0, row);
            }
          } else {
            const int row = (2 * pr + part) * p.D + ch;    // reads past B are out of bounds -> zero filled
            for (int a = 0; a < 2; ++a) ptx::tma_load_3d(z + part * 16384 + a * 8192, &tmVX, &z_full[buf], 64 * a, 64 * c, row);
          }
        }
      };
      for (int it = 0; it < min(n, NZ); ++it) load_z(it);
      for (int j = 0; j < n; ++j) {
        const int buf = j % NZ, item = item0 + u_item(j), c = u_chunk(j);
        const int ch = item / p.n_pairs, pr = item % p.n_pairs;
        uint8_t* z = smem + OFF2_Z + buf * Z_BYTES;
        ptx::mbar_wait(&out_ready[buf], (j / NZ) & 1);
        if constexpr (P4) {   // 32 rows of each of the four reads (stores of reads past B are clipped)
          for (int q4 = 0; q4 < 4; ++q4)
            ptx::tma_store_3d(&tmOut, z + (q4 >> 1) * 16384 + (q4 & 1) * 8192, 0, 0, (4 * pr + q4) * p.D + ch);
        } else {
          ptx::tma_store_3d(&tmOut, z, 0, 64 * c, 2 * pr * p.D + ch);
          if (2 * pr + 1 < p.B) ptx::tma_store_3d(&tmOut, z + 16384, 0, 64 * c, (2 * pr + 1) * p.D + ch);
        }
        ptx::tma_store_commit();
        if (j + NZ < n) {
          ptx::tma_store_wait_read<0>();
          load_z(j + NZ);
        }
      }
      ptx::tma_store_wait<0>();
    }
  } else if (warp == W_MMA) {
    // =========================== MMA issuer (whole warp, uniform control flow; one elected lane issues) ===========
    // Steps 3 and 5 (data from TMEM x constants) share one rolled loop, steps 1 and 7 (constants x data / data x constants
    // from shared memory) another: ~300 instructions, so that this warp's code stays resident in the instruction cache.
    ptx::setmaxnreg_dec<104>();
    constexpr uint32_t id1 = idesc2(128, false, true, false, false), id1n = idesc2(128, false, true, true, false);
    constexpr uint32_t id35 = idesc2(128, false, false, false, false), id35n = idesc2(128, false, false, false, true);
    constexpr uint32_t id7 = idesc2(P4 ? 32 : 64, true, false, false, false), id7n = idesc2(P4 ? 32 : 64, true, false, false, true);
    const uint32_t sS = ptx::smem_u32(smem + OFF2_S), sBT = ptx::smem_u32(smem + OFF2_BT);
    const uint64_t dS = ptx::smem_desc_k_sw128(sS);
    auto s_desc = [&](int row0, int kk) -> uint64_t {    // constant rows row0.., K index kk (multiple of 16)
      return dS + (uint64_t)(((kk >> 6) * S2_PANEL + row0 * 128 + (kk & 63) * 2) >> 4);
    };
    const uint64_t dBT = ptx::smem_desc_mn_sw128(sBT, BT_ATOM, 1024);
    const int n_slots = 8 * (n / 2 + 1);
#pragma unroll 1
    for (int s = 0; s < n_slots; ++s) {
      const int j = s & 7, k2x = (s >> 3) * 2;
      // slot order: M1 A, M7 B-, M3 A, M1 B, M5 A, M3 B, M7 A, M5 B   (item offset + 1 in 2 bits, step in 4 bits per slot)
      const int it = k2x + int((0x9991u >> (2 * j)) & 3u) - 1;
      if (it < 0 || it >= n) continue;
      const int step = int((0x57351371u >> (4 * j)) & 7u);
      const uint32_t type = it & 1, k = it >> 1, ph = k & 1;
      if constexpr (TR) stamp(0);
      if (step == 3 || step == 5) {
        // step 3: S_re = P_re Fre - P_im Fim -> U1,  S_im = P_re Fim + P_im Fre -> U3;  P1 = U0
        // step 5: B_im = -P_re Fim + P_im Fre -> U2,  B_re = P_re Fre + P_im Fim -> U3;  P2 = U1
        const uint32_t is5 = step == 5;
        uint64_t* pfull = p1_full + 4 * is5 + 2 * type;   // p2_full = p1_full + 4
        ptx::mbar_wait(&pfull[0], ph);
        if (is5 && type == 0) {
          // U2: the B-type partner's A_im (its E1 pulls it into registers first), or - without a partner - the previous
          // B-type item's z'
          if (it + 1 < n) ptx::mbar_wait(imd, ph);
          else if (k >= 1) ptx::mbar_wait(&e4_done[1], (k - 1) & 1);
        }
        ptx::tc_fence_after_sync();
        if constexpr (TR) stamp(0);
        const uint32_t src = is5 ? U1 : U0;
#pragma unroll 1
        for (uint32_t o = 0; o < 2; ++o) {
          if (o == 1 && !is5 && type == 1) {   // U3 still holds the A-type partner's B_re until its E3 has pulled it into registers
            ptx::mbar_wait(bdr, ph);
            ptx::tc_fence_after_sync();
          }
          const uint32_t dst = o == 0 ? (is5 ? U2 : U1) : U3;
          // constant block of the re (t = 0) / im (t = 1) K slices and the one negated product (first half only)
          const int rows0 = (o != is5) ? FIM : FRE, rows1 = (o != is5) ? FRE : FIM;
          const uint32_t idt0 = (o == 0 && is5) ? id35n : id35, idt1 = (o == 0 && !is5) ? id35n : id35;
          uint64_t dc0 = s_desc(rows0, 0), dc1 = s_desc(rows1, 0);
          uint32_t sa = src;
#pragma unroll 1
          for (int hq = 0; hq < 2; ++hq) {               // K panel = index half; packed K order: per run of 16 indices, re then im
            if (o == 0 && hq == 1) {   // second index half (step 5: also, E2 has finished reading S_im, the second half's target)
              ptx::mbar_wait(&pfull[1], ph);
              ptx::tc_fence_after_sync();
            }
#pragma unroll
            for (int q2 = 0; q2 < 4; ++q2) {
              ptx::umma_f16_ts_e(dst, sa + 16 * q2, dc0 + 2 * q2, idt0, (hq | q2) != 0);
              ptx::umma_f16_ts_e(dst, sa + 16 * q2 + 8, dc1 + 2 * q2, idt1, 1u);
            }
            dc0 += S2_PANEL >> 4; dc1 += S2_PANEL >> 4; sa += 64;
          }
        }
        ptx::umma_commit_e(y_full + 2 * is5 + type);   // x2_full = y_full + 2
      } else {
        // step 1: A_re = Fre Zre - Fim Zim -> U0,  A_im = Fim Zre + Fre Zim -> U1 (A-type) / U2 (B-type)   (constants x z)
        // step 7: z_re = Bre Fre + Bim Fim, z_im = Bim Fre - Bre Fim (n1 < 64) -> U0 (A-type) / U2 (B-type)   (BT x constants)
        const uint32_t is7 = step == 7;
        uint64_t dD;
        uint32_t d0, d1;
        if (!is7) {
          const int buf = it % NZ;
          if (k >= 1) ptx::mbar_wait(&e4_done[type], (k - 1) & 1);   // U0 / U2 held the previous same-type item's z'
          ptx::mbar_wait(&z_full[buf], (it / NZ) & 1);
          dD = ptx::smem_desc_mn_sw128(ptx::smem_u32(smem + OFF2_Z + buf * Z_BYTES), 8192, 1024);
          d0 = U0;
          d1 = type ? U2 : U1;
        } else {
          ptx::mbar_wait(&bt_full[type], ph);
          dD = dBT;
          d0 = type ? U2 : U0;
          d1 = d0 + 64;
        }
        ptx::tc_fence_after_sync();
        if constexpr (TR) stamp(0);
#pragma unroll 1
        for (uint32_t op = 0; op < 4; ++op) {             // (output half o, input part: z re / im rows, or BT's B_re / B_im rows)
          const uint32_t o = op >> 1, part = op & 1;
          const int rows = (o ^ part) ? FIM : FRE;
          const bool neg = (o == is7) && (part != is7);   // step 1: -Fim Zim in A_re; step 7: -Bre Fim in z_im
          const uint32_t dst = o ? d1 : d0;
          const uint64_t dc = s_desc(rows, 0), dd = dD + (uint64_t)((part * 16384) >> 4);
          if (!is7) {   // 64 K rows: one pass
            const uint32_t id = neg ? id1n : id1;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj)   // P4: input rows n1 = 0..31 and 64..95 -> K slices 0, 16 of panel 0 and of panel 1
              ptx::umma_f16_e(dst, P4 ? dc + (jj >> 1) * (S2_PANEL >> 4) + 2 * (jj & 1) : dc + 2 * jj, dd + 128 * jj, id, (part | jj) != 0);
          } else {      // 128 K rows: both panels of the constants
            const uint32_t id = neg ? id7n : id7;
            if constexpr (P4) {   // output rows n1 = 0..31 -> columns [0,32) of the half, n1 = 64..95 -> columns [32,64)
#pragma unroll 1
              for (int g = 0; g < 2; ++g) {
                const uint64_t dcg = dc + (uint64_t)((g * 64 * 128) >> 4);
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                  ptx::umma_f16_e(dst + 32 * g, dd + 128 * jj, dcg + (jj >> 2) * (S2_PANEL >> 4) + 2 * (jj & 3), id, (part | jj) != 0);
              }
            } else {
#pragma unroll
              for (int jj = 0; jj < 8; ++jj)
                ptx::umma_f16_e(dst, dd + 128 * jj, dc + (jj >> 2) * (S2_PANEL >> 4) + 2 * (jj & 3), id, (part | jj) != 0);
            }
          }
        }
        ptx::umma_commit_e(is7 ? &o_full[type] : &x_full[type]);
      }
      if constexpr (TR) stamp(0);
    }
  } else if (warp == W_AUX || warp == W_TAIL) {
    // =========================== gate-tile loads (warp W_AUX) / tail tokens (warp W_TAIL) ===========================
    // (two warps on two schedulers: whatever a helper warp executes is taken from the two epilogue warps it shares a
    // scheduler with, and the slowest epilogue warp sets the pace of every barrier)
    ptx::setmaxnreg_dec<104>();
    const bool gate_warp = warp == W_AUX;
    float a_prev = 0.f, b_prev = 0.f;   // tail recurrence over the chunks of an item (chunked form)
    for (int k = (gate_warp || (!P4 && nt > 0)) ? 0 : n; 2 * k <= n; ++k) {
#pragma unroll 1
      for (int sub = 0; sub < 4; ++sub) {
        // order of the events: x_full(A_k) [gate], bt_full(B_k-1) [tail], x_full(B_k) [gate], bt_full(A_k) [tail]
        const int it = 2 * k + (sub == 1 ? -1 : (sub == 2 ? 1 : 0));
        if (it < 0 || it >= n || ((sub & 1) == 0) != gate_warp) continue;
        const uint32_t type = it & 1, ph = (it >> 1) & 1;
        const int item = item0 + u_item(it), cc = u_chunk(it);
        const int ch = item / p.n_pairs, pr = item % p.n_pairs;
        const int b0 = 2 * pr, b1 = 2 * pr + 1;
        const bool has1 = b1 < p.B;
        if (sub == 0 || sub == 2) {
          if (lane == 0) {   // z has been consumed: its buffer now receives the x0 gate tile [n1][n2] of both reads
            const int buf = it % NZ;
            uint8_t* zb = smem + OFF2_Z + buf * Z_BYTES;
            ptx::mbar_wait(&x_full[type], ph);
            if constexpr (P4) {   // 32 rows of each of the four reads (reads past B: zero filled)
              ptx::mbar_expect_tx(&g_full[buf], Z_BYTES);
              for (int q4 = 0; q4 < 4; ++q4)
                ptx::tma_load_3d(zb + (q4 >> 1) * 16384 + (q4 & 1) * 8192, &tmX0, &g_full[buf], 0, 0, (4 * pr + q4) * p.D + ch);
            } else {
              ptx::mbar_expect_tx(&g_full[buf], has1 ? Z_BYTES : Z_BYTES / 2);
              ptx::tma_load_3d(zb, &tmX0, &g_full[buf], 0, 64 * cc, b0 * p.D + ch);
              if (has1) ptx::tma_load_3d(zb + 16384, &tmX0, &g_full[buf], 0, 64 * cc, b1 * p.D + ch);
            }
          }
          __syncwarp();
        } else if (nt > 0) {
          // Tail tokens t = NC C + j (j < nt <= 8).  Row n1 = 64 of a unit's inverse transform - an alternating sum over BT,
          // conj(F)[k1][64] = (-1)^k1 - is the value at C + j of that transform.  Single transform (NC = 1): that is the
          // taps 0..8191 part of output C + j, and the <= 2 (j + 1) products the transform wraps around are added directly:
          // x[a C + i] k'[(NC - a) C + j - i], a = 0..NC, i <= j.  Chunked: with the V-form tables unit m's transform is
          // w_m + shift(w_{m-1}), w_m = IFFT(sum_{c+s=m} S_c G_s), so its first-row outputs are F_m = a_m + b_{m-1} and its
          // row 64 is R_m = b_m + a_{m-1} (a / b = value j of the first / second half of w_m); the recurrence
          // a_m = F_m - b_{m-1}, b_m = R_m - a_{m-1} over the item's units yields b_{NC-1}, the pairs c + s = NC - 1, and
          // the pairs c + s = NC are the same direct products.  F_m comes from the E4 thread that owns it (tailF).
          // Lane = (j = lane & 7, a parity = bit 3, read = bit 4) for the products, (j, k1 mod 4 = lane >> 3) for the BT sums.
          const int jl = lane & 7, abit = (lane >> 3) & 1, rd = lane >> 4;
          const bool live = jl < nt && (rd == 0 || has1);
          const bool last = cc == NC - 1;
          const long long rowb = ((long long)(rd ? b1 : b0) * p.D + ch) * p.Tp;
          const float inva = __ldg(p.inva + ch), osc = __ldg(p.osc + ch);
          float tcv = 0.f, tx = 0.f;
          if (last) {
            const float* kq = p.k + (long long)ch * p.Lk;
#pragma unroll 1
            for (int a = abit; a <= NC; a += 2)
#pragma unroll 1
              for (int i = 0; i < nt; ++i)   // rolled (code size); nt = 1 for a maximum-length read
                if (live && i <= jl) {
                  const int tap = (NC - a) * C + jl - i;
                  float kv = __ldg(kq + tap);
                  if (tap == 0) kv += __ldg(p.dbias + ch);   // tap 0 carries the bias skip
                  tcv = fmaf(__half2float(p.vx[rowb + (long long)a * C + i]), kv, tcv);
                }
            if (live) tx = __bfloat162float(p.x0[rowb + (long long)NC * C + jl]);
          }
          tcv += __shfl_xor_sync(0xffffffffu, tcv, 8);
          ptx::mbar_wait(&bt_full[type], ph);
          // n2 = 0..7 is 16-byte chunk 0 of atom 0, stored at chunk position k1 & 7 of its 128-byte row
          float sre = 0.f, sim = 0.f;
          {
            // lane group g = lane >> 3 takes rows k1 = 4 i + g: the four groups of one load hit four different chunk
            // positions (k1 & 7), i.e. different banks; the sign (-1)^k1 = (-1)^g is a per-lane constant
            const int g4 = lane >> 3;
            const uint8_t* base = smem + OFF2_BT + jl * 2;
#pragma unroll 2
            for (int i = 0; i < 32; ++i) {
              const int rr = 4 * i + g4;
              const uint8_t* rowp = base + rr * 128 + ((rr & 7) << 4);
              sre += __half2float(*reinterpret_cast<const __half*>(rowp));
              sim += __half2float(*reinterpret_cast<const __half*>(rowp + 128 * 128));
            }
            if (g4 & 1) { sre = -sre; sim = -sim; }
          }
          sre += __shfl_xor_sync(0xffffffffu, sre, 8);
          sim += __shfl_xor_sync(0xffffffffu, sim, 8);
          sre += __shfl_xor_sync(0xffffffffu, sre, 16);
          sim += __shfl_xor_sync(0xffffffffu, sim, 16);
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&bt_read[type]);
          float tail = (rd ? sim : sre) * osc;
          if constexpr (CH) {
            const int buf = it % NZ;
            ptx::mbar_wait(&out_ready[buf], (it / NZ) & 1);   // E4 of this unit has published F_m
            const float F = tailF[(buf * 2 + rd) * 8 + jl];
            if (cc == 0) { a_prev = 0.f; b_prev = 0.f; }
            const float a_m = F - b_prev, b_m = tail - a_prev;
            a_prev = a_m;
            b_prev = b_m;
            tail = b_m;
          }
          if (last && live && abit == 0) p.out[rowb + (long long)NC * C + jl] = __float2bfloat16((tail + tcv * inva) * tx);
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    ptx::setmaxnreg_inc<200>();
    const int q = warp & 3, hf = (warp - EPI2_W0) >> 2;
    const int r = q * 32 + lane;                         // TMEM lane: k1 (E1-E3) or n2 (E4)
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t sBT = ptx::smem_u32(smem + OFF2_BT);
    // twiddles of this thread's row: step w = exp(-2 pi i r / N) and one seed per run of 16 indices
    float2 wstep, seed[4];
    sincospif(-2.0f * float(r) / float(N), &wstep.y, &wstep.x);
#pragma unroll
    for (int u = 0; u < 4; ++u) sincospif(-2.0f * float((r * (64 * hf + 16 * u)) % N) / float(N), &seed[u].y, &seed[u].x);
    // E1 / E2 walk the index (n2, k2) so that EVERY warp finishes the first half [0, 64) before the second
    auto col12 = [&](int u) { return (u < 2 ? 0 : 64) + 32 * hf + 16 * (u & 1); };
    float2 seed1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) sincospif(-2.0f * float((r * col12(u)) % N) / float(N), &seed1[u].y, &seed1[u].x);
    const bool tr = TR && trace && lane == 0;   // every epilogue warp records its own row (1 + warp)
    const float2 w2 = make_float2(wstep.x * wstep.x - wstep.y * wstep.y, 2.0f * wstep.x * wstep.y);
    const int n_slots = 8 * (n / 2 + 1);
    // channel of local item `it` without a division per slot (a CTA's items span at most a few channels)
    const int ch0 = item0 / p.n_pairs, pr0 = item0 % p.n_pairs;
    // (items per CTA <= 256 n_pairs / 148 + 1, so pr0 + it < 3 n_pairs + 1)
    auto ch_of = [&](int it) {
      const int x = pr0 + it, np = p.n_pairs;
      return ch0 + (x >= np) + (x >= 2 * np) + (x >= 3 * np);
    };
    // spectrum-table lines of this thread: uint4 (4 consecutive k2) at (((ch * 4 + k1 / 32) * 2 + k2 / 64) * 16 + (k2 % 64) / 4) * 32
    // + k1 % 32; an index half of this warp (k2 = 64 h + 32 hf + 0..31) is 8 uint4, 32 apart.  The first half of the NEXT E2
    // is fetched during the phase before it (E3 or E4: the L2 round trip would otherwise open every E2).
    // (`it` = work unit; chunked form: table of filter segment `seg`, the unit's own chunk always takes segment 0)
    auto g_ptr = [&](int it, int h, int seg = 0) {
      return p.G + (CH ? (size_t)seg * p.g_seg_stride : 0) + ((((size_t)ch_of(u_item(it)) * 4 + q) * 2) * 16 + 16 * h + 8 * hf) * 32 + lane;
    };
    // chunked form: this CTA's parked spectra, chunk c at uint4 [c * 4096, (c + 1) * 4096): run (h, uu), uint4 v of thread
    // tid_e at ((h * 2 + uu) * 4 + v) * 256 + tid_e (fp16 pairs in the spectrum-table layout)
    uint4* park = reinterpret_cast<uint4*>(p.scratch + (long long)blockIdx.x * p.scratch_per_cta);
    const int tid_e = (warp - EPI2_W0) * 32 + lane;
    uint4 g[8];        // table lines of the index half E2 works on next (lives across the phases: fetched one phase ahead)
    int gpre_it = -1;
#pragma unroll
    for (int v = 0; v < 8; ++v) g[v] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
    for (int s = 0; s < n_slots; ++s) {
      const int j = s & 7, k2x = (s >> 3) * 2;
      // slot order: E3 B-, E1 A, E4 B-, E2 A, E1 B, E3 A, E2 B, E4 A   (item offset + 1 in 2 bits, phase in 4 bits per slot)
      const int it = k2x + int((0x6644u >> (2 * j)) & 3u) - 1;
      if (it < 0 || it >= n) continue;
      const int phase = int((0x42312413u >> (4 * j)) & 7u);
      const uint32_t type = it & 1, ph = (it >> 1) & 1;
      if constexpr (TR) { if (tr) stamp(1 + warp - EPI2_W0); }
      // Twiddle seeds are re-materialised per phase: without the barrier ptxas precomputes all 256 twiddle values of the
      // thread once and keeps them in LOCAL memory (an L2 round trip per use with this shared-memory carve-out).
      float2 ws = wstep;
      float w2x = w2.x, w2y = w2.y;
      asm volatile("" : "+f"(ws.x), "+f"(ws.y), "+f"(w2x), "+f"(w2y));
      const f2t W2X = f2_pack(w2x, w2x), W2Y = f2_pack(w2y, w2y), NW2Y = f2_pack(-w2y, -w2y);
      // The phases walk their 64 indices as two rolled halves of two unrolled runs: the whole steady-state loop (all
      // warps) has to fit the 32 KB L1.5 instruction cache - fully unrolled it was 87 KB and every change of phase cost
      // hundreds of cycles of instruction fetch from L2.  Values pulled out of TMEM up front are consumed from the low
      // half of their register array, the high half moves down after the first pass.
      if (phase == 1) {
        // ------------------------------------------------ E1: P1 = fp16(S1 tw .* A), packed in place over A_re (U0)
        const uint32_t t_im = (type ? U2 : U1) + lane_addr, t_re = U0 + lane_addr;
        ptx::mbar_wait(&x_full[type], ph);
        ptx::tc_fence_after_sync();
        if constexpr (TR) { if (tr) stamp(1 + warp - EPI2_W0); }
        uint32_t xi[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u) tmem_ld16(t_im + col12(u), xi[u]);
        ptx::tmem_ld_wait();
        if (type == 1) {   // the A-type partner's step 5 writes its first half into this unit
          ptx::tc_fence_before_sync();
          __syncwarp();
          ptx::mbar_arrive_lane0(imd, lane);
        }
        float2 sc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          sc[u] = seed1[u];
          asm volatile("" : "+f"(sc[u].x), "+f"(sc[u].y));
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t xr2[2][16];
          tmem_ld16(t_re + 64 * h + 32 * hf, xr2[0]);
          tmem_ld16(t_re + 64 * h + 32 * hf + 16, xr2[1]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            uint32_t w[16];
            const uint32_t (&xr)[16] = xr2[uu];
            const uint32_t col = 64 * h + 32 * hf + 16 * uu;
            const float2 t0 = make_float2(sc[uu].x * S1, sc[uu].y * S1);
            const float2 t1 = make_float2(t0.x * ws.x - t0.y * ws.y, t0.x * ws.y + t0.y * ws.x);
            f2t TWX = f2_pack(t0.x, t1.x), TWY = f2_pack(t0.y, t1.y);   // twiddles of elements (2 j, 2 j + 1)
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const f2t XR = f2_packu(xr[2 * e], xr[2 * e + 1]), XI = f2_packu(xi[uu][2 * e], xi[uu][2 * e + 1]);
              w[e] = f2_to_h2(f2_sub(f2_mul(XR, TWX), f2_mul(XI, TWY)));
              w[8 + e] = f2_to_h2(f2_fma(XR, TWY, f2_mul(XI, TWX)));
              if (e < 7) {   // advance both twiddles by w^2
                const f2t NX = f2_fma(TWY, NW2Y, f2_mul(TWX, W2X));
                TWY = f2_fma(TWY, W2X, f2_mul(TWX, W2Y));
                TWX = NX;
              }
            }
            tmem_st8(t_re + col, w);          // K slice 2 (col / 16) (re)
            tmem_st8(t_re + col + 8, w + 8);  // K slice 2 (col / 16) + 1 (im)
          }
          ptx::tmem_st_wait();   // an index half is complete
          ptx::tc_fence_before_sync();
          __syncwarp();
          ptx::mbar_arrive_lane0(&p1_full[2 * type + h], lane);
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            sc[uu] = sc[2 + uu];
#pragma unroll
            for (int e = 0; e < 16; ++e) xi[uu][e] = xi[2 + uu][e];
          }
        }
      } else if (phase == 2) {
        // ------------------------------------------------ E2: P2 = fp16(S .* G'), packed in place over S_re (U1)
        const uint32_t t_re = U1 + lane_addr, t_im = U3 + lane_addr;
        uint4 g2[8];
        {
          if constexpr (!CH) {   // (chunked form: the second half's table lines are fetched between the halves - registers)
            const uint4* g1 = g_ptr(it, 1);
#pragma unroll
            for (int v = 0; v < 8; ++v) g2[v] = __ldg(g1 + v * 32);
          }
          if (gpre_it != it) {   // first item of the CTA: nothing ran before this E2
            const uint4* g0 = g_ptr(it, 0);
#pragma unroll
            for (int v = 0; v < 8; ++v) g[v] = __ldg(g0 + v * 32);
          }
        }
        ptx::mbar_wait(&y_full[type], ph);
        ptx::tc_fence_after_sync();
        if constexpr (TR) { if (tr) stamp(1 + warp - EPI2_W0); }
        const int cc = u_chunk(it);
        // table scale factors of the later filter segments (chunked form), fetched once per phase: a load inside the segment
        // loop put an L2 round trip in front of every product
        float rel1 = 1.f, rel2 = 1.f, rel3 = 1.f;
        if constexpr (CH) {
          const float* relp = p.rel + ch_of(u_item(it));
          if (cc >= 1) rel1 = __ldg(relp + p.D);
          if (cc >= 2) rel2 = __ldg(relp + 2 * p.D);
          if (cc >= 3) rel3 = __ldg(relp + 3 * p.D);
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t xr2[2][16], xi2[2][16];
          if constexpr (!CH) {
#pragma unroll
            for (int uu = 0; uu < 2; ++uu) {
              tmem_ld16(t_re + 64 * h + 32 * hf + 16 * uu, xr2[uu]);
              tmem_ld16(t_im + 64 * h + 32 * hf + 16 * uu, xi2[uu]);
            }
            ptx::tmem_ld_wait();
          }
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            uint32_t w[16];
            uint4 gq[4], sq[4];   // chunked form: operands of the first extra product, requested before anything else of the run
            const size_t slot = (size_t)((h * 2 + uu) * 4) * 256 + tid_e;
            if constexpr (CH) {   // one run at a time: the segment sum below needs the registers
              if (cc > 0) {
                const uint4* gj = g_ptr(it, h, 1) + (4 * uu) * 32;
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                  gq[v] = __ldg(gj + v * 32);
                  sq[v] = park[(size_t)(cc - 1) * 4096 + slot + (size_t)v * 256];
                }
              }
              tmem_ld16(t_re + 64 * h + 32 * hf + 16 * uu, xr2[uu]);
              tmem_ld16(t_im + 64 * h + 32 * hf + 16 * uu, xi2[uu]);
              ptx::tmem_ld_wait();
            }
            const uint32_t (&xr)[16] = xr2[uu];
            const uint32_t (&xi)[16] = xi2[uu];
            const uint32_t col = 64 * h + 32 * hf + 16 * uu;
            f2t ar[8], ai[8];   // products for elements (2 m, 2 m + 1)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              const uint32_t gw[4] = {g[4 * uu + v].x, g[4 * uu + v].y, g[4 * uu + v].z, g[4 * uu + v].w};
#pragma unroll
              for (int hp = 0; hp < 2; ++hp) {   // elements 4 v + 2 hp, + 1
                const int idx = 4 * v + 2 * hp;
                const f2t GR = h2_to_f2(gw[2 * hp]), GI = h2_to_f2(gw[2 * hp + 1]);
                const f2t XR = f2_packu(xr[idx], xr[idx + 1]), XI = f2_packu(xi[idx], xi[idx + 1]);
                ar[idx / 2] = f2_sub(f2_mul(XR, GR), f2_mul(XI, GI));
                ai[idx / 2] = f2_fma(XR, GI, f2_mul(XI, GR));
              }
            }
            if constexpr (CH) {
              // park this chunk's spectrum for the later chunks (fp16: the sum below is rounded to fp16 anyway), then add the
              // earlier chunks' spectra times the later V-form tables: V_c = sum_j S_{c-j} H_j; the loads of product j + 1
              // fly while product j is accumulated
              if (cc < NC - 1) {
#pragma unroll
                for (int v = 0; v < 4; ++v)
                  park[(size_t)cc * 4096 + slot + (size_t)v * 256] =
                      make_uint4(pack_f16(__uint_as_float(xr[4 * v]), __uint_as_float(xr[4 * v + 1])),
                                 pack_f16(__uint_as_float(xi[4 * v]), __uint_as_float(xi[4 * v + 1])),
                                 pack_f16(__uint_as_float(xr[4 * v + 2]), __uint_as_float(xr[4 * v + 3])),
                                 pack_f16(__uint_as_float(xi[4 * v + 2]), __uint_as_float(xi[4 * v + 3])));
              }
              if (cc > 0) {
#pragma unroll 1
                for (int j = 1; j <= cc; ++j) {
                  if (j > 1) {   // (product 1 was requested at the top of the run)
                    const uint4* gj = g_ptr(it, h, j) + (4 * uu) * 32;
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                      gq[v] = __ldg(gj + v * 32);
                      sq[v] = park[(size_t)(cc - j) * 4096 + slot + (size_t)v * 256];
                    }
                  }
                  const float relj = j == 1 ? rel1 : (j == 2 ? rel2 : rel3);   // 2^(e_0 - e_j): table j -> table 0's scale
                  const f2t REL = f2_pack(relj, relj);
#pragma unroll
                  for (int v = 0; v < 4; ++v) {
                    const uint32_t gw[4] = {gq[v].x, gq[v].y, gq[v].z, gq[v].w};
                    const uint32_t sw[4] = {sq[v].x, sq[v].y, sq[v].z, sq[v].w};
#pragma unroll
                    for (int hp = 0; hp < 2; ++hp) {
                      const int m = 2 * v + hp;   // elements 2 m, 2 m + 1
                      const f2t GR = f2_mul(h2_to_f2(gw[2 * hp]), REL), GI = f2_mul(h2_to_f2(gw[2 * hp + 1]), REL);
                      const f2t XR = h2_to_f2(sw[2 * hp]), XI = h2_to_f2(sw[2 * hp + 1]);
                      ar[m] = f2_sub(f2_fma(XR, GR, ar[m]), f2_mul(XI, GI));
                      ai[m] = f2_fma(XI, GR, f2_fma(XR, GI, ai[m]));
                    }
                  }
                }
              }
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) {
              w[m] = f2_to_h2(ar[m]);
              w[8 + m] = f2_to_h2(ai[m]);
            }
            tmem_st8(t_re + col, w);
            tmem_st8(t_re + col + 8, w + 8);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before_sync();
          __syncwarp();
          ptx::mbar_arrive_lane0(&p2_full[2 * type + h], lane);
          if constexpr (CH) {
            if (h == 0) {
              const uint4* g1 = g_ptr(it, 1);
#pragma unroll
              for (int v = 0; v < 8; ++v) g[v] = __ldg(g1 + v * 32);
            }
          } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) g[v] = g2[v];
          }
        }
      } else if (phase == 3) {
        // ------------------------------------------------ E3: BT = fp16(conj(tw) .* B), shared memory
        const uint32_t t_bim = U2 + lane_addr + 64 * hf, t_bre = U3 + lane_addr + 64 * hf;
        if (type == 0 && it + 1 < n) {   // E3 of an A-type item: the next phase is E2 of its partner, item it + 1
          const uint4* g0 = g_ptr(it + 1, 0);
#pragma unroll
          for (int v = 0; v < 8; ++v) g[v] = __ldg(g0 + v * 32);
          gpre_it = it + 1;
        }
        ptx::mbar_wait(&x2_full[type], ph);
        ptx::tc_fence_after_sync();
        if constexpr (TR) { if (tr) stamp(1 + warp - EPI2_W0); }
        uint32_t xr[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u) tmem_ld16(t_bre + 16 * u, xr[u]);
        ptx::tmem_ld_wait();
        if (type == 0) {   // the B-type partner's step 3 writes S_im into this unit
          ptx::tc_fence_before_sync();
          __syncwarp();
          ptx::mbar_arrive_lane0(bdr, lane);
        }
        if (nt > 0 && it >= 1) ptx::mbar_wait(&bt_read[type ^ 1], ((it - 1) >> 1) & 1);   // the tail warp is done with BT
        float2 sc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          sc[u] = seed[u];
          asm volatile("" : "+f"(sc[u].x), "+f"(sc[u].y));
        }
        // row r (B_re) and row 128 + r (B_im) of atom hf; run u covers 16-byte chunks 2 u, 2 u + 1
        const uint32_t base_re = sBT + hf * BT_ATOM + r * 128, base_im = base_re + 128 * 128;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t xi2[2][16];
          tmem_ld16(t_bim + 32 * h, xi2[0]);
          tmem_ld16(t_bim + 32 * h + 16, xi2[1]);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            uint32_t wr[8], wi[8];
            const uint32_t (&xi)[16] = xi2[uu];
            const int u = 2 * h + uu;
            const float2 t0 = sc[uu];
            const float2 t1 = make_float2(t0.x * ws.x - t0.y * ws.y, t0.x * ws.y + t0.y * ws.x);
            f2t TWX = f2_pack(t0.x, t1.x), TWY = f2_pack(t0.y, t1.y);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const f2t BR = f2_packu(xr[uu][2 * e], xr[uu][2 * e + 1]), BI = f2_packu(xi[2 * e], xi[2 * e + 1]);
              wr[e] = f2_to_h2(f2_fma(BI, TWY, f2_mul(BR, TWX)));      // (br + i bi)(tw.x - i tw.y)
              wi[e] = f2_to_h2(f2_sub(f2_mul(BI, TWX), f2_mul(BR, TWY)));
              if (e < 7) {
                const f2t NX = f2_fma(TWY, NW2Y, f2_mul(TWX, W2X));
                TWY = f2_fma(TWY, W2X, f2_mul(TWX, W2Y));
                TWX = NX;
              }
            }
#pragma unroll
            for (int jj = 0; jj < 2; ++jj) {
              const uint32_t off = uint32_t((2 * u + jj) ^ (r & 7)) << 4;
              ptx::st_shared_v4(base_re + off, wr[4 * jj], wr[4 * jj + 1], wr[4 * jj + 2], wr[4 * jj + 3]);
              ptx::st_shared_v4(base_im + off, wi[4 * jj], wi[4 * jj + 1], wi[4 * jj + 2], wi[4 * jj + 3]);
            }
          }
#pragma unroll
          for (int uu = 0; uu < 2; ++uu) {
            sc[uu] = sc[2 + uu];
#pragma unroll
            for (int e = 0; e < 16; ++e) xr[uu][e] = xr[2 + uu][e];
          }
        }
        ptx::fence_proxy_async_smem();
        ptx::tc_fence_before_sync();
        __syncwarp();
        ptx::mbar_arrive_lane0(&bt_full[type], lane);
      } else {
        // ------------------------------------------------ E4: out = z' * x0, in place over the gate tile, TMA store
        const int buf = it % NZ;
        uint8_t* zb = smem + OFF2_Z + buf * Z_BYTES;
        // output scale of this channel (exact power of two; P4: times 2^(e - e4), the 4096-tap table has its own exponent)
        const float osc = P4 ? __ldg(p.osc + ch_of(it)) * __ldg(p.osc_adj + ch_of(it)) : __ldg(p.osc + ch_of(u_item(it)));
        const uint32_t t_z = (type ? U2 : U0) + lane_addr + 32 * hf;
        if (type == 1 && it + 1 < n) {   // E4 of a B-type item: the next phase is E2 of item it + 1
          const uint4* g0 = g_ptr(it + 1, 0);
#pragma unroll
          for (int v = 0; v < 8; ++v) g[v] = __ldg(g0 + v * 32);
          gpre_it = it + 1;
        }
        ptx::mbar_wait(&o_full[type], ph);
        ptx::tc_fence_after_sync();
        if constexpr (TR) { if (tr) stamp(1 + warp - EPI2_W0); }
        uint32_t zr[2][16], zi[2][16];
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          tmem_ld16(t_z + 16 * h2, zr[h2]);
          tmem_ld16(t_z + 64 + 16 * h2, zi[h2]);
        }
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        ptx::mbar_arrive_lane0(&e4_done[type], lane);   // the unit is free for the next same-type item's step 1
        ptx::mbar_wait(&g_full[buf], (it / NZ) & 1);
        // [n1][n2] bf16, 256 B per n1 row: gate in, product out
        unsigned short* st0 = reinterpret_cast<unsigned short*>(zb) + r + 32 * hf * 128;
        // One output per thread is enough for the range check: an overflow in P1 / P2 reaches every output of the item, one
        // in BT reaches every output of its row n2 - and a row is a thread here.
        const float chk = fmaf(__uint_as_float(zr[0][0]), 0.f, __uint_as_float(zi[0][0]) * 0.f);
        if constexpr (CH) {   // first-row outputs (n1 = 0, n2 = r < 8) before the gate: F_m of the tail recurrence
          if (nt > 0 && hf == 0 && r < 8) {
            tailF[(buf * 2 + 0) * 8 + r] = __uint_as_float(zr[0][0]) * osc;
            tailF[(buf * 2 + 1) * 8 + r] = __uint_as_float(zi[0][0]) * osc;
          }
        }
        // two n1 rows per packed multiply; bf16 -> fp32 of the gate by byte permute (ALU pipe: the FMA pipe is the busy one)
        const f2t OSC = f2_pack(osc, osc);
        auto gate2 = [&](const unsigned short* g) {
          return f2_packu(__byte_perm(uint32_t(g[0]), 0u, 0x1044), __byte_perm(uint32_t(g[128]), 0u, 0x1044));
        };
        auto put2 = [&](unsigned short* g, f2t v) {   // one rounding, fp32 -> bf16, as before
          float lo, hi;
          f2_unpack(v, lo, hi);
          const __nv_bfloat162 pk = __floats2bfloat162_rn(lo, hi);
          const uint32_t u = *reinterpret_cast<const uint32_t*>(&pk);
          g[0] = (unsigned short)u;
          g[128] = (unsigned short)(u >> 16);
        };
#pragma unroll 1
        for (int h2 = 0; h2 < 2; ++h2) {
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            // (an odd batch's last item has no second read: that half of the buffer is neither loaded nor stored)
            unsigned short* ga = st0 + e * 128;
            unsigned short* gb = ga + 8192;
            const f2t GA = gate2(ga), GB = gate2(gb);
            put2(ga, f2_mul(f2_mul(f2_packu(zr[0][e], zr[0][e + 1]), OSC), GA));
            put2(gb, f2_mul(f2_mul(f2_packu(zi[0][e], zi[0][e + 1]), OSC), GB));
          }
          st0 += 16 * 128;
#pragma unroll
          for (int e = 0; e < 16; ++e) { zr[0][e] = zr[1][e]; zi[0][e] = zi[1][e]; }
        }
        if (chk != chk) atomicOr(p.err, 2);   // an fp16 operand overflowed somewhere in this item: the launch is reported
        ptx::fence_proxy_async_smem();
        __syncwarp();
        ptx::mbar_arrive_lane0(&out_ready[buf], lane);
      }
      if constexpr (TR) { if (tr) stamp(1 + warp - EPI2_W0); }
    }
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == W_MMA) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace clm
