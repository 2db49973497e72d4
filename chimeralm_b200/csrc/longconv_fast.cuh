// Single-chunk long convolution, tuned version (T <= N/2 + LONGCONV_TAIL_MAX).
//
// Same math as longconv_kernel (longconv.cuh) - y = causal_conv(vx, k) + bias*vx, out = y*x0, two reads
// per complex FFT - restructured around what the profile of the first version showed
// (profiles/r1_v4_longconv.txt: 33% long-scoreboard stalls on global loads between FFT passes, 11%
// barrier stalls, issue slots dominated by twiddle generation and scalar complex adds):
//   * the first forward pass reads its inputs straight from global memory (upper half is the
//     zero padding, so half of each radix-R0 butterfly's inputs are known zeros),
//   * the last forward radix-16 pass, the product with the filter spectrum and the first inverse
//     radix-16 pass touch the same 16 contiguous points, so they are one pass with the
//     spectrum loaded coalesced from a [16][N/16] transposed table,
//   * the last inverse pass computes only the outputs that survive overlap-save (upper half),
//     applies the x0 gate and writes bf16 to global from registers,
//   * radix-16 twiddles come from shared-memory tables instead of sincospif + a power chain,
//   * complex add/sub use packed fp32x2 instructions (FADD2 on sm_100),
//   * the bias skip is folded into filter tap 0 (k'[0] = k[0] + bias), so the output pass does
//     not re-read vx; the next item's vx rows are copied into a shared-memory staging buffer with
//     cp.async (and its x0 rows prefetched into L2) while this item is transformed.
// Barriers per item: 7 (was 11); shared-memory round trips: 6 (was 11).
#pragma once
#include <cuda_bf16.h>

#include "fft.cuh"
#include "longconv.cuh"

namespace clm {
namespace f2 {

typedef unsigned long long cx;  // packed (re, im) fp32 pair in an aligned register pair

__device__ __forceinline__ cx mk(float x, float y) {
  cx r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ void un(cx a, float& x, float& y) { asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a)); }
__device__ __forceinline__ cx cadd(cx a, cx b) { cx d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ cx csub(cx a, cx b) { cx d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ cx cmul(cx a, float2 w) {   // a * w
  float x, y;
  un(a, x, y);
  return mk(x * w.x - y * w.y, x * w.y + y * w.x);
}
__device__ __forceinline__ cx cmulc(cx a, float2 w) {  // a * conj(w)
  float x, y;
  un(a, x, y);
  return mk(x * w.x + y * w.y, y * w.x - x * w.y);
}
template <bool INV>
__device__ __forceinline__ cx mul_mi(cx a) {  // a * (-i) forward, a * (+i) inverse
  float x, y;
  un(a, x, y);
  return INV ? mk(-y, x) : mk(y, -x);
}
template <bool INV>
__device__ __forceinline__ void dft4(cx& a, cx& b, cx& c, cx& d) {
  const cx s0 = cadd(a, c), d0 = csub(a, c), s1 = cadd(b, d), d1 = mul_mi<INV>(csub(b, d));
  a = cadd(s0, s1);
  c = csub(s0, s1);
  b = cadd(d0, d1);
  d = csub(d0, d1);
}
template <bool INV, int E>
__device__ __forceinline__ cx mul_w16(cx a) {
  constexpr float C1 = 0.9238795325112867f, S1 = 0.3826834323650898f, H = 0.7071067811865476f;
  constexpr int e = E & 15;
  if constexpr (e == 0) return a;
  else if constexpr (e == 4) return mul_mi<INV>(a);
  else if constexpr (e == 8) { float x, y; un(a, x, y); return mk(-x, -y); }
  else if constexpr (e == 12) return mul_mi<!INV>(a);
  else {
    constexpr float cr = (e == 1 || e == 15) ? C1 : (e == 2 || e == 14) ? H : (e == 3 || e == 13) ? S1
                       : (e == 5 || e == 11) ? -S1 : (e == 6 || e == 10) ? -H : -C1;
    constexpr float sn = (e == 1 || e == 7) ? S1 : (e == 2 || e == 6) ? H : (e == 3 || e == 5) ? C1
                       : (e == 9 || e == 15) ? -S1 : (e == 10 || e == 14) ? -H : -C1;
    constexpr float wi = INV ? sn : -sn;
    float x, y;
    un(a, x, y);
    return mk(x * cr - y * wi, x * wi + y * cr);
  }
}
// natural-order 16-point DFT (4x4), same index algebra as fft::Dft<16>
template <bool INV>
__device__ __forceinline__ void dft16(cx (&x)[16]) {
  dft4<INV>(x[0], x[4], x[8], x[12]);
  dft4<INV>(x[1], x[5], x[9], x[13]);
  dft4<INV>(x[2], x[6], x[10], x[14]);
  dft4<INV>(x[3], x[7], x[11], x[15]);
  x[5] = mul_w16<INV, 1>(x[5]);   x[9] = mul_w16<INV, 2>(x[9]);    x[13] = mul_w16<INV, 3>(x[13]);
  x[6] = mul_w16<INV, 2>(x[6]);   x[10] = mul_w16<INV, 4>(x[10]);  x[14] = mul_w16<INV, 6>(x[14]);
  x[7] = mul_w16<INV, 3>(x[7]);   x[11] = mul_w16<INV, 6>(x[11]);  x[15] = mul_w16<INV, 9>(x[15]);
  dft4<INV>(x[0], x[1], x[2], x[3]);
  dft4<INV>(x[4], x[5], x[6], x[7]);
  dft4<INV>(x[8], x[9], x[10], x[11]);
  dft4<INV>(x[12], x[13], x[14], x[15]);
  cx t;
#define CLM_SWAP(a, b) t = x[a]; x[a] = x[b]; x[b] = t;
  CLM_SWAP(1, 4) CLM_SWAP(2, 8) CLM_SWAP(3, 12) CLM_SWAP(6, 9) CLM_SWAP(7, 13) CLM_SWAP(11, 14)
#undef CLM_SWAP
}

// Twiddle tables, laid out so that a warp (consecutive j) reads consecutive addresses:
//   tA[(q-1)*256 + j] = W_4096^(j q), j < 256, q = 1..15;   tB[(q-1)*16 + j] = W_256^(j q), j < 16.
constexpr int TA = 15 * 256, TB = 15 * 16;

template <int NB>
__device__ __forceinline__ float2 tw_lookup(const float2* tA, const float2* tB, int j, int q) {
  if constexpr (NB == 256) return tB[(q - 1) * 16 + j];
  else return tA[(q - 1) * 256 + j];
}

// in-place radix-16 pass over blocks of NB (4096 or 256), twiddles from the tables
template <int N, int NB, bool INV, int THREADS>
__device__ __forceinline__ void pass16(cx* z, const float2* tA, const float2* tB, int tid) {
  constexpr int S = NB / 16;
#pragma unroll 1
  for (int i = tid; i < N / 16; i += THREADS) {
    const int j = i % S;
    const int base = (i / S) * NB + j;
    cx x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = z[fft::pad_idx(base + r * S)];
    if constexpr (!INV) {
      dft16<false>(x);
#pragma unroll
      for (int q = 1; q < 16; ++q) x[q] = cmul(x[q], tw_lookup<NB>(tA, tB, j, q));
    } else {
#pragma unroll
      for (int q = 1; q < 16; ++q) x[q] = cmulc(x[q], tw_lookup<NB>(tA, tB, j, q));
      dft16<true>(x);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) z[fft::pad_idx(base + r * S)] = x[r];
  }
}

}  // namespace f2

template <int LOGN>
struct FastCfg {
  static constexpr int N = 1 << LOGN, C = N / 2;
  static constexpr int R0 = (LOGN % 4) ? (1 << (LOGN % 4)) : 16;   // leading radix 2/4/8, or 16 when logN is a multiple of 4
  static constexpr bool kSupported = (LOGN >= 8) && (LOGN <= 14);
  static constexpr int THREADS = ConvCfg<LOGN>::THREADS;
  static constexpr int OFF_TA = fft::padded_size(N) * 8;
  static constexpr int OFF_TB = OFF_TA + f2::TA * 8;
  // leading-pass twiddles W_N^j, j < N/R0 <= 4096, as W_N^(64 a) * W_N^b (two 64-entry tables, one complex multiply)
  static constexpr int OFF_TC = OFF_TB + f2::TB * 8;
  static constexpr int OFF_STG = (OFF_TC + 2 * 64 * 8 + 15) / 16 * 16;   // bf16 staging of the NEXT item's vx rows: [2][C]
  static constexpr int SMEM = OFF_STG + 2 * C * 2;
};

// Spectra in the layout the fused middle pass wants: gT[m][c][q][b] = G_{m,c}[16 b + q], b < N/16, for filter
// segment m (g_m[p] = k[mC + p - C], longconv.cuh); the bias skip is folded into segment 0.
template <int LOGN>
__global__ void __launch_bounds__(ConvCfg<LOGN>::THREADS) filter_spectrum_fast_kernel(const float* __restrict__ k,
                                                                                       long long Lk, int L,
                                                                                       const float* __restrict__ dbias,
                                                                                       float2* __restrict__ gT, int D) {
  using Cfg = ConvCfg<LOGN>;
  constexpr int N = Cfg::N, C = Cfg::C, TH = Cfg::THREADS;
  extern __shared__ float2 zs[];
  const int c = blockIdx.x, m = blockIdx.y, tid = threadIdx.x;   // m = filter segment (overlap-save chunk lag)
  const float* kc = k + (long long)c * Lk;
  for (int pidx = tid; pidx < N; pidx += TH) {
    const long long s = (long long)m * C + pidx - C;
    float v = (pidx >= 1 && s >= 0 && s < L) ? kc[s] : 0.f;
    if (s == 0) v += dbias[c];  // bias skip y += bias*vx  ==  k'[0] = k[0] + bias
    zs[fft::pad_idx(pidx)] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft::fft_forward<LOGN, TH>(zs, tid);
  float2* out = gT + ((long long)m * D + c) * N;
  const float sc = 1.0f / (float)N;
  for (int i = tid; i < N; i += TH) {
    const float2 v = zs[fft::pad_idx(i)];
    out[(i & 15) * (N / 16) + (i >> 4)] = make_float2(v.x * sc, v.y * sc);
  }
}

struct LongConvFastParams {
  const __nv_bfloat16* vx;
  const __nv_bfloat16* x0;
  __nv_bfloat16* out;
  const float2* gT;     // [n_seg][D][16][N/16]
  float2* scratch;      // [gridDim.x][n_chunks][N] chunk spectra (n_chunks > 1 only)
  int n_chunks;         // overlap-save chunks of C = N/2 tokens; outputs [0, n_chunks*C) come from the FFT path
  const float* k;       // [D][Lk]   (ragged end)
  const float* dbias;   // [D]       (ragged end)
  long long Lk;
  int B, D, T, Tp, n_items;
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int LOGN>
__global__ void __launch_bounds__(FastCfg<LOGN>::THREADS, 1) longconv_fast_kernel(LongConvFastParams p) {
  using F = FastCfg<LOGN>;
  using f2::cx;
  constexpr int N = F::N, C = F::C, TH = F::THREADS, R0 = F::R0, S0 = N / R0, HALF = R0 / 2;
  static_assert(F::kSupported, "longconv_fast_kernel supports 256 <= N <= 16384");
  extern __shared__ __align__(16) uint8_t smem_f[];
  cx* z = reinterpret_cast<cx*>(smem_f);
  float2* tA = reinterpret_cast<float2*>(smem_f + F::OFF_TA);
  float2* tB = reinterpret_cast<float2*>(smem_f + F::OFF_TB);
  __shared__ float red[2][TH / 32];
  const int tid = threadIdx.x;
  for (int e = tid; e < f2::TA; e += TH) {
    float sn, cs;
    sincospif((float)((e & 255) * ((e >> 8) + 1)) / 2048.0f, &sn, &cs);   // 2*pi*(j q)/4096
    tA[e] = make_float2(cs, -sn);
  }
  for (int e = tid; e < f2::TB; e += TH) {
    float sn, cs;
    sincospif((float)((e & 15) * ((e >> 4) + 1)) / 128.0f, &sn, &cs);     // 2*pi*(j q)/256
    tB[e] = make_float2(cs, -sn);
  }
  float2* tC = reinterpret_cast<float2*>(smem_f + F::OFF_TC);   // [0,64): W_N^(64 a);  [64,128): W_N^b
  for (int e = tid; e < 128; e += TH) {
    float sn, cs;
    const int k = (e < 64) ? 64 * e : (e - 64);
    sincospif(2.0f * (float)k / (float)N, &sn, &cs);
    tC[e] = make_float2(cs, -sn);
  }
  auto lead_twiddle = [&](int j) -> float2 {   // W_N^j for j < 4096
    return fft::cmul(tC[j >> 6], tC[64 + (j & 63)]);
  };
  __nv_bfloat16* stg = reinterpret_cast<__nv_bfloat16*>(smem_f + F::OFF_STG);
  // asynchronous copy of one item's vx rows (both reads, first min(C, Tp) tokens) into the staging buffer
  auto stage_item = [&](int it_, int ch_) {
    const int c_ = it_ / ((p.B + 1) / 2), b_ = (it_ % ((p.B + 1) / 2)) * 2;
    const long long o0 = ((long long)b_ * p.D + c_) * p.Tp + (long long)ch_ * C;
    const int ncopy = max(0, min(C, p.Tp - ch_ * C)) / 8;   // 16-byte pieces per read
    const int nseq = (b_ + 1 < p.B) ? 2 : 1;
    for (int i = tid; i < nseq * ncopy; i += TH) {
      const int sq = i / ncopy, pc = i % ncopy;
      const __nv_bfloat16* src = p.vx + o0 + (long long)sq * p.D * p.Tp + pc * 8;
      const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(stg + sq * C + pc * 8));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if ((int)blockIdx.x < p.n_items) stage_item(blockIdx.x, 0);
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const int n_pairs = (p.B + 1) / 2;
  const int t_fft = min(p.n_chunks * C, p.T);
  __syncthreads();

  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const int c = item / n_pairs;
    const int b0 = (item % n_pairs) * 2;
    const bool has_b1 = (b0 + 1) < p.B;
    const long long off0 = ((long long)b0 * p.D + c) * p.Tp;
    const long long off1 = has_b1 ? off0 + (long long)p.D * p.Tp : off0;
    const __nv_bfloat16* va = p.vx + off0;
    const __nv_bfloat16* vb = p.vx + off1;
    for (int ch = 0; ch < p.n_chunks; ++ch) {
    const int tbase = ch * C;
    // the unit (item, chunk) processed after this one
    const bool last_chunk = (ch + 1 == p.n_chunks);
    const int n_item = last_chunk ? item + (int)gridDim.x : item;
    const int n_ch = last_chunk ? 0 : ch + 1;
    // ---- L2 prefetch of the next unit's x0 rows
    {
      const int nitem = n_item;
      if (nitem < p.n_items) {
        const int nc = nitem / n_pairs, nb0 = (nitem % n_pairs) * 2;
        const long long no0 = ((long long)nb0 * p.D + nc) * p.Tp + (long long)n_ch * C;
        const long long no1 = (nb0 + 1 < p.B) ? no0 + (long long)p.D * p.Tp : no0;
        const int lines = (max(0, min(p.T - n_ch * C, C)) * 2 + 127) / 128;
        for (int l = tid; l < 2 * lines; l += TH) {     // x0 rows only: vx is staged through shared memory
          const int which = l / lines, ln = l % lines;
          prefetch_l2(p.x0 + (which ? no1 : no0) + ln * 64);
        }
      }
    }
    // ---- phase A: global load fused with the leading radix-R0 DIF pass (inputs r >= R0/2 are zero padding)
    {
      constexpr int PAIRS = S0 / 2;   // two adjacent butterflies per step
#pragma unroll 2
      for (int i = tid; i < PAIRS; i += TH) {
        const int j = 2 * i;
        float xa[HALF][2], xb[HALF][2];
#pragma unroll
        for (int r = 0; r < HALF; ++r) {
          const int t = j + r * S0;
          uint32_t wa = 0, wb = 0;
          if (tbase + t < p.T) {  // t even: the pair (t, t+1) was staged (tbase + t + 1 < Tp)
            wa = *reinterpret_cast<const uint32_t*>(stg + t);
            if (has_b1) wb = *reinterpret_cast<const uint32_t*>(stg + C + t);
          }
          xa[r][0] = __uint_as_float(wa << 16);
          xa[r][1] = (tbase + t + 1 < p.T) ? __uint_as_float(wa & 0xffff0000u) : 0.f;
          xb[r][0] = __uint_as_float(wb << 16);
          xb[r][1] = (tbase + t + 1 < p.T) ? __uint_as_float(wb & 0xffff0000u) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          cx x[R0];
#pragma unroll
          for (int r = 0; r < R0; ++r) x[r] = (r < HALF) ? f2::mk(xa[r < HALF ? r : 0][u], xb[r < HALF ? r : 0][u]) : f2::mk(0.f, 0.f);
          const float2 w1 = lead_twiddle(j + u);
          if constexpr (R0 == 2) {
            const cx a = x[0];   // x[1] == 0
            x[1] = f2::cmul(a, w1);
          } else if constexpr (R0 == 4) {
            // DFT4 with x2 = x3 = 0: X0 = a+b, X1 = a - i b, X2 = a - b, X3 = a + i b
            const cx a = x[0], b = x[1], ib = f2::mul_mi<false>(b);
            const float2 w2 = fft::cmul(w1, w1), w3 = fft::cmul(w2, w1);
            x[0] = f2::cadd(a, b);
            x[1] = f2::cmul(f2::cadd(a, ib), w1);
            x[2] = f2::cmul(f2::csub(a, b), w2);
            x[3] = f2::cmul(f2::csub(a, ib), w3);
          } else if constexpr (R0 == 16) {  // DFT16 on (x0..x7, 0 x 8); twiddles W_N^(j q) straight from the tables
            f2::dft16<false>(x);
#pragma unroll
            for (int q = 1; q < 16; ++q) x[q] = f2::cmul(x[q], f2::tw_lookup<N>(tA, tB, j + u, q));
          } else {  // R0 == 8: generic DFT8 on (x0..x3, 0, 0, 0, 0)
            float2 xf[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) { float a_, b_; f2::un(x[r], a_, b_); xf[r] = make_float2(a_, b_); }
            fft::Dft<8, false>::run(xf);
            float2 w = w1;
            x[0] = f2::mk(xf[0].x, xf[0].y);
#pragma unroll
            for (int q = 1; q < 8; ++q) {
              const float2 y = fft::cmul(xf[q], w);
              x[q] = f2::mk(y.x, y.y);
              w = fft::cmul(w, w1);
            }
          }
#pragma unroll
          for (int q = 0; q < R0; ++q) z[fft::pad_idx(j + u + q * S0)] = x[q];
        }
      }
    }
    __syncthreads();
    if (n_item < p.n_items) stage_item(n_item, n_ch);   // overlaps the whole transform
    // ---- forward radix-16 passes down to blocks of 256
    if constexpr (S0 >= 4096) {
      f2::pass16<N, 4096, false, TH>(z, tA, tB, tid);
      __syncthreads();
    }
    if constexpr (S0 >= 256) {
      f2::pass16<N, 256, false, TH>(z, tA, tB, tid);
      __syncthreads();
    }
    // ---- phase M: last forward radix-16 (blocks of 16) * spectrum * first inverse radix-16
    if (p.n_chunks == 1) {
      const float2* g = p.gT + (long long)c * N;
#pragma unroll 1
      for (int bfly = tid; bfly < N / 16; bfly += TH) {
        float2 gq[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) gq[q] = __ldg(g + q * (N / 16) + bfly);
        cx x[16];
        const int base = fft::pad_idx(bfly * 16);   // 16 contiguous padded slots
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = z[base + r];
        f2::dft16<false>(x);
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = f2::cmul(x[q], gq[q]);
        f2::dft16<true>(x);
#pragma unroll
        for (int r = 0; r < 16; ++r) z[base + r] = x[r];
      }
    } else {
      // overlap-save: Y_ch = sum_{j <= ch} U_j . G_{ch-j}; earlier chunks' spectra U_j are parked in this CTA's scratch
      // (L2-resident) in the same [q][butterfly] layout, so every load is coalesced
      const float2* g = p.gT + (long long)c * N;
      const long long gseg = (long long)p.D * N;
      cx* sc = reinterpret_cast<cx*>(p.scratch) + (long long)blockIdx.x * p.n_chunks * N;
#pragma unroll 1
      for (int bfly = tid; bfly < N / 16; bfly += TH) {
        cx x[16];
        const int base = fft::pad_idx(bfly * 16);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = z[base + r];
        f2::dft16<false>(x);
        if (!last_chunk) {
#pragma unroll
          for (int q = 0; q < 16; ++q) sc[((long long)ch * 16 + q) * (N / 16) + bfly] = x[q];
        }
        float2 acc[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          float xr, xi;
          f2::un(x[q], xr, xi);
          acc[q] = fft::cmul(make_float2(xr, xi), __ldg(g + q * (N / 16) + bfly));
        }
        for (int j = 0; j < ch; ++j) {
          const float2* gj = g + (long long)(ch - j) * gseg;
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            float ur, ui;
            f2::un(sc[((long long)j * 16 + q) * (N / 16) + bfly], ur, ui);
            const float2 gv = __ldg(gj + q * (N / 16) + bfly);
            acc[q].x += ur * gv.x - ui * gv.y;
            acc[q].y += ur * gv.y + ui * gv.x;
          }
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = f2::mk(acc[q].x, acc[q].y);
        f2::dft16<true>(x);
#pragma unroll
        for (int r = 0; r < 16; ++r) z[base + r] = x[r];
      }
    }
    __syncthreads();
    // x0 for the output pass is fetched now (L2-prefetched rows) so its latency hides behind the inverse passes
    constexpr int ZITER = (S0 / 2 + TH - 1) / TH;
    uint32_t gxa[ZITER][HALF], gxb[ZITER][HALF];
#pragma unroll
    for (int zi = 0; zi < ZITER; ++zi) {
      const int j = 2 * (tid + zi * TH);
#pragma unroll
      for (int r = 0; r < HALF; ++r) {
        const int t = j + r * S0;
        gxa[zi][r] = gxb[zi][r] = 0;
        if (j < S0 && tbase + t < t_fft) {
          gxa[zi][r] = __ldg(reinterpret_cast<const unsigned int*>(p.x0 + off0 + tbase + t));
          if (has_b1) gxb[zi][r] = __ldg(reinterpret_cast<const unsigned int*>(p.x0 + off1 + tbase + t));
        }
      }
    }
    // ---- inverse radix-16 passes back up
    if constexpr (S0 >= 256) {
      f2::pass16<N, 256, true, TH>(z, tA, tB, tid);
      __syncthreads();
    }
    if constexpr (S0 >= 4096) {
      f2::pass16<N, 4096, true, TH>(z, tA, tB, tid);
      __syncthreads();
    }
    // ---- phase Z: last inverse radix-R0 pass fused with the output (only r >= R0/2 survive overlap-save)
    {
      constexpr int PAIRS = S0 / 2;
#pragma unroll
      for (int zi = 0; zi < ZITER; ++zi) {
        const int i = tid + zi * TH;
        if (i >= PAIRS) break;
        const int j = 2 * i;
        float oa[HALF][2], ob[HALF][2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float2 w1 = lead_twiddle(j + u);
          cx y[R0];
#pragma unroll
          for (int q = 0; q < R0; ++q) y[q] = z[fft::pad_idx(j + u + q * S0)];
          cx xo[HALF];
          if constexpr (R0 == 2) {
            xo[0] = f2::csub(y[0], f2::cmulc(y[1], w1));                       // x1 = Y0 - conj(w) Y1
          } else if constexpr (R0 == 4) {
            const float2 w2 = fft::cmul(w1, w1), w3 = fft::cmul(w2, w1);
            const cx y1 = f2::cmulc(y[1], w1), y2 = f2::cmulc(y[2], w2), y3 = f2::cmulc(y[3], w3);
            // x2 = Y0 - Y1 + Y2 - Y3 ;  x3 = Y0 - i Y1 - Y2 + i Y3   (inverse: W4^-1 = +i)
            const cx s02 = f2::cadd(y[0], y2), d02 = f2::csub(y[0], y2);
            const cx s13 = f2::cadd(y1, y3), d13 = f2::mul_mi<true>(f2::csub(y1, y3));   // i (Y1 - Y3)
            xo[0] = f2::csub(s02, s13);
            xo[1] = f2::csub(d02, d13);
          } else if constexpr (R0 == 16) {
#pragma unroll
            for (int q = 1; q < 16; ++q) y[q] = f2::cmulc(y[q], f2::tw_lookup<N>(tA, tB, j + u, q));
            f2::dft16<true>(y);
#pragma unroll
            for (int r = 0; r < HALF; ++r) xo[r] = y[HALF + r];
          } else {
            float2 yf[8];
            float2 w = w1;
            { float a_, b_; f2::un(y[0], a_, b_); yf[0] = make_float2(a_, b_); }
#pragma unroll
            for (int q = 1; q < 8; ++q) {
              float a_, b_;
              f2::un(y[q], a_, b_);
              yf[q] = fft::cmul_conj(make_float2(a_, b_), w);
              w = fft::cmul(w, w1);
            }
            fft::Dft<8, true>::run(yf);
#pragma unroll
            for (int r = 0; r < HALF; ++r) xo[r] = f2::mk(yf[HALF + r].x, yf[HALF + r].y);
          }
#pragma unroll
          for (int r = 0; r < HALF; ++r) {
            float ya, yb;
            f2::un(xo[r], ya, yb);
            const float xa_ = u ? __uint_as_float(gxa[zi][r] & 0xffff0000u) : __uint_as_float(gxa[zi][r] << 16);
            const float xb_ = u ? __uint_as_float(gxb[zi][r] & 0xffff0000u) : __uint_as_float(gxb[zi][r] << 16);
            oa[r][u] = ya * xa_;
            ob[r][u] = yb * xb_;
          }
        }
#pragma unroll
        for (int r = 0; r < HALF; ++r) {
          const int t = tbase + j + r * S0;
          if (t + 1 < t_fft) {
            __nv_bfloat162 va2 = __floats2bfloat162_rn(oa[r][0], oa[r][1]);
            *reinterpret_cast<__nv_bfloat162*>(p.out + off0 + t) = va2;
            if (has_b1) {
              __nv_bfloat162 vb2 = __floats2bfloat162_rn(ob[r][0], ob[r][1]);
              *reinterpret_cast<__nv_bfloat162*>(p.out + off1 + t) = vb2;
            }
          } else if (t < t_fft) {
            p.out[off0 + t] = __float2bfloat16(oa[r][0]);
            if (has_b1) p.out[off1 + t] = __float2bfloat16(ob[r][0]);
          }
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();   // this unit's reads of z are complete and the next unit's staged rows have landed
    }  // chunk loop
    // ---- ragged end: direct causal dot products for t in [t_fft, T)
    for (int t = t_fft; t < p.T; ++t) {
      const float* kc = p.k + (long long)c * p.Lk;
      float sa = 0.f, sb = 0.f;
      for (int s = tid; s <= t; s += TH) {
        const float kv = kc[s];
        sa += kv * bf16_ld(va + (t - s));
        if (has_b1) sb += kv * bf16_ld(vb + (t - s));
      }
      for (int o = 16; o > 0; o >>= 1) {
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
        sb += __shfl_xor_sync(0xffffffffu, sb, o);
      }
      if ((tid & 31) == 0) {
        red[0][tid >> 5] = sa;
        red[1][tid >> 5] = sb;
      }
      __syncthreads();
      if (tid == 0) {
        float ta = 0.f, tb = 0.f;
        for (int w = 0; w < TH / 32; ++w) {
          ta += red[0][w];
          tb += red[1][w];
        }
        const float dc = p.dbias[c];
        p.out[off0 + t] = __float2bfloat16((ta + dc * bf16_ld(va + t)) * bf16_ld(p.x0 + off0 + t));
        if (has_b1) p.out[off1 + t] = __float2bfloat16((tb + dc * bf16_ld(vb + t)) * bf16_ld(p.x0 + off1 + t));
      }
      __syncthreads();
    }
  }
}

}  // namespace clm
