// Hyena implicit long filter + causal long convolution, fused with both gates' tail end.
//
// Reference semantics (HF HyenaDNA remote code; call site
// chimeralm/models/components/hyena.py:249):
//   k      = HyenaFilter.filter(T)                        [D, T]   (filter_gen_kernel)
//   y[c,t] = sum_{s<=t} k[c,s] * vx[c,t-s] + bias[c] * vx[c,t]      (fftconv, fp32)
//   out    = y * x0                                                 (second gate)
//
// Layout: vx, x0, out are channel-major bf16 [B][D][Tp] (time contiguous, Tp = T rounded up).
// One work item = (channel c, pair of batch rows): two real sequences ride one complex FFT
// (re = row 2p, im = row 2p+1); convolution with a real filter keeps them separate.
//
// Long sequences use overlap-save over chunks of C = N/2 outputs:
//   Y_i = sum_{j<=i} U_j . G_{i-j},   y_i = IFFT(Y_i)[C..2C)
// with U_j the spectrum of input chunk j zero-padded to N and G_m the spectrum of the filter
// segment g_m[p] = k[mC + p - C], p in [1,2C) (zero where the index is negative).  Spectra of
// earlier chunks are parked in a per-CTA global scratch (L2-resident).  A short ragged end
// (T mod C <= LONGCONV_TAIL_MAX, e.g. the trailing [SEP] of a maximum-length read) is finished
// by direct dot products instead of paying for another FFT chunk.
#pragma once
#include <cuda_bf16.h>

#include "fft.cuh"

namespace clm {

constexpr int LONGCONV_TAIL_MAX = 8;
constexpr int LONGCONV_MAX_LOGN = 14;  // 16384 complex points = 139 KB padded smem

template <int LOGN>
struct ConvCfg {
  static constexpr int N = 1 << LOGN;
  static constexpr int C = N / 2;
  static constexpr int THREADS = (N / 32) < 64 ? 64 : ((N / 32) > 512 ? 512 : (N / 32));
  static constexpr int SMEM = fft::padded_size(N) * (int)sizeof(float2);
};

// ---------------------------------------------------------------------------------------------
// Implicit filter: one block per time step t computes k[:, t] for one layer.
//   z[t,0:E] -> sin(f*(W0 z + b0)) -> sin(f*(W1 h + b1)) -> sin(f*(W2 h + b2)) -> W3 h   [D]
//   k[c,t] = h[c] * (exp(-tpos[t] * |delta[c]|) + shift)
struct FilterGenParams {
  const float* z;       // [Lmax, E]
  const float* tpos;    // [Lmax]
  const float* w[4];    // w[0]: [F,E], w[1..2]: [F,F], w[3]: [D,F]
  const float* b[3];    // [F]
  const float* freq[3]; // [F]
  const float* deltas;  // [D]
  float shift;
  int E, F, D, L;
  float* k_out;         // [D, Lk]
  long long Lk;
};

__global__ void __launch_bounds__(256) filter_gen_kernel(FilterGenParams p) {
  __shared__ float h0[64], h1[64];
  const int t = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < p.F) {
    float a = p.b[0][tid];
    for (int e = 0; e < p.E; ++e) a += p.w[0][tid * p.E + e] * p.z[(long long)t * p.E + e];
    h0[tid] = sinf(p.freq[0][tid] * a);
  }
  __syncthreads();
  if (tid < p.F) {
    float a = p.b[1][tid];
    for (int e = 0; e < p.F; ++e) a += p.w[1][tid * p.F + e] * h0[e];
    h1[tid] = sinf(p.freq[1][tid] * a);
  }
  __syncthreads();
  if (tid < p.F) {
    float a = p.b[2][tid];
    for (int e = 0; e < p.F; ++e) a += p.w[2][tid * p.F + e] * h1[e];
    h0[tid] = sinf(p.freq[2][tid] * a);
  }
  __syncthreads();
  for (int c = tid; c < p.D; c += blockDim.x) {
    float a = 0.f;
    for (int e = 0; e < p.F; ++e) a += p.w[3][c * p.F + e] * h0[e];
    float decay = expf(-p.tpos[t] * fabsf(p.deltas[c]));
    p.k_out[(long long)c * p.Lk + t] = a * (decay + p.shift);
  }
}

// ---------------------------------------------------------------------------------------------
// Filter segment spectra: gspec[m][c][0..N) = FFT_N(g_m) / N in the forward transform's
// digit-reversed order, for m in [0, n_seg).  Grid (D, n_seg).
template <int LOGN>
__global__ void __launch_bounds__(ConvCfg<LOGN>::THREADS) filter_spectrum_kernel(const float* __restrict__ k,
                                                                                  long long Lk, int L,
                                                                                  float2* __restrict__ gspec, int D) {
  using Cfg = ConvCfg<LOGN>;
  constexpr int N = Cfg::N, C = Cfg::C, TH = Cfg::THREADS;
  extern __shared__ float2 zs[];
  const int c = blockIdx.x, m = blockIdx.y, tid = threadIdx.x;
  const float* kc = k + (long long)c * Lk;
  for (int pidx = tid; pidx < N; pidx += TH) {
    long long s = (long long)m * C + pidx - C;
    float v = (pidx >= 1 && s >= 0 && s < L) ? kc[s] : 0.f;
    zs[fft::pad_idx(pidx)] = make_float2(v, 0.f);
  }
  __syncthreads();
  fft::fft_forward<LOGN, TH>(zs, tid);
  float2* out = gspec + ((long long)m * D + c) * N;
  const float sc = 1.0f / (float)N;
  for (int i = tid; i < N; i += TH) {
    float2 v = zs[fft::pad_idx(i)];
    out[i] = make_float2(v.x * sc, v.y * sc);
  }
}

// ---------------------------------------------------------------------------------------------
struct LongConvParams {
  const __nv_bfloat16* vx;   // [B][D][Tp]
  const __nv_bfloat16* x0;   // [B][D][Tp]
  __nv_bfloat16* out;        // [B][D][Tp]
  const float2* gspec;       // [n_seg][D][N]
  const float* k;            // [D][Lk] time-domain filter (ragged-end dot products)
  const float* dbias;        // [D]
  float2* scratch;           // [gridDim.x][n_chunks][N]
  long long Lk;
  int B, D, T, Tp;
  int n_chunks;              // FFT chunks; outputs [0, n_chunks*C) come from the FFT path
  int n_items;               // D * ceil(B/2)
};

__device__ __forceinline__ float bf16_ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }

template <int LOGN>
__global__ void __launch_bounds__(ConvCfg<LOGN>::THREADS, 1) longconv_kernel(LongConvParams p) {
  using Cfg = ConvCfg<LOGN>;
  constexpr int N = Cfg::N, C = Cfg::C, TH = Cfg::THREADS;
  extern __shared__ float2 zs[];
  __shared__ float red[2][TH / 32];
  const int tid = threadIdx.x;
  const int n_pairs = (p.B + 1) / 2;
  const int t_fft = min(p.n_chunks * C, p.T);  // outputs produced by the FFT path

  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
    const int c = item / n_pairs;
    const int b0 = (item % n_pairs) * 2;
    const bool has_b1 = (b0 + 1) < p.B;
    const long long off0 = ((long long)b0 * p.D + c) * p.Tp;
    const long long off1 = off0 + (long long)p.D * p.Tp;
    const __nv_bfloat16* va = p.vx + off0;
    const __nv_bfloat16* vb = p.vx + off1;
    const float dc = p.dbias[c];

    for (int ch = 0; ch < p.n_chunks; ++ch) {
      const int tbase = ch * C;
      // ---- load chunk (zero beyond T), upper half zero
      for (int i = tid; i < C; i += TH) {
        const int t = tbase + i;
        float a = 0.f, b = 0.f;
        if (t < p.T) {
          a = bf16_ld(va + t);
          if (has_b1) b = bf16_ld(vb + t);
        }
        zs[fft::pad_idx(i)] = make_float2(a, b);
        zs[fft::pad_idx(C + i)] = make_float2(0.f, 0.f);
      }
      __syncthreads();
      fft::fft_forward<LOGN, TH>(zs, tid);
      // ---- frequency domain: Y = sum_{j<=ch} U_j . G_{ch-j}
      float2* my_scratch = p.scratch + (long long)blockIdx.x * p.n_chunks * N;
      const float2* g0 = p.gspec + (long long)c * N;
      const long long gstride = (long long)p.D * N;
      for (int i = tid; i < N; i += TH) {
        const float2 u = zs[fft::pad_idx(i)];
        if (p.n_chunks > 1 && ch + 1 < p.n_chunks) my_scratch[(long long)ch * N + i] = u;
        float2 acc = fft::cmul(u, g0[i]);
        for (int j = 0; j < ch; ++j) {
          const float2 uj = my_scratch[(long long)j * N + i];
          const float2 g = g0[(long long)(ch - j) * gstride + i];
          acc.x += uj.x * g.x - uj.y * g.y;
          acc.y += uj.x * g.y + uj.y * g.x;
        }
        zs[fft::pad_idx(i)] = acc;
      }
      __syncthreads();
      fft::fft_inverse<LOGN, TH>(zs, tid);
      // ---- outputs y[tbase + r] = z[C + r]; add bias skip, apply x0 gate, store bf16
      for (int r = tid; r < C; r += TH) {
        const int t = tbase + r;
        if (t < t_fft) {
          const float2 y = zs[fft::pad_idx(C + r)];
          const float oa = (y.x + dc * bf16_ld(va + t)) * bf16_ld(p.x0 + off0 + t);
          p.out[off0 + t] = __float2bfloat16(oa);
          if (has_b1) {
            const float ob = (y.y + dc * bf16_ld(vb + t)) * bf16_ld(p.x0 + off1 + t);
            p.out[off1 + t] = __float2bfloat16(ob);
          }
        }
      }
      __syncthreads();
    }
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
        const float oa = (ta + dc * bf16_ld(va + t)) * bf16_ld(p.x0 + off0 + t);
        p.out[off0 + t] = __float2bfloat16(oa);
        if (has_b1) {
          const float ob = (tb + dc * bf16_ld(vb + t)) * bf16_ld(p.x0 + off1 + t);
          p.out[off1 + t] = __float2bfloat16(ob);
        }
      }
      __syncthreads();
    }
  }
}

}  // namespace clm
