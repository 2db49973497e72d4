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

#ifdef CLM_EXPERIMENTS   // the first long-convolution kernel and its spectrum tables: cross-check only, not in the product build
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
#endif  // CLM_EXPERIMENTS

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
// streaming 16-byte load (8 bf16), bypassing L1 allocation
__device__ __forceinline__ uint4 ld_nc_16(const __nv_bfloat16* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& r, float (&f)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __nv_bfloat162 v = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    w[i] = *reinterpret_cast<uint32_t*>(&v);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}

#ifdef CLM_EXPERIMENTS
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
      // ---- load chunk (zero beyond T), upper half zero.  8 bf16 (16 B) per sequence per step;
      //      row starts are 128 B aligned (Tp % 64 == 0) and Tp >= any t touched here.
#pragma unroll 2
      for (int i = tid * 8; i < C; i += TH * 8) {
        const int t = tbase + i;
        uint4 ra = make_uint4(0, 0, 0, 0), rb = make_uint4(0, 0, 0, 0);
        if (t < p.T) {
          ra = ld_nc_16(va + t);
          if (has_b1) rb = ld_nc_16(vb + t);
        }
        float fa[8], fb[8];
        unpack_bf16x8(ra, fa);
        unpack_bf16x8(rb, fb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool ok = (t + j) < p.T;
          zs[fft::pad_idx(i + j)] = make_float2(ok ? fa[j] : 0.f, ok ? fb[j] : 0.f);
          zs[fft::pad_idx(C + i + j)] = make_float2(0.f, 0.f);
        }
      }
      __syncthreads();
      fft::fft_forward<LOGN, TH>(zs, tid);
      // ---- frequency domain: Y = sum_{j<=ch} U_j . G_{ch-j}; two bins (one float4) per step
      float2* my_scratch = p.scratch + (long long)blockIdx.x * p.n_chunks * N;
      const float2* g0 = p.gspec + (long long)c * N;
      const long long gstride = (long long)p.D * N;
      const bool park = p.n_chunks > 1 && ch + 1 < p.n_chunks;
#pragma unroll 4
      for (int i = tid * 2; i < N; i += TH * 2) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(g0 + i));
        const int zi = fft::pad_idx(i);
        const float2 u0 = zs[zi], u1 = zs[zi + 1];
        if (park) *reinterpret_cast<float4*>(my_scratch + (long long)ch * N + i) = make_float4(u0.x, u0.y, u1.x, u1.y);
        float2 a0 = fft::cmul(u0, make_float2(g.x, g.y)), a1 = fft::cmul(u1, make_float2(g.z, g.w));
        for (int j = 0; j < ch; ++j) {
          const float4 uj = *reinterpret_cast<const float4*>(my_scratch + (long long)j * N + i);
          const float4 gj = __ldg(reinterpret_cast<const float4*>(g0 + (long long)(ch - j) * gstride + i));
          a0.x += uj.x * gj.x - uj.y * gj.y;
          a0.y += uj.x * gj.y + uj.y * gj.x;
          a1.x += uj.z * gj.z - uj.w * gj.w;
          a1.y += uj.z * gj.w + uj.w * gj.z;
        }
        zs[zi] = a0;
        zs[zi + 1] = a1;
      }
      __syncthreads();
      fft::fft_inverse<LOGN, TH>(zs, tid);
      // ---- outputs y[tbase + r] = z[C + r]; add bias skip, apply x0 gate, store bf16 (8 per step)
#pragma unroll 2
      for (int r = tid * 8; r < C; r += TH * 8) {
        const int t = tbase + r;
        if (t < t_fft) {
          const uint4 rva = ld_nc_16(va + t), rxa = ld_nc_16(p.x0 + off0 + t);
          uint4 rvb = make_uint4(0, 0, 0, 0), rxb = make_uint4(0, 0, 0, 0);
          if (has_b1) {
            rvb = ld_nc_16(vb + t);
            rxb = ld_nc_16(p.x0 + off1 + t);
          }
          float fva[8], fxa[8], fvb[8], fxb[8], oa[8], ob[8];
          unpack_bf16x8(rva, fva);
          unpack_bf16x8(rxa, fxa);
          unpack_bf16x8(rvb, fvb);
          unpack_bf16x8(rxb, fxb);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float2 y = zs[fft::pad_idx(C + r + j)];
            oa[j] = (y.x + dc * fva[j]) * fxa[j];
            ob[j] = (y.y + dc * fvb[j]) * fxb[j];
          }
          if (t + 8 <= t_fft) {
            *reinterpret_cast<uint4*>(p.out + off0 + t) = pack_bf16x8(oa);
            if (has_b1) *reinterpret_cast<uint4*>(p.out + off1 + t) = pack_bf16x8(ob);
          } else {
            for (int j = 0; j < 8 && t + j < t_fft; ++j) {
              p.out[off0 + t + j] = __float2bfloat16(oa[j]);
              if (has_b1) p.out[off1 + t + j] = __float2bfloat16(ob[j]);
            }
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

#endif  // CLM_EXPERIMENTS

}  // namespace clm
