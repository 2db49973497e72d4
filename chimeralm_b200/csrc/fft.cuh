// Shared-memory complex FFT building blocks for the long convolution.
//
// Replaces `torch.fft.rfft / irfft` inside HF HyenaDNA `fftconv` (call chain
// chimeralm/models/components/hyena.py:249 -> HyenaOperator -> HyenaFilter.forward).
//
// Design: in-place mixed-radix FFT over N = 2^LOGN complex fp32 points held in shared memory.
//   forward  : decimation-in-frequency, natural order in  -> digit-reversed order out
//   inverse  : the exact mirror,        digit-reversed in -> natural order out (unscaled)
// Because the only thing done in the frequency domain is a pointwise product with a filter
// spectrum produced by the SAME forward transform, no reordering pass is ever needed.
// Radix plan: one leading pass of radix 2^(LOGN%4) (skipped when 1) followed by radix-16 passes,
// each radix-16 butterfly done in registers as 4x4.  Elements are padded by one slot every 16
// (index + index/16) which makes every pass bank-conflict-free for 64-bit accesses.
#pragma once
#include <cuda_runtime.h>

// host+device so the transform can be unit-tested on the CPU (tests/test_fft_host.py)
#define CLM_HD __host__ __device__ __forceinline__
#ifdef __CUDA_ARCH__
#define CLM_SYNC() __syncthreads()
#else
#define CLM_SYNC() ((void)0)
#endif

namespace clm {
namespace fft {

__host__ __device__ constexpr int pad_idx(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int padded_size(int n) { return n + (n >> 4); }

CLM_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
CLM_HD float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
CLM_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
CLM_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
CLM_HD float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
CLM_HD void dft2(float2& a, float2& b) {
  float2 t = a;
  a = cadd(t, b);
  b = csub(t, b);
}

// natural-order 4-point DFT: X_k = sum_n x_n W4^{nk}, W4 = -i (forward) / +i (inverse)
template <bool INV>
CLM_HD void dft4(float2& a, float2& b, float2& c, float2& d) {
  float2 s0 = cadd(a, c), d0 = csub(a, c);
  float2 s1 = cadd(b, d), d1 = mul_mi<INV>(csub(b, d));
  a = cadd(s0, s1);
  c = csub(s0, s1);
  b = cadd(d0, d1);
  d = csub(d0, d1);
}

// W_16^e for e in [0,16): (cos(2*pi*e/16), -sin(2*pi*e/16)); inverse uses the conjugate.
template <bool INV, int E>
CLM_HD float2 mul_w16(float2 a) {
  constexpr float C1 = 0.9238795325112867f, S1 = 0.3826834323650898f, H = 0.7071067811865476f;
  constexpr int e = E & 15;
  if constexpr (e == 0) return a;
  if constexpr (e == 4) return mul_mi<INV>(a);
  if constexpr (e == 8) return make_float2(-a.x, -a.y);
  if constexpr (e == 12) return mul_mi<!INV>(a);
  constexpr float cr = (e == 1 || e == 15) ? C1 : (e == 2 || e == 14) ? H : (e == 3 || e == 13) ? S1
                     : (e == 5 || e == 11) ? -S1 : (e == 6 || e == 10) ? -H : -C1;           // e == 7 || e == 9
  constexpr float sn = (e == 1 || e == 7) ? S1 : (e == 2 || e == 6) ? H : (e == 3 || e == 5) ? C1
                     : (e == 9 || e == 15) ? -S1 : (e == 10 || e == 14) ? -H : -C1;          // e == 11 || e == 13
  // forward twiddle = (cr, -sn); inverse = (cr, +sn)
  constexpr float wi = INV ? sn : -sn;
  return make_float2(a.x * cr - a.y * wi, a.x * wi + a.y * cr);
}

// Natural-order R-point DFT in registers, R in {2,4,8,16}.
template <int R, bool INV>
struct Dft;

template <bool INV>
struct Dft<2, INV> {
  static CLM_HD void run(float2 (&x)[2]) { dft2<INV>(x[0], x[1]); }
};
template <bool INV>
struct Dft<4, INV> {
  static CLM_HD void run(float2 (&x)[4]) { dft4<INV>(x[0], x[1], x[2], x[3]); }
};
// n = n1 + 2 n2 (n1<2, n2<4), k = k2 + 4 k1:  X[k2+4k1] = sum_n1 W2^{n1 k1} W8^{n1 k2} DFT4_n2(x[n1+2n2])[k2]
template <bool INV>
struct Dft<8, INV> {
  static CLM_HD void run(float2 (&x)[8]) {
    dft4<INV>(x[0], x[2], x[4], x[6]);  // n1 = 0 -> results at k2 slots 0,2,4,6
    dft4<INV>(x[1], x[3], x[5], x[7]);  // n1 = 1
    x[3] = mul_w16<INV, 2>(x[3]);       // W8^1
    x[5] = mul_w16<INV, 4>(x[5]);       // W8^2
    x[7] = mul_w16<INV, 6>(x[7]);       // W8^3
    // t[n1][k2] lives at x[n1 + 2 k2]; outer DFT2 over n1 gives X[k2] and X[k2+4]
    float2 y[8];
#pragma unroll
    for (int k2 = 0; k2 < 4; ++k2) {
      y[k2] = cadd(x[2 * k2], x[2 * k2 + 1]);
      y[k2 + 4] = csub(x[2 * k2], x[2 * k2 + 1]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = y[i];
  }
};
// n = n1 + 4 n2, k = k2 + 4 k1:  X[k2+4k1] = sum_n1 W4^{n1 k1} W16^{n1 k2} DFT4_n2(x[n1+4n2])[k2]
template <bool INV>
struct Dft<16, INV> {
  static CLM_HD void run(float2 (&x)[16]) {
    // inner DFT4 over n2 for each n1; t[n1][k2] stored at x[n1 + 4 k2]
    dft4<INV>(x[0], x[4], x[8], x[12]);
    dft4<INV>(x[1], x[5], x[9], x[13]);
    dft4<INV>(x[2], x[6], x[10], x[14]);
    dft4<INV>(x[3], x[7], x[11], x[15]);
    // twiddle W16^{n1 k2}
    x[1 + 4 * 1] = mul_w16<INV, 1>(x[5]);
    x[1 + 4 * 2] = mul_w16<INV, 2>(x[9]);
    x[1 + 4 * 3] = mul_w16<INV, 3>(x[13]);
    x[2 + 4 * 1] = mul_w16<INV, 2>(x[6]);
    x[2 + 4 * 2] = mul_w16<INV, 4>(x[10]);
    x[2 + 4 * 3] = mul_w16<INV, 6>(x[14]);
    x[3 + 4 * 1] = mul_w16<INV, 3>(x[7]);
    x[3 + 4 * 2] = mul_w16<INV, 6>(x[11]);
    x[3 + 4 * 3] = mul_w16<INV, 9>(x[15]);
    // outer DFT4 over n1 for each k2: inputs x[0+4k2..3+4k2] -> X[k2 + 4 k1] for k1 = 0..3
    dft4<INV>(x[0], x[1], x[2], x[3]);      // k2 = 0: X[0], X[4], X[8], X[12]
    dft4<INV>(x[4], x[5], x[6], x[7]);      // k2 = 1: X[1], X[5], X[9], X[13]
    dft4<INV>(x[8], x[9], x[10], x[11]);    // k2 = 2
    dft4<INV>(x[12], x[13], x[14], x[15]);  // k2 = 3
    // now x[k1 + 4 k2] holds X[k2 + 4 k1]: transpose the 4x4 index grid
    float2 t;
#define CLM_SWAP(a, b) t = x[a]; x[a] = x[b]; x[b] = t;
    CLM_SWAP(1, 4) CLM_SWAP(2, 8) CLM_SWAP(3, 12) CLM_SWAP(6, 9) CLM_SWAP(7, 13) CLM_SWAP(11, 14)
#undef CLM_SWAP
  }
};

// One in-place pass over z[N] (padded indexing).  Block size NB, radix R, sub-stride S = NB / R.
// Forward: y_q[j] = W_NB^{j q} * DFT_R(x[j + S r])[q], stored at (q S + j).
// Inverse: x[j + S r] = IDFT_R( conj(W_NB)^{j q} * y_q[j] )[r]   (unscaled).
template <int N, int NB, int R, bool INV, int THREADS>
CLM_HD void fft_pass(float2* z, int tid) {
  constexpr int S = NB / R;
  constexpr int NBF = N / R;  // butterflies
#pragma unroll 1
  for (int i = tid; i < NBF; i += THREADS) {
    const int j = i % S;
    const int base = (i / S) * NB + j;
    float2 x[R];
#pragma unroll
    for (int r = 0; r < R; ++r) x[r] = z[pad_idx(base + r * S)];
    if constexpr (S > 1) {
      // w1 = exp(-+ 2 pi i j / NB); j / NB is exact in fp32, sincospif reduces exactly
      float sn, cs;
      sincospif(2.0f * (float)j / (float)NB, &sn, &cs);
      float2 w[R];
      w[1] = make_float2(cs, -sn);  // forward twiddle
      if constexpr (R > 2) {
#pragma unroll
        for (int q = 2; q < R; ++q) w[q] = (q & 1) ? cmul(w[q - 1], w[1]) : cmul(w[q / 2], w[q / 2]);
      }
      if constexpr (!INV) {
        Dft<R, false>::run(x);
#pragma unroll
        for (int q = 1; q < R; ++q) x[q] = cmul(x[q], w[q]);
      } else {
#pragma unroll
        for (int q = 1; q < R; ++q) x[q] = cmul_conj(x[q], w[q]);
        Dft<R, true>::run(x);
      }
    } else {
      Dft<R, INV>::run(x);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) z[pad_idx(base + r * S)] = x[r];
  }
}

template <int LOGN>
struct Plan {
  static constexpr int N = 1 << LOGN;
  static constexpr int R0 = 1 << (LOGN % 4);     // leading radix (1 = none)
  static constexpr int K16 = LOGN / 4;           // number of radix-16 passes
};

// radix-16 passes p = 0..K16-1 over blocks NB = N/R0 / 16^p
template <int N, int NB, bool INV, int THREADS>
CLM_HD void radix16_chain_fwd(float2* z, int tid) {
  if constexpr (NB >= 16) {
    fft_pass<N, NB, 16, false, THREADS>(z, tid);
    CLM_SYNC();
    radix16_chain_fwd<N, NB / 16, INV, THREADS>(z, tid);
  }
}
template <int N, int NB, int NBTOP, int THREADS>
CLM_HD void radix16_chain_inv(float2* z, int tid) {
  // runs passes from the smallest block (16) up to NBTOP
  if constexpr (NB <= NBTOP) {
    fft_pass<N, NB, 16, true, THREADS>(z, tid);
    CLM_SYNC();
    radix16_chain_inv<N, NB * 16, NBTOP, THREADS>(z, tid);
  }
}

// Forward FFT of z (natural order) -> digit-reversed spectrum.  Ends with __syncthreads().
template <int LOGN, int THREADS>
CLM_HD void fft_forward(float2* z, int tid) {
  using P = Plan<LOGN>;
  if constexpr (P::R0 > 1) {
    fft_pass<P::N, P::N, P::R0, false, THREADS>(z, tid);
    CLM_SYNC();
  }
  radix16_chain_fwd<P::N, P::N / P::R0, false, THREADS>(z, tid);
}
// Inverse of fft_forward, unscaled (caller folds 1/N into the filter spectrum).
template <int LOGN, int THREADS>
CLM_HD void fft_inverse(float2* z, int tid) {
  using P = Plan<LOGN>;
  radix16_chain_inv<P::N, 16, P::N / P::R0, THREADS>(z, tid);
  if constexpr (P::R0 > 1) {
    fft_pass<P::N, P::N, P::R0, true, THREADS>(z, tid);
    CLM_SYNC();
  }
}

}  // namespace fft
}  // namespace clm
