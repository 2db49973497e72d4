// Long convolution on the tensor cores: a 16384-point FFT as two 128-point DFT matrix products
// (Monarch / four-step decomposition, N = 128 x 128), all four matrix stages on tcgen05 with fp16
// operands and fp32 accumulation in TMEM.  Same math as longconv_fast_kernel<14> for reads of
// 2057..8200 tokens (longer reads: chunked form below): y = causal_conv(vx, k) + bias * vx (bias folded into tap 0), out = y * x0;
// replaces fftconv() (reference chimeralm/models/components/hyena.py via HF modeling_hyena.fftconv,
// SURVEY.md A.5).
//
// One item = one channel of two reads, z = vx_a + i vx_b (the filter is real, so the two convolutions
// come back as the real and imaginary parts; no spectrum unpacking).  With n = 128 n1 + n2 and
// k = k1 + 128 k2, F = the 128-point DFT matrix and tw[k1][n2] = exp(-2 pi i k1 n2 / N):
//   step 1   A[k1][n2] = sum_{n1 < 64} F[k1][n1] z[n1][n2]          (rows n1 >= 64 are the zero padding)
//   E1       P1 = fp16(s1 * tw .* A)
//   step 3   S[k1][k2] = sum_{n2} P1[k1][n2] F[n2][k2]              (the spectrum, fp32 in TMEM)
//   E2       P2 = fp16(S .* G'),  G' = FFT(k) / (N s1) in the same [k1][k2] order (fp16 table)
//   step 5   B[k1][n2] = sum_{k2} P2[k1][k2] conj(F)[k2][n2]
//   E3       BT = fp16(conj(tw) .* B)  -> shared memory, [k1][n2] with n2 contiguous
//   step 7   z'[n2][n1] = sum_{k1} BT[k1][n2] conj(F)[k1][n1], n1 < 64 (outputs n >= 8192 are discarded)
//   E4       out[t = 128 n1 + n2] = z' * x0 -> bf16, global
// Operand forms: step 1 = A constants (shared, K-major) x B data (shared, MN-major, written by TMA);
// steps 3/5 = A data from TMEM (packed fp16) x B constants; step 7 = A data (shared, MN-major) x B
// constants.  All constant operands are slices of ONE 96 KB stack S = [-Im F; Re F; Im F]
// (384 rows x 128 K, fp16, 128-byte swizzled K-major), resident in shared memory.
// Error of the fp16 operand rounding: ~4e-4 relative L2 on real layer inputs, a third of the bf16
// rounding the output gets anyway (tests/probes/fp16_monarch_fft_emulation.py).
//
// TMEM: X = cols [0,256), Y = cols [256,512).  step 1 -> X (A_re | A_im); E1 packs P1 in place over
// X[0,128); step 3 -> Y (S_re | S_im); E2 packs P2 over Y[0,128); step 5 -> X (B_im | B_re);
// step 7 -> Y[0,128) (z'_re | z'_im).  Packed K order (both P1 and P2): per 32-index chunk c, 16
// columns of re pairs then 16 columns of im pairs, so each thread overwrites only columns it has
// just read.
//
// Warp roles: warp 0 = TMA producer (z tiles, double buffered), warp 1 = MMA issuer, warps 2..9 =
// epilogue (TMEM lane quarter = warp % 4, index half = (warp - 2) / 4).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_tcgen05.cuh"
#include "ptx.cuh"

namespace clm {

struct LongConvTcParams {
  const __nv_bfloat16* x0;   // [B][D][Tp] gate
  __nv_bfloat16* out;        // [B][D][Tp]
  const uint4* S;            // 96 KB shared-memory image of the constant stack (tc::build_s_kernel)
  const uint4* G;            // per channel 4096 uint4: fp16 (re, im) spectrum, lane-interleaved (tc::spectrum_kernel)
  int B, D, Tp, n_items, n_pairs;
  int T;                     // tokens; outputs t in [8192, T) are the tail (direct dot products inside the kernel)
  const __half* vx;          // [B][D][Tp] fp16 (same buffer tmVX describes; the tail reads x[t >= 8192] from it)
  const float* k;            // [D][Lk] filter taps
  const float* dbias;        // [D] bias skip (folded into tap 0)
  long long Lk;
  // chunked mode (reads longer than 8200 tokens, overlap-add over chunks of 8192 tokens): n_chunks > 1
  int n_chunks;              // transforms per item
  int nt;                    // tail tokens after the last chunk (0..LONGCONV_TAIL_MAX), finished by direct products
  float* scratch;            // per CTA: (n_chunks - 1) parked spectra (16384 x fp16 (re, im) each) + carry (2 x 8192 float)
  long long scratch_per_cta; // floats
  long long g_seg_stride;    // uint4 between the spectrum tables of consecutive filter segments
  // Dynamic-range handling (all factors are powers of two, so they are exact): the input rows hold a[ch] * v*x1, the
  // spectrum table of segment j holds FFT(k_j) * 2^e[j][ch] (largest bin in [1/16, 1/8)); `osc` undoes both and the
  // transform's N * S1, `rel[j]` = 2^(e[0] - e[j]) brings the later filter segments to segment 0's scale before the
  // sum over segments, `inva` = 1 / a is for the tail tokens' direct products.
  const float* osc;          // [D]
  const float* inva;         // [D]
  const float* rel;          // [n_seg][D]
  int helpers_low;           // longconv_tc2_kernel, A/B switch: 1 = helper warps on the LOWEST warp ids (default 0: highest)
  const float* osc_adj;      // [D] four-reads-per-item form only: 2^(e[0][ch] - e4[ch]), e4 = exponent of the 4096-tap table
  int* err;                  // status word: bit 1 (value 2) is set when an output is not finite (fp16 range exceeded)
  long long* trace;          // optional [2][64] clock64 stamps of CTA 0: row 0 = MMA issuer, row 1 = epilogue warp 2
};

namespace tc {
constexpr int R = 128, N = R * R, C = N / 2;
constexpr float S1 = 0.125f;                      // scale of P1 (keeps |A| far from the fp16 limits)
constexpr int S_ROWS = 384;
constexpr int S_PANEL = S_ROWS * 128;             // bytes per 64-wide K panel
constexpr int S_BYTES = 2 * S_PANEL;              // 98304
constexpr int Z_BYTES = 32768;                    // [re | im][atom 2][64 rows][128 B]
constexpr int BT_ATOM = 256 * 128;                // 32 KB: 256 K rows x 64 n2
constexpr int BT_BYTES = 2 * BT_ATOM;
constexpr int OFF_S = 0;
constexpr int OFF_Z = OFF_S + S_BYTES;            // 2 buffers
constexpr int OFF_BT = OFF_Z + 2 * Z_BYTES;
constexpr int OFF_BAR = OFF_BT + BT_BYTES;        // 229376
constexpr int SMEM_TOTAL = OFF_BAR + 256;
constexpr int THREADS = 320;        // single-chunk kernel: warp 0 producer, warp 1 MMA issuer, warps 2..9 epilogue
// Chunked kernel: three warpgroups (producer, MMA issuer and two idle warps | eight epilogue warps) so that `setmaxnreg` can
// move registers to the epilogue warps (200 each instead of the 168 ptxas allows 320 threads): the segment sum and the
// carry of the overlap-add keep ~60 more values live than the single-chunk epilogue.
constexpr int THREADS_CH = 384;
constexpr int ROW_NFIM = 0, ROW_FRE = 128, ROW_FIM = 256;   // row blocks of S
constexpr int ZIM_COL = 80;                                 // step 7: z'_re at Y[0, 80), z'_im at Y[80, 160)

// kind::f16 instruction descriptor, fp16 x fp16 -> fp32 (formats 0), M = 128.
__host__ __device__ constexpr uint32_t idesc(uint32_t n, bool a_mn, bool b_mn) {
  return (1u << 4) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

__device__ __forceinline__ f2t h2_to_f2(uint32_t h) {
  const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&h));
  return f2_pack(f.x, f.y);
}

// The constant stack as it sits in shared memory (byte image): panel p (K = 64 p .. 64 p + 63), row r
// (128 B), 16-byte chunk j stored at chunk position j ^ (r & 7).
__global__ void build_s_kernel(__half* __restrict__ img) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S_ROWS * 128) return;
  const int r = i / 128, k = i % 128;
  const int blk = r / 128, row = r % 128;
  float s, c;
  sincospif(-2.0f * float((row * k) % 128) / 128.0f, &s, &c);   // F = c + i s
  const float v = blk == 0 ? -s : (blk == 1 ? c : s);
  const int p = k / 64, kk = k % 64;
  const size_t off = (size_t)p * (S_PANEL / 2) + (size_t)r * 64 + (size_t)(((kk / 8) ^ (r & 7)) * 8 + kk % 8);
  img[off] = __float2half_rn(v);
}

// G'[k1][k2] = FFT_N(k')[k1 + 128 k2] * 2^e, k' = filter taps [seg C, seg C + C) with the bias skip folded into
// tap 0 of segment 0, computed with the same two-stage decomposition in fp32.  e = gexp[seg][ch] is chosen per
// (segment, channel) so that the largest bin lands in [1/16, 1/8): the table is fp16, and with one global scale the
// bins of a decayed segment (or of a trained filter of another magnitude) would sit in the subnormal range.  Stored
// as one uint4 per 4 consecutive k2:
// {re(k2, k2+1), im(k2, k2+1), re(k2+2, k2+3), im(k2+2, k2+3)} at uint4 index
//   (((ch * 4 + k1 / 32) * 2 + k2 / 64) * 16 + (k2 % 64) / 4) * 32 + k1 % 32
// so that a warp of 32 consecutive k1 reads 512 contiguous bytes per 16-byte load.
// vform: the table of segment j also carries segment j - 1 in the SECOND half of the 16384-point window,
//   H_j = FFT([k_j | k_{j-1}]) = G_j + (-1)^k G_{j-1}.
// With these tables the overlap-add over chunks needs no carry: the output chunk m is the FIRST half of
// IFFT(sum_c S_c H_{m-c}), because the second half of x_c * k_j (which belongs to chunk c + j + 1) is the first half of
// the same product shifted by half a period, i.e. of S_c G_j (-1)^k (longconv_tc2_kernel, chunked form).
__global__ void __launch_bounds__(256) spectrum_kernel(const float* __restrict__ k, long long Lk, int n_taps,
                                                       const float* __restrict__ dbias, __half2* __restrict__ G_all,
                                                       int* __restrict__ gexp, int vform = 0) {
  extern __shared__ float2 sm_a[];            // A[k1][n2], 128 KB
  __shared__ float2 w128[128];
  __shared__ float red[8];
  __shared__ float s_scale;
  // blockIdx.y = filter segment: taps [seg * C, (seg + 1) * C) (zero past n_taps); the bias skip lives in tap 0 only
  const int ch = blockIdx.x, seg = blockIdx.y;
  const float* kc = k + (long long)ch * Lk + (long long)seg * C;
  const int n_here = max(0, min(C, n_taps - seg * C));
  __half2* G = G_all + (size_t)seg * gridDim.x * N;
  if (threadIdx.x < 128) {
    float s, c;
    sincospif(-2.0f * float(threadIdx.x) / 128.0f, &s, &c);
    w128[threadIdx.x] = make_float2(c, s);
  }
  __syncthreads();
  const float tap0 = n_here > 0 ? kc[0] + (seg == 0 ? dbias[ch] : 0.f) : 0.f;
  for (int o = threadIdx.x; o < N; o += blockDim.x) {
    const int k1 = o / R, n2 = o % R;
    float ar = 0.f, ai = 0.f;
    for (int n1 = 0; n1 < 64; ++n1) {
      const int t = 128 * n1 + n2;
      const float v = t == 0 ? tap0 : (t < n_here ? kc[t] : 0.f);
      const float2 w = w128[(k1 * n1) & 127];
      ar = fmaf(v, w.x, ar);
      ai = fmaf(v, w.y, ai);
    }
    if (vform && seg > 0) {   // second half of the window: the previous segment (all of its C taps exist; tap 0 of the
      const float* kp = kc - C;   // filter, with the bias skip, sits at the start of segment 0)
      for (int n1 = 64; n1 < 128; ++n1) {
        const int t = 128 * (n1 - 64) + n2;
        const float v = kp[t] + ((seg == 1 && t == 0) ? dbias[ch] : 0.f);
        const float2 w = w128[(k1 * n1) & 127];
        ar = fmaf(v, w.x, ar);
        ai = fmaf(v, w.y, ai);
      }
    }
    float s, c;
    sincospif(-2.0f * float((k1 * n2) % N) / float(N), &s, &c);
    sm_a[o] = make_float2(ar * c - ai * s, ar * s + ai * c);
  }
  __syncthreads();
  auto bin = [&](int o, float& gr, float& gi) {
    const int k1 = o / R, k2 = o % R;
    gr = 0.f; gi = 0.f;
    for (int n2 = 0; n2 < R; ++n2) {
      const float2 a = sm_a[k1 * R + n2];
      const float2 w = w128[(n2 * k2) & 127];
      gr += a.x * w.x - a.y * w.y;
      gi += a.x * w.y + a.y * w.x;
    }
  };
  // pass 1: largest |re|, |im| over the segment's bins -> exponent
  float mx = 0.f;
  for (int o = threadIdx.x; o < N; o += blockDim.x) {
    float gr, gi;
    bin(o, gr, gi);
    mx = fmaxf(mx, fmaxf(fabsf(gr), fabsf(gi)));
  }
  for (int d = 16; d > 0; d >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < 8; ++i) m = fmaxf(m, red[i]);
    int e = 0;
    if (m > 0.f && m < INFINITY) {
      int ex;
      frexpf(m, &ex);            // m = f * 2^ex, f in [0.5, 1)  ->  m * 2^(-3 - ex) in [1/16, 1/8)
      e = max(-100, min(100, -3 - ex));
    }
    gexp[seg * gridDim.x + ch] = e;
    s_scale = ldexpf(1.0f, e);
  }
  __syncthreads();
  const float scale = s_scale;
  // pass 2: recompute (the 128 KB of stage-1 results leave no room to keep the bins) and store scaled
  for (int o = threadIdx.x; o < N; o += blockDim.x) {
    const int k1 = o / R, k2 = o % R;
    float gr, gi;
    bin(o, gr, gi);
    const size_t u4 = ((((size_t)ch * 4 + k1 / 32) * 2 + k2 / 64) * 16 + (k2 % 64) / 4) * 32 + (k1 % 32);
    __half* gh = reinterpret_cast<__half*>(G) + u4 * 8 + ((k2 % 4) / 2) * 4 + (k2 % 2);
    gh[0] = __float2half_rn(gr * scale);
    gh[2] = __float2half_rn(gi * scale);
  }
}

// Largest |vx| per channel of a channel-major bf16 [B][D][Tp] buffer (tokens < T), accumulated with atomicMax on the
// float bit pattern (non-negative floats order like unsigned integers).  Grid (D, B).  Used once, by the calibration
// forward of clm_finalize, and by the unit-level auto-scaling entry point.
__global__ void __launch_bounds__(256) amax_cm_kernel(const __nv_bfloat16* __restrict__ vx, int D, int Tp, int T,
                                                      unsigned int* __restrict__ amax_bits) {
  __shared__ float red[8];
  const int ch = blockIdx.x, b = blockIdx.y;
  const __nv_bfloat16* row = vx + ((long long)b * D + ch) * Tp;
  float m = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float v = fabsf(__bfloat162float(row[t]));
    if (v < INFINITY) m = fmaxf(m, v);
  }
  for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, d));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
    atomicMax(amax_bits + ch, __float_as_uint(m));
  }
}

constexpr float VX_TARGET_AMAX = 8.0f;   // scaled |v*x1| of the calibration draw lands in (4, 8]

// Per-channel scale factors of one layer (see LongConvTcParams).  amax_bits == nullptr: no input scaling (a = 1).
// Budget: with |a z| <= 8 the largest possible spectrum bin (a constant input - e.g. the [PAD] prefix of a left-padded
// batch - puts 8192 equal values into the DC bin) is 8 * 8192 * S1 = 8192 in P1 x F and in the parked fp16 spectra, and
// <= 1024 after the filter product: 8x below the fp16 limit; the input itself has 8000x of headroom over the
// calibration draw.  Anything beyond that overflows to inf/NaN, is detected in E4 and reported (clm_forward_status).
// `shift` (normally 0) moves the input scale by 2^shift: a test hook that makes real data overflow.
__global__ void scales_kernel(const int* __restrict__ gexp, const unsigned int* __restrict__ amax_bits, float* __restrict__ vx_scale,
                              float* __restrict__ osc, float* __restrict__ inva, float* __restrict__ rel, int D, int n_seg,
                              int shift) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= D) return;
  int ea = 0;
  if (amax_bits) {
    const float m = __uint_as_float(amax_bits[ch]);
    if (m > 0.f && m < INFINITY) {
      int ex;
      frexpf(m, &ex);                       // m in [2^(ex-1), 2^ex)  ->  m * 2^(3 - ex) in [4, 8)
      ea = max(-40, min(40, 3 - ex));
    }
  }
  ea += shift;
  const int e0 = gexp[ch];
  vx_scale[ch] = ldexpf(1.0f, ea);
  inva[ch] = ldexpf(1.0f, -ea);
  osc[ch] = ldexpf(1.0f, -e0 - ea - 11);    // 1 / (2^e0 * a * N * S1), N * S1 = 2048
  for (int j = 0; j < n_seg; ++j) rel[j * D + ch] = ldexpf(1.0f, max(-120, min(120, e0 - gexp[j * D + ch])));
}

// rel[j][ch] = 2^(e[0][ch] - e[j][ch]) for a table set of its own (the V-form tables of the chunked two-in-flight kernel)
__global__ void rel_kernel(const int* __restrict__ gexp, float* __restrict__ rel, int D, int n_seg) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= D) return;
  for (int j = 0; j < n_seg; ++j) rel[j * D + ch] = ldexpf(1.0f, max(-120, min(120, gexp[ch] - gexp[j * D + ch])));
}

// adj[ch] = 2^(e[ch] - e4[ch]): brings the output scale of the full-filter table to the 4096-tap table's exponent
__global__ void exp_adj_kernel(const int* __restrict__ gexp, const int* __restrict__ gexp4, float* __restrict__ adj, int D) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch < D) adj[ch] = ldexpf(1.0f, max(-120, min(120, gexp[ch] - gexp4[ch])));
}

// bf16 -> fp16 with the per-channel input scale and zeros past T (what block_in emits in the forward)
__global__ void __launch_bounds__(256) scale_to_f16_kernel(const __nv_bfloat16* __restrict__ in, __half* __restrict__ out,
                                                           const float* __restrict__ vx_scale, int D, int Tp, int T) {
  const int ch = blockIdx.x, b = blockIdx.y;
  const long long base = ((long long)b * D + ch) * Tp;
  const float a = vx_scale[ch];
  for (int t = threadIdx.x; t < Tp; t += blockDim.x) out[base + t] = __float2half_rn(t < T ? __bfloat162float(in[base + t]) * a : 0.f);
}

}  // namespace tc

template <bool CH>   // CH: chunked mode (n_chunks > 1)
__global__ void __launch_bounds__(CH ? tc::THREADS_CH : tc::THREADS, 1)
longconv_tc_kernel(const __grid_constant__ CUtensorMap tmVX, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ CUtensorMap tmX0, LongConvTcParams p) {
  using namespace tc;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* z_full = bars;        // [2] TMA landed
  uint64_t* z_empty = bars + 2;   // [2] step 1 finished reading z
  uint64_t* x_full = bars + 4;    // step 1 accumulators complete
  uint64_t* p1_full = bars + 16;  // [2] E1 wrote the P1 columns of index half h (steps 3 / 5 sum over that index: the MMAs
                                  //     over the first half are issued while the epilogue still packs the second)
  uint64_t* y_full = bars + 6;    // step 3 complete
  uint64_t* p2_full = bars + 18;  // [2] E2 wrote the P2 columns of index half h
  uint64_t* x2_full = bars + 8;   // step 5 complete
  uint64_t* bt_full = bars + 9;   // E3 wrote BT
  uint64_t* o_full = bars + 10;   // step 7 complete
  // [2] x0 gate tile landed in z buffer `buf` (TMA issued by ONE epilogue thread once step 1 has consumed z).  Per buffer on
  // purpose: that thread may run a whole E4 ahead of a slower warp (the one doing the tail-token loads), and with a single
  // barrier it could complete the NEXT item's phase before the slow warp has observed this one - the waiter would then be
  // lapped and the kernel would dead-lock (seen as a rare launch failure from the bounded wait).
  uint64_t* g_full = bars + 14;
  uint64_t* out_ready = bars + 12; // [2] E4 wrote the output tile into the z buffer (8 warp arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous item ranges per CTA: consecutive items share the channel (and its spectrum lines in L2)
  const int per = (p.n_items + gridDim.x - 1) / gridDim.x;
  const int item0 = blockIdx.x * per, item1 = min(p.n_items, item0 + per);
  // work units = (item, chunk); all the pipeline state (buffers, barrier phases) is indexed by the unit counter `it`
  const int NC = CH ? p.n_chunks : 1;
  const int n_units = max(0, item1 - item0) * NC;
  constexpr int ZIM = CH ? 128 : ZIM_COL;   // step 7: z'_im column offset inside Y
  long long* trace = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
  int trace_n = 0;
  auto stamp = [&](int role) {
    if (trace && (threadIdx.x & 31) == 0 && trace_n < 64) trace[role * 64 + trace_n++] = clock64();
  };

  // constant stack -> shared memory
  {
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_S);
    for (int i = threadIdx.x; i < S_BYTES / 16; i += (CH ? THREADS_CH : THREADS)) dst[i] = __ldg(p.S + i);
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmVX); ptx::prefetch_tmap(&tmOut); ptx::prefetch_tmap(&tmX0);
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&z_full[i], 1); ptx::mbar_init(&z_empty[i], 1); ptx::mbar_init(&out_ready[i], 8); }
    ptx::mbar_init(x_full, 1); ptx::mbar_init(&p1_full[0], 8); ptx::mbar_init(&p1_full[1], 8);
    ptx::mbar_init(y_full, 1); ptx::mbar_init(&p2_full[0], 8); ptx::mbar_init(&p2_full[1], 8);
    ptx::mbar_init(x2_full, 1); ptx::mbar_init(bt_full, 8);
    ptx::mbar_init(o_full, 1); ptx::mbar_init(&g_full[0], 1); ptx::mbar_init(&g_full[1], 1);
    ptx::fence_mbar_init();
  } else if (warp == 1) {
    ptx::tmem_alloc<512>(tmem_ptr);
  }
  ptx::fence_proxy_async_smem();   // S was written with generic stores, UMMA reads it through the async proxy
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t TM_X = tmem_base, TM_Y = tmem_base + 256;

  constexpr int EPI_W0 = CH ? 4 : 2;   // first epilogue warp
  if (warp == 0) {
    // =========================== TMA producer (+ output stores) ===========================
    if constexpr (CH) ptx::setmaxnreg_dec<104>();
    // Buffer b cycles: z tile of item i -> (step 1 done) x0 gate tile of item i -> (E4) output tile of item i -> stored
    // -> z tile of item i + 2.  This thread issues the z loads and the output stores; the gate load is issued by an
    // epilogue thread, which is the one that knows when step 1 has finished.
    if (lane == 0) {
      for (int it = 0; it < n_units + 2; ++it) {
        const uint32_t buf = it & 1;
        uint8_t* z = smem + OFF_Z + buf * Z_BYTES;
        if (it >= 2) {   // retire unit it - 2: its output tile is complete in z
          const int un = it - 2, item = item0 + un / NC, c = un % NC;
          const int ch = item / p.n_pairs, pr = item % p.n_pairs;
          ptx::mbar_wait(&out_ready[buf], (un >> 1) & 1);
          ptx::tma_store_3d(&tmOut, z, 0, 64 * c, 2 * pr * p.D + ch);
          if (2 * pr + 1 < p.B) ptx::tma_store_3d(&tmOut, z + 16384, 0, 64 * c, (2 * pr + 1) * p.D + ch);
          ptx::tma_store_commit();
          ptx::tma_store_wait_read<0>();
        }
        if (it < n_units) {
          const int item = item0 + it / NC, c = it % NC;
          const int ch = item / p.n_pairs, pr = item % p.n_pairs;
          if (it >= 2) ptx::mbar_wait(&z_empty[buf], ((it - 2) >> 1) & 1);   // long since true
          ptx::mbar_expect_tx(&z_full[buf], Z_BYTES);
          for (int part = 0; part < 2; ++part) {
            const int row = (2 * pr + part) * p.D + ch;    // reads past B are out of bounds -> zero filled
            for (int a = 0; a < 2; ++a)
              ptx::tma_load_3d(z + part * 16384 + a * 8192, &tmVX, &z_full[buf], 64 * a, 64 * c, row);
          }
        }
      }
      ptx::tma_store_wait<0>();
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if constexpr (CH) ptx::setmaxnreg_dec<104>();
    {   // whole warp, uniform control flow; one elected lane issues (ptx::umma_f16_e)
      constexpr uint32_t id1 = idesc(128, false, true);    // step 1: A K-major, B MN-major, N = 128
      constexpr uint32_t id35 = idesc(256, false, false);  // steps 3, 5: A from TMEM, B K-major, N = 256
      // step 7: A MN-major, B K-major; N = 64 output rows n1, or 80 when tail tokens are wanted (row n1 = 64 = outputs
      // 8192..8319 of the same transform)
      const uint32_t id7 = CH ? idesc(128, true, false) : (p.nt > 0 ? idesc(80, true, false) : idesc(64, true, false));
      const uint32_t sS = ptx::smem_u32(smem + OFF_S), sBT = ptx::smem_u32(smem + OFF_BT);
      uint64_t dS = ptx::smem_desc_k_sw128(sS);
      auto s_desc = [&](int row0, int kk) -> uint64_t {    // constant rows row0.., K index kk (multiple of 16)
        return dS + (uint64_t)(((kk >> 6) * S_PANEL + row0 * 128 + (kk & 63) * 2) >> 4);
      };
      for (uint32_t it = 0; it < (uint32_t)n_units; ++it) {
        const uint32_t buf = it & 1, zph = (it >> 1) & 1, ph = it & 1;
        const uint32_t sZ = ptx::smem_u32(smem + OFF_Z + buf * Z_BYTES);
        asm volatile("" : "+l"(dS));   // keeps ptxas from tabulating every descriptor in local memory across items
        // ---- step 1: X[0,128) = A_re = Fre Zre - Fim Zim, X[128,256) = A_im = Fim Zre + Fre Zim
        stamp(0);
        ptx::mbar_wait(&z_full[buf], zph);
        ptx::tc_fence_after_sync();
        stamp(0);
        {
          const uint64_t dZ = ptx::smem_desc_mn_sw128(sZ, 8192, 1024);
#pragma unroll
          for (int op = 0; op < 4; ++op) {                 // (output half o, input part)
            const int o = op >> 1, part = op & 1;
            const int rows = o == 0 ? (part == 0 ? ROW_FRE : ROW_NFIM) : (part == 0 ? ROW_FIM : ROW_FRE);
#pragma unroll
            for (int j = 0; j < 4; ++j)                    // 16 n1 rows per MMA
              ptx::umma_f16_e(TM_X + o * 128, s_desc(rows, 16 * j), dZ + (uint64_t)((part * 16384 + 2048 * j) >> 4), id1,
                            (part | j) != 0);
          }
        }
        ptx::umma_commit_e(&z_empty[buf]);
        ptx::umma_commit_e(x_full);
        stamp(0);
        // ---- step 3: Y = [S_re | S_im] = P1 x F
        ptx::mbar_wait(&p1_full[0], ph);
        ptx::tc_fence_after_sync();
        stamp(0);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          if (s == 8) {   // second index half
            ptx::mbar_wait(&p1_full[1], ph);
            ptx::tc_fence_after_sync();
          }
          const int t = s & 1, n2 = 16 * (s >> 1);          // packed K order: per run of 16 indices, re then im
          const int rows = t == 0 ? ROW_FRE : ROW_NFIM;    // re part: [Fre; Fim], im part: [-Fim; Fre]
          ptx::umma_f16_ts_e(TM_Y, TM_X + 8 * s, s_desc(rows, n2), id35, s != 0);
        }
        ptx::umma_commit_e(y_full);
        stamp(0);
        // ---- step 5: X = [B_im | B_re] = P2 x conj(F)
        ptx::mbar_wait(&p2_full[0], ph);
        ptx::tc_fence_after_sync();
        stamp(0);
#pragma unroll
        for (int s = 0; s < 16; ++s) {
          if (s == 8) {
            ptx::mbar_wait(&p2_full[1], ph);
            ptx::tc_fence_after_sync();
          }
          const int t = s & 1, k2 = 16 * (s >> 1);
          const int rows = t == 0 ? ROW_NFIM : ROW_FRE;    // re part: [-Fim; Fre], im part: [Fre; Fim]
          ptx::umma_f16_ts_e(TM_X, TM_Y + 8 * s, s_desc(rows, k2), id35, s != 0);
        }
        ptx::umma_commit_e(x2_full);
        stamp(0);
        // ---- step 7: Y[0,64) = z_re = Bre Fre + Bim Fim, Y[64,128) = z_im = Bim Fre - Bre Fim   (n1 < 64)
        ptx::mbar_wait(bt_full, ph);
        ptx::tc_fence_after_sync();
        stamp(0);
        {
          const uint64_t dBT = ptx::smem_desc_mn_sw128(sBT, BT_ATOM, 1024);
#pragma unroll
          for (int op = 0; op < 4; ++op) {                 // (output half o, part: 0 = K rows B_re, 1 = B_im)
            const int o = op >> 1, part = op & 1;
            const int rows = o == 0 ? (part == 0 ? ROW_FRE : ROW_FIM) : (part == 0 ? ROW_NFIM : ROW_FRE);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              ptx::umma_f16_e(TM_Y + o * ZIM, dBT + (uint64_t)(((part * 128 + 16 * j) * 128) >> 4), s_desc(rows, 16 * j), id7,
                            (part | j) != 0);
          }
        }
        ptx::umma_commit_e(o_full);
        stamp(0);
      }
    }
  } else if (warp < EPI_W0) {
    if constexpr (CH) ptx::setmaxnreg_dec<104>();   // idle warps of the first warpgroup (chunked kernel only)
  } else {
    // =========================== epilogue warps ===========================
    if constexpr (CH) ptx::setmaxnreg_inc<200>();
    const int q = warp & 3, hf = (warp - EPI_W0) >> 2;
    const int r = q * 32 + lane;                         // TMEM lane: k1 (E1-E3) or n2 (E4)
    const uint32_t lane_addr = uint32_t(q * 32) << 16;
    const uint32_t sBT = ptx::smem_u32(smem + OFF_BT);
    // twiddles of this thread's row: step w = exp(-2 pi i r / N) and one seed per run of 16 indices
    float2 wstep, seed[4];
    sincospif(-2.0f * float(r) / float(N), &wstep.y, &wstep.x);
#pragma unroll
    for (int u = 0; u < 4; ++u) sincospif(-2.0f * float((r * (64 * hf + 16 * u)) % N) / float(N), &seed[u].y, &seed[u].x);
    // E1 / E2 walk the index (n2, k2) so that EVERY warp finishes the first half [0, 64) before the second: run u covers
    // 16 indices from col12(u); E3 keeps the plain split (its output rows go to the shared-memory atom hf)
    auto col12 = [&](int u) { return (u < 2 ? 0 : 64) + 32 * hf + 16 * (u & 1); };
    float2 seed1[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) sincospif(-2.0f * float((r * col12(u)) % N) / float(N), &seed1[u].y, &seed1[u].x);
    const bool tr = trace && warp == EPI_W0 && lane == 0;
    const int nt = p.nt;      // tail tokens after the last chunk, 0..LONGCONV_TAIL_MAX
    const float2 w2 = make_float2(wstep.x * wstep.x - wstep.y * wstep.y, 2.0f * wstep.x * wstep.y);
    // chunked mode: this CTA's scratch = parked spectra of chunks 0..NC-2 (float4 = two complex values; a warp's 32 lanes
    // write 512 contiguous bytes) followed by the carry (second halves of the last inverse transform, [2][64][128] fp32)
    uint4* park = reinterpret_cast<uint4*>(p.scratch + (long long)blockIdx.x * p.scratch_per_cta);
    float* carry = p.scratch + (long long)blockIdx.x * p.scratch_per_cta + (long long)(NC - 1) * N;
    const int e_warp = warp - EPI_W0;
    for (uint32_t it = 0; it < (uint32_t)n_units; ++it) {
      const uint32_t ph = it & 1, buf = it & 1;
      const int item = item0 + (int)it / NC, c = (int)it % NC;   // c: chunk (tokens [c C, c C + C))
      const int ch = item / p.n_pairs, pr = item % p.n_pairs;
      uint8_t* zb = smem + OFF_Z + buf * Z_BYTES;
      const int b0 = 2 * pr, b1 = 2 * pr + 1;
      const bool has1 = b1 < p.B;
      const long long row0 = ((long long)b0 * p.D + ch) * p.Tp, row1 = ((long long)b1 * p.D + ch) * p.Tp;
      const float osc = __ldg(p.osc + ch);   // output scale of this channel (exact power of two)
      float chk = 0.f;                       // becomes NaN when any output of this unit is inf / NaN
      if (tr) stamp(1);
      // Twiddle seeds are re-materialised per item: without the barrier ptxas precomputes all 256 twiddle values of the
      // thread once and keeps them in LOCAL memory (an L2 round trip per use with this shared-memory carve-out).
      float2 ws = wstep, sd[4];
      float w2x = w2.x, w2y = w2.y;
      asm volatile("" : "+f"(ws.x), "+f"(ws.y), "+f"(w2x), "+f"(w2y));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        sd[u] = seed[u];
        asm volatile("" : "+f"(sd[u].x), "+f"(sd[u].y));
      }
      const f2t W2X = f2_pack(w2x, w2x), W2Y = f2_pack(w2y, w2y), NW2Y = f2_pack(-w2y, -w2y);
      // ------------------------------------------------ E1: P1 = fp16(S1 tw .* A), 16 indices per pass, in place
      ptx::mbar_wait(x_full, ph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(1);
      if (threadIdx.x == EPI_W0 * 32) {   // z has been consumed: its buffer now receives the x0 gate tile [n1][n2] of both reads
        ptx::mbar_expect_tx(&g_full[buf], has1 ? Z_BYTES : Z_BYTES / 2);
        ptx::tma_load_3d(zb, &tmX0, &g_full[buf], 0, 64 * c, b0 * p.D + ch);
        if (has1) ptx::tma_load_3d(zb + 16384, &tmX0, &g_full[buf], 0, 64 * c, b1 * p.D + ch);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t xr[16], xi[16], w[16];
        const uint32_t col = col12(u);
        tmem_ld16(TM_X + lane_addr + col, xr);
        tmem_ld16(TM_X + lane_addr + 128 + col, xi);
        ptx::tmem_ld_wait();
        float2 s1v = seed1[u];
        asm volatile("" : "+f"(s1v.x), "+f"(s1v.y));
        const float2 t0 = make_float2(s1v.x * S1, s1v.y * S1);
        const float2 t1 = make_float2(t0.x * ws.x - t0.y * ws.y, t0.x * ws.y + t0.y * ws.x);
        f2t TWX = f2_pack(t0.x, t1.x), TWY = f2_pack(t0.y, t1.y);   // twiddles of elements (2 j, 2 j + 1)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const f2t XR = f2_packu(xr[2 * j], xr[2 * j + 1]), XI = f2_packu(xi[2 * j], xi[2 * j + 1]);
          w[j] = f2_to_h2(f2_sub(f2_mul(XR, TWX), f2_mul(XI, TWY)));
          w[8 + j] = f2_to_h2(f2_fma(XR, TWY, f2_mul(XI, TWX)));
          if (j < 7) {   // advance both twiddles by w^2
            const f2t NX = f2_fma(TWY, NW2Y, f2_mul(TWX, W2X));
            TWY = f2_fma(TWY, W2X, f2_mul(TWX, W2Y));
            TWX = NX;
          }
        }
        tmem_st8(TM_X + lane_addr + col, w);          // K slice 2 (col / 16) (re)
        tmem_st8(TM_X + lane_addr + col + 8, w + 8);  // K slice 2 (col / 16) + 1 (im)
        if (u & 1) {   // an index half is complete
          ptx::tmem_st_wait();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&p1_full[u >> 1]);
        }
      }
      if (tr) stamp(1);
      // ------------------------------------------------ E2: P2 = fp16(sum_j S_{c-j} .* G'_j)   (j = 0 only when not chunked)
      // spectrum table: uint4 (4 consecutive k2) at (((ch * 4 + k1 / 32) * 2 + k2 / 64) * 16 + (k2 % 64) / 4) * 32 + k1 % 32
      const uint4* gp = p.G + (((size_t)ch * 4 + q) * 2) * 16 * 32 + lane;
      auto g_off = [&](int u) { const int k2 = col12(u); return (size_t)((k2 >> 6) * 16 + ((k2 & 63) >> 2)) * 32; };
      uint4 g[16];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) g[4 * u + v] = __ldg(gp + g_off(u) + v * 32);
      ptx::mbar_wait(y_full, ph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t xr[16], xi[16], w[16];
        const uint32_t col = col12(u);
        tmem_ld16(TM_Y + lane_addr + col, xr);
        tmem_ld16(TM_Y + lane_addr + 128 + col, xi);
        ptx::tmem_ld_wait();
        f2t ar[8], ai[8];   // products for elements (2 m, 2 m + 1)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const uint32_t gw[4] = {g[4 * u + v].x, g[4 * u + v].y, g[4 * u + v].z, g[4 * u + v].w};
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
          // park this chunk's spectrum for the later chunks, then add the earlier chunks' spectra times the later filter
          // segments: W_c = sum_j S_{c-j} G_j (overlap-add in the frequency domain, one inverse transform per chunk)
          // (parked as fp16 in the layout of the spectrum table - the sum is rounded to fp16 right below anyway - so a run
          // is 4 x 16 B per thread; chunk region = 4096 uint4)
          const size_t slot = ((size_t)u * 4 * 8 + e_warp) * 32 + lane;   // + v * 256 per uint4, + chunk * 4096
          if (c < NC - 1) {
#pragma unroll
            for (int v = 0; v < 4; ++v)
              park[(size_t)c * 4096 + slot + (size_t)v * 256] =
                  make_uint4(pack_f16(__uint_as_float(xr[4 * v]), __uint_as_float(xr[4 * v + 1])),
                             pack_f16(__uint_as_float(xi[4 * v]), __uint_as_float(xi[4 * v + 1])),
                             pack_f16(__uint_as_float(xr[4 * v + 2]), __uint_as_float(xr[4 * v + 3])),
                             pack_f16(__uint_as_float(xi[4 * v + 2]), __uint_as_float(xi[4 * v + 3])));
          }
#pragma unroll 1
          for (int j = 1; j <= c; ++j) {
            const float relj = __ldg(p.rel + j * p.D + ch);   // 2^(e_0 - e_j): segment j's table -> segment 0's scale
            const f2t REL = f2_pack(relj, relj);
            const uint4* gj = gp + (size_t)j * p.g_seg_stride + g_off(u);
            uint4 gq[4], sq[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              gq[v] = __ldg(gj + v * 32);
              sq[v] = park[(size_t)(c - j) * 4096 + slot + (size_t)v * 256];
            }
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
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          w[m] = f2_to_h2(ar[m]);
          w[8 + m] = f2_to_h2(ai[m]);
        }
        tmem_st8(TM_Y + lane_addr + col, w);
        tmem_st8(TM_Y + lane_addr + col + 8, w + 8);
        if (u & 1) {
          ptx::tmem_st_wait();
          ptx::tc_fence_before_sync();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&p2_full[u >> 1]);
        }
      }
      if (tr) stamp(1);
      // ------------------------------------------------ E3: BT = fp16(conj(tw) .* B), shared memory
      ptx::mbar_wait(x2_full, ph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(1);
      asm volatile("" : "+f"(ws.x), "+f"(ws.y));
      {
        // row r (B_re) and row 128 + r (B_im) of atom hf; run u covers 16-byte chunks 2 u, 2 u + 1
        const uint32_t base_re = sBT + hf * BT_ATOM + r * 128, base_im = base_re + 128 * 128;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint32_t xi[16], xr[16], wr[8], wi[8];
          const uint32_t col = 64 * hf + 16 * u;
          tmem_ld16(TM_X + lane_addr + col, xi);          // B_im
          tmem_ld16(TM_X + lane_addr + 128 + col, xr);    // B_re
          ptx::tmem_ld_wait();
          float2 t0 = sd[u];
          asm volatile("" : "+f"(t0.x), "+f"(t0.y));
          const float2 t1 = make_float2(t0.x * ws.x - t0.y * ws.y, t0.x * ws.y + t0.y * ws.x);
          f2t TWX = f2_pack(t0.x, t1.x), TWY = f2_pack(t0.y, t1.y);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const f2t BR = f2_packu(xr[2 * j], xr[2 * j + 1]), BI = f2_packu(xi[2 * j], xi[2 * j + 1]);
            wr[j] = f2_to_h2(f2_fma(BI, TWY, f2_mul(BR, TWX)));      // (br + i bi)(tw.x - i tw.y)
            wi[j] = f2_to_h2(f2_sub(f2_mul(BI, TWX), f2_mul(BR, TWY)));
            if (j < 7) {
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
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bt_full);
      if (tr) stamp(1);
      // ------------------------------------------------ E4: out = z' * x0, staged in the (free) z buffer, TMA store
      // Tail tokens t = NC C + j (j < nt, last chunk only): the last transform's outputs C + j (row n1 = 64, lane j) hold
      // every (input chunk a, filter segment s) pair with a + s = NC - 1; what is missing are the pairs with a + s = NC,
      // i.e. per input chunk a (including the tail itself, a = NC) the <= j + 1 products x[a C + i] k'[(NC - a) C + j - i].
      const bool last_chunk = c == NC - 1;
      const bool tail_thread = nt > 0 && last_chunk && hf == 0 && q == 0 && lane < nt;
      float tc0 = 0.f, tc1 = 0.f, tx0 = 0.f, tx1 = 0.f;
      if (tail_thread) {
        const float* kq = p.k + (long long)ch * p.Lk;
        const int j = lane;
        for (int a = 0; a <= NC; ++a)
          for (int i = 0; i <= j; ++i) {
            const int tap = (NC - a) * C + j - i;
            const float kv = tap == 0 ? kq[0] + p.dbias[ch] : kq[tap];
            tc0 = fmaf(__half2float(p.vx[row0 + a * C + i]), kv, tc0);
            if (has1) tc1 = fmaf(__half2float(p.vx[row1 + a * C + i]), kv, tc1);
          }
        const float inva = __ldg(p.inva + ch);   // the vx rows hold a * v*x1
        tc0 *= inva;
        tc1 *= inva;
        tx0 = __bfloat162float(p.x0[row0 + NC * C + j]);
        if (has1) tx1 = __bfloat162float(p.x0[row1 + NC * C + j]);
      }
      // chunked mode: the previous chunk's overlap (this thread's 64 carry values, an L2 round trip each) is fetched into
      // registers BEFORE waiting for step 7, so the loads fly under that wait instead of inside the output loop
      // (first half here, second half while the first is being multiplied: all 64 at once do not fit the register file)
      float cra[CH ? 32 : 1], crb[CH ? 32 : 1];
      if constexpr (CH) {
        if (c > 0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            cra[i] = carry[(32 * hf + i) * 128 + r];
            crb[i] = carry[8192 + (32 * hf + i) * 128 + r];
          }
        }
      }
      ptx::mbar_wait(o_full, ph);
      ptx::tc_fence_after_sync();
      if (tr) stamp(1);
      ptx::mbar_wait(&g_full[buf], (it >> 1) & 1);
      unsigned short* st0 = reinterpret_cast<unsigned short*>(zb) + r;   // [n1][n2] bf16, 256 B per n1 row: gate in, product out
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t zr[16], zi[16];
        tmem_ld16(TM_Y + lane_addr + 32 * hf + 16 * h2, zr);
        tmem_ld16(TM_Y + lane_addr + ZIM + 32 * hf + 16 * h2, zi);
        if constexpr (CH) {
          if (c > 0 && h2 == 0) {
#pragma unroll
            for (int i = 16; i < 32; ++i) {
              cra[i] = carry[(32 * hf + i) * 128 + r];
              crb[i] = carry[8192 + (32 * hf + i) * 128 + r];
            }
          }
        }
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n1 = 32 * hf + 16 * h2 + j;
          float va = __uint_as_float(zr[j]), vb = __uint_as_float(zi[j]);
          if constexpr (CH) {   // + the previous chunk's overlap
            if (c > 0) {
              va += cra[16 * h2 + j];
              vb += crb[16 * h2 + j];
            }
          }
          va *= osc;
          vb *= osc;
          // One output per thread is enough: an overflow in P1 / P2 / the parked spectra reaches every output of the unit,
          // one in BT reaches every output of its row n2 - and a row is a thread here.
          if (j == 0) chk = fmaf(va, 0.f, fmaf(vb, 0.f, chk));
          const float ga = __uint_as_float(uint32_t(st0[n1 * 128]) << 16);
          const __nv_bfloat16 oa = __float2bfloat16(va * ga);
          st0[n1 * 128] = *reinterpret_cast<const unsigned short*>(&oa);
          if (has1) {
            const float gb = __uint_as_float(uint32_t(st0[8192 + n1 * 128]) << 16);
            const __nv_bfloat16 ob = __float2bfloat16(vb * gb);
            st0[8192 + n1 * 128] = *reinterpret_cast<const unsigned short*>(&ob);
          }
        }
      }
      if constexpr (CH) {
        // second half of this transform (rows n1 >= 64) = what it contributes to the next chunk's outputs.  The thread that
        // reads carry[n1][r] for n1 in [32 hf, 32 hf + 32) above is the one that rewrites exactly those entries here.
        if (!last_chunk) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t zr[16], zi[16];
            tmem_ld16(TM_Y + lane_addr + 64 + 32 * hf + 16 * h2, zr);
            tmem_ld16(TM_Y + lane_addr + ZIM + 64 + 32 * hf + 16 * h2, zi);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n1 = 32 * hf + 16 * h2 + j;
              carry[n1 * 128 + r] = __uint_as_float(zr[j]);
              carry[8192 + n1 * 128 + r] = __uint_as_float(zi[j]);
            }
          }
        }
      }
      if (nt > 0 && last_chunk && hf == 0 && q == 0) {   // warp-uniform
        uint32_t zr[16], zi[16];
        tmem_ld16(TM_Y + lane_addr + 64, zr);
        tmem_ld16(TM_Y + lane_addr + ZIM + 64, zi);
        ptx::tmem_ld_wait();
        if (tail_thread) {
          p.out[row0 + NC * C + lane] = __float2bfloat16((__uint_as_float(zr[0]) * osc + tc0) * tx0);
          if (has1) p.out[row1 + NC * C + lane] = __float2bfloat16((__uint_as_float(zi[0]) * osc + tc1) * tx1);
        }
      }
      if (chk != chk) atomicOr(p.err, 2);   // an fp16 operand overflowed somewhere in this unit: the launch is reported
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&out_ready[buf]);
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
