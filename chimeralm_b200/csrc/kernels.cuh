// Bandwidth-bound kernels of the predict path: token encoding, embedding gather, LayerNorm,
// depthwise short conv + first gate (with the token-major -> channel-major transpose the long
// convolution wants), the inverse transpose, attention pooling and the classifier head.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "ptx.cuh"

namespace clm {

// ---------------------------------------------------------------------------------------------
// Character tokenisation.  Reference: chimeralm/data/tokenizer.py:264-268 (char -> id, unknown
// -> [UNK]=6), :297-306 ([CLS] + ids + [SEP]; Hub flavour ids + [SEP]), truncation keeps the
// first bases, DataCollator pads with [PAD]=4 on `padding_side` (:152-159).
__constant__ uint8_t c_base_lut[256];

struct EncodeParams {
  const uint8_t* bases;     // concatenated ASCII reads
  const int64_t* offsets;   // [B+1]
  uint8_t* ids;             // [B, T_pad]
  int32_t* lens;            // [B] token count incl. specials (may be null)
  int B, T_pad, add_cls, add_sep, pad_left, max_bases;
};

__global__ void __launch_bounds__(256) encode_kernel(EncodeParams p) {
  const int b = blockIdx.y;
  const int64_t beg = p.offsets[b];
  int64_t nb = p.offsets[b + 1] - beg;
  if (nb > p.max_bases) nb = p.max_bases;
  const int ntok = (int)nb + p.add_cls + p.add_sep;
  const int start = p.pad_left ? (p.T_pad - ntok) : 0;
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.lens) p.lens[b] = ntok;
  uint8_t* row = p.ids + (int64_t)b * p.T_pad;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.T_pad; t += gridDim.x * blockDim.x) {
    const int j = t - start;  // index inside the token sequence
    uint8_t id = 4;           // [PAD]
    if (j >= 0 && j < ntok) {
      if (p.add_cls && j == 0) id = 0;
      else if (p.add_sep && j == ntok - 1) id = 1;
      else id = c_base_lut[p.bases[beg + j - p.add_cls]];
    }
    row[t] = id;
  }
}

// ---------------------------------------------------------------------------------------------
// Embedding gather (HyenaEmbeddings.word_embeddings, A.2): one warp per token, fp32 residual out
// in the R32 blocked layout (ptx::r32_off).
// It also writes the layer-0 LayerNorm input xn = (E[id] - mean) / std as bf16 token-major, looked up from a
// per-vocabulary-row table built at finalize (the block kernels fold the LayerNorm affine into their weights).
template <typename IdT>
__global__ void __launch_bounds__(256) embed_kernel(const IdT* __restrict__ ids, const float* __restrict__ E,
                                                    const __nv_bfloat16* __restrict__ En, float* __restrict__ R,
                                                    __nv_bfloat16* __restrict__ XN, long long M, int D, int vocab_rows,
                                                    int* __restrict__ err) {
  // One CTA per 32 consecutive rows = one contiguous 32 KB block of the R32 layout.  For the fp32 residual a warp's lanes
  // are the 32 ROWS and it walks column groups, so every store instruction writes 512 contiguous bytes (lane-per-column
  // wrote 16 bytes every 512).  The embedding table (<= 16 rows) sits in shared memory, rows padded by 16 B against bank
  // conflicts between lanes holding different ids.  The bf16 xn rows are token-major: one warp per row, 512 B per store.
  constexpr int DD = 256, PADF = DD + 4;
  __shared__ __align__(16) float Es[16 * PADF];
  __shared__ int s_id[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row0 = (long long)blockIdx.x * 32;
  const int nv = min(vocab_rows, 16);
  for (int i = threadIdx.x; i < nv * (DD / 4); i += 256) {
    const int v = i / (DD / 4), c4 = i % (DD / 4);
    *reinterpret_cast<float4*>(Es + v * PADF + 4 * c4) = __ldg(reinterpret_cast<const float4*>(E + (long long)v * D) + c4);
  }
  if (threadIdx.x < 32) {
    const long long row = row0 + threadIdx.x;
    long long id = row < M ? (long long)ids[row] : 0;
    if (id < 0 || id >= nv) {
      if (row < M) atomicOr(err, 1);
      id = 0;
    }
    s_id[threadIdx.x] = (int)id;
  }
  __syncthreads();
  {
    const long long row = row0 + lane;
    const float* src = Es + s_id[lane] * PADF;
    if (row < M) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c4 = warp + 8 * k;   // column group (4 floats)
        *reinterpret_cast<float4*>(R + ptx::r32_off(row, 4 * c4)) = *reinterpret_cast<const float4*>(src + 4 * c4);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int rr = warp * 4 + k;
    const long long row = row0 + rr;
    if (XN && row < M) reinterpret_cast<uint4*>(XN + row * D)[lane] = __ldg(reinterpret_cast<const uint4*>(En + (long long)s_id[rr] * D) + lane);
  }
}

// En[v,:] = bf16((E[v,:] - mean) * rsqrt(var + eps)), one warp per vocabulary row
__global__ void embed_norm_table_kernel(const float* __restrict__ E, __nv_bfloat16* __restrict__ En, int rows, float eps) {
  constexpr int D = 256;
  const int v = blockIdx.x, lane = threadIdx.x;
  if (v >= rows) return;
  float x[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = E[v * D + lane + 32 * i]; s += x[i]; }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / D;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] -= mean; ss += x[i] * x[i]; }
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float rstd = rsqrtf(ss / D + eps);
#pragma unroll
  for (int i = 0; i < 8; ++i) En[v * D + lane + 32 * i] = __float2bfloat16(x[i] * rstd);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over D=256 (eps inside rsqrt, biased variance, as torch.nn.LayerNorm).  Input: fp32
// residual in the R32 blocked layout; output: bf16 token-major (GEMM operand).  One block per
// 32-row group: the group's 32 KB are contiguous in R32 and are copied to shared memory with
// fully coalesced 16-byte loads, then each warp normalises 4 rows.
__global__ void __launch_bounds__(256) layernorm_bf16_kernel(const float* __restrict__ X, const float* __restrict__ g,
                                                             const float* __restrict__ bta,
                                                             __nv_bfloat16* __restrict__ Y, long long M, float eps) {
  constexpr int D = 256, CH = 64, CS = 33;  // 64 column chunks; chunk stride 33 float4 (pad 1) against bank conflicts
  __shared__ float4 sx[CH * CS];
  const long long row0 = (long long)blockIdx.x * 32;
  const float4* src = reinterpret_cast<const float4*>(X + row0 * D);  // r32_off(row0, 0) == row0 * 256
  for (int i = threadIdx.x; i < CH * 32; i += 256) sx[(i >> 5) * CS + (i & 31)] = src[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g) + lane), g1 = __ldg(reinterpret_cast<const float4*>(g) + lane + 32);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bta) + lane), b1 = __ldg(reinterpret_cast<const float4*>(bta) + lane + 32);
#pragma unroll
  for (int rr = 0; rr < 4; ++rr) {
    const int r = warp * 4 + rr;
    const long long row = row0 + r;
    if (row >= M) break;
    const float4 a = sx[lane * CS + r], b = sx[(lane + 32) * CS + r];
    float s = a.x + a.y + a.z + a.w + b.x + b.y + b.z + b.w;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / D);
    float v[8] = {a.x - mean, a.y - mean, a.z - mean, a.w - mean, b.x - mean, b.y - mean, b.z - mean, b.w - mean};
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += v[i] * v[i];
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float rstd = rsqrtf(ss * (1.0f / D) + eps);
    __nv_bfloat162 o0 = __floats2bfloat162_rn(v[0] * rstd * g0.x + b0.x, v[1] * rstd * g0.y + b0.y);
    __nv_bfloat162 o1 = __floats2bfloat162_rn(v[2] * rstd * g0.z + b0.z, v[3] * rstd * g0.w + b0.w);
    __nv_bfloat162 o2 = __floats2bfloat162_rn(v[4] * rstd * g1.x + b1.x, v[5] * rstd * g1.y + b1.y);
    __nv_bfloat162 o3 = __floats2bfloat162_rn(v[6] * rstd * g1.z + b1.z, v[7] * rstd * g1.w + b1.w);
    uint2* y2 = reinterpret_cast<uint2*>(Y + row * D);
    y2[lane] = make_uint2(*reinterpret_cast<uint32_t*>(&o0), *reinterpret_cast<uint32_t*>(&o1));
    y2[lane + 32] = make_uint2(*reinterpret_cast<uint32_t*>(&o2), *reinterpret_cast<uint32_t*>(&o3));
  }
}

// ---------------------------------------------------------------------------------------------
// Depthwise causal short conv (k=3, Conv1d padding=2 truncated to T) + first gate, A.3:
//   uc[ch,t] = w[ch,0] u[ch,t-2] + w[ch,1] u[ch,t-1] + w[ch,2] u[ch,t] + b[ch]   (u = 0 for t < 0)
//   x0 = uc[0:D], x1 = uc[D:2D], v = uc[2D:3D];   vx = v * x1
// Input U token-major bf16 [B*T, 3D]; outputs channel-major bf16 [B][D][Tp].
// Block = 64 tokens x 32 channels (x3 groups), transposed through shared memory.
__global__ void __launch_bounds__(256) shortconv_gate_kernel(const __nv_bfloat16* __restrict__ U,
                                                             const float* __restrict__ w,   // [3D,3]
                                                             const float* __restrict__ bias,  // [3D]
                                                             __nv_bfloat16* __restrict__ VX,
                                                             __nv_bfloat16* __restrict__ X0, int T, int Tp, int D) {
  constexpr int TT = 64, TC = 32;
  __shared__ float su[3][TT + 2][TC + 1];
  const int b = blockIdx.z, t0 = blockIdx.x * TT, c0 = blockIdx.y * TC;
  const int tid = threadIdx.x;
  const long long rowbase = (long long)b * T;
  // load: (TT+2) tokens x 3 groups x TC channels, 2 channels (one bf16x2) per thread-iteration
  for (int i = tid; i < (TT + 2) * 3 * (TC / 2); i += 256) {
    const int cp = i % (TC / 2);
    const int g = (i / (TC / 2)) % 3;
    const int tt = i / (3 * TC / 2);
    const int t = t0 - 2 + tt;
    float2 f = make_float2(0.f, 0.f);
    if (t >= 0 && t < T) {
      const __nv_bfloat162 v =
          *reinterpret_cast<const __nv_bfloat162*>(U + (rowbase + t) * (3LL * D) + g * D + c0 + 2 * cp);
      f = __bfloat1622float2(v);
    }
    su[g][tt][2 * cp] = f.x;
    su[g][tt][2 * cp + 1] = f.y;
  }
  __syncthreads();
  // compute + transposed store: thread handles channel cl, tokens (tl, tl+1)
  for (int i = tid; i < TC * (TT / 2); i += 256) {
    const int tl = (i % (TT / 2)) * 2;
    const int cl = i / (TT / 2);
    const int c = c0 + cl;
    float uc[3][2];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int ch = g * D + c;
      const float w0 = __ldg(w + ch * 3), w1 = __ldg(w + ch * 3 + 1), w2 = __ldg(w + ch * 3 + 2), bb = __ldg(bias + ch);
#pragma unroll
      for (int k = 0; k < 2; ++k)
        uc[g][k] = w0 * su[g][tl + k][cl] + w1 * su[g][tl + k + 1][cl] + w2 * su[g][tl + k + 2][cl] + bb;
    }
    const int t = t0 + tl;
    if (t < Tp) {
      const long long o = ((long long)b * D + c) * Tp + t;
      const float vx0 = (t < T) ? uc[2][0] * uc[1][0] : 0.f, vx1 = (t + 1 < T) ? uc[2][1] * uc[1][1] : 0.f;
      const float x00 = (t < T) ? uc[0][0] : 0.f, x01 = (t + 1 < T) ? uc[0][1] : 0.f;
      *reinterpret_cast<__nv_bfloat162*>(VX + o) = __floats2bfloat162_rn(vx0, vx1);
      *reinterpret_cast<__nv_bfloat162*>(X0 + o) = __floats2bfloat162_rn(x00, x01);
    }
  }
}

// Channel-major bf16 [B][D][Tp] -> token-major bf16 [B*T, D] (out_proj's A operand).
__global__ void __launch_bounds__(256) transpose_ct_kernel(const __nv_bfloat16* __restrict__ Y,
                                                           __nv_bfloat16* __restrict__ YT, int T, int Tp, int D) {
  constexpr int TT = 64, TC = 64;
  __shared__ __nv_bfloat16 s[TC][TT + 2];
  const int b = blockIdx.z, t0 = blockIdx.x * TT, c0 = blockIdx.y * TC;
  const int tid = threadIdx.x;
  for (int i = tid; i < TC * TT; i += 256) {
    const int tl = i % TT, cl = i / TT;
    const int t = t0 + tl;
    s[cl][tl] = (t < T) ? Y[((long long)b * D + c0 + cl) * Tp + t] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = tid; i < TT * TC; i += 256) {
    const int cl = i % TC, tl = i / TC;
    const int t = t0 + tl;
    if (t < T) YT[((long long)b * T + t) * D + c0 + cl] = s[cl][tl];
  }
}

// ---------------------------------------------------------------------------------------------
// Attention pooling (chimeralm/models/components/hyena.py:117-132, mask == None):
//   w = softmax_t(score[b,:]) over ALL T positions;  pooled[b,:] = sum_t w_t * ln_f(R[b,t,:])
// Split over the sequence: each block reduces a slice with a running max (online softmax) and
// writes (max, sum, acc[D]); the head kernel merges slices.  The rows are the bf16 normalised tokens; the
// weighted average over thousands of tokens washes out their 2^-9 rounding.
__global__ void __launch_bounds__(256) pool_partial_kernel(const __nv_bfloat16* __restrict__ XN, const float* __restrict__ score,
                                                           const float* __restrict__ g, const float* __restrict__ bta,
                                                           float* __restrict__ part,  // [B][S][2+D]
                                                           int T, int n_split, int pool_mode) {
  // XN = (r - mean) * rstd of the final residual (bf16, token-major, emitted by the last block or the LayerNorm
  // kernel with unit affine); ln_f's gamma/beta are applied ONCE to the pooled sum: sum_t w_t (xn_t g + b) =
  // g * sum_t w_t xn_t + b * sum_t w_t.  One warp per token row (512 contiguous bytes), 8 bf16 per lane.
  constexpr int D = 256;
  __shared__ float sm_m[8], sm_l[8];
  __shared__ float sm_acc[8][D];
  const int b = blockIdx.y, sp = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (T + n_split - 1) / n_split;
  const int tb = sp * per, te = min(T, tb + per);
  float m = -INFINITY, l = 0.f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 2
  for (int t = tb + warp; t < te; t += 8) {
    const long long row = (long long)b * T + t;
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(XN + row * D) + lane);
    float sc = __ldg(score + row);
    if (pool_mode == 1) sc = 0.f;                                  // mean: uniform weights
    else if (pool_mode == 3) {                                     // cls: position 0 only
      if (t != 0) continue;
      sc = 0.f;
    }
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    if (pool_mode == 2) {   // max: acc[] carries the running max, m/l are unused; negative gammas are handled by the caller's
                            // sign trick below (max of gamma * x = gamma * (gamma >= 0 ? max x : min x)) - here per element
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c0 = lane * 8 + 2 * i;
        const float g0 = __ldg(g + c0), g1 = __ldg(g + c0 + 1);
        const float h0 = g0 * __uint_as_float(w[i] << 16), h1 = g1 * __uint_as_float(w[i] & 0xffff0000u);
        acc[2 * i] = (l == 0.f) ? h0 : fmaxf(acc[2 * i], h0);
        acc[2 * i + 1] = (l == 0.f) ? h1 : fmaxf(acc[2 * i + 1], h1);
      }
      l = 1.f;
      m = 0.f;
      continue;
    }
    const float mn = fmaxf(m, sc);
    const float corr = __expf(m - mn);  // exp(-inf) = 0 on the first step
    const float pw = __expf(sc - mn);
    l = l * corr + pw;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] = acc[2 * i] * corr + pw * __uint_as_float(w[i] << 16);
      acc[2 * i + 1] = acc[2 * i + 1] * corr + pw * __uint_as_float(w[i] & 0xffff0000u);
    }
    m = mn;
  }
  if (lane == 0) { sm_m[warp] = m; sm_l[warp] = l; }
#pragma unroll
  for (int i = 0; i < 8; ++i) sm_acc[warp][lane * 8 + i] = acc[i];
  __syncthreads();
  float M = -INFINITY;
  for (int w = 0; w < 8; ++w) M = fmaxf(M, sm_m[w]);
  float* out = part + ((long long)b * n_split + sp) * (2 + D);
  const int d = threadIdx.x;  // 256 threads == D
  float a = 0.f, L = 0.f;
  for (int w = 0; w < 8; ++w) {
    const float f = (sm_m[w] == -INFINITY) ? 0.f : __expf(sm_m[w] - M);
    a += sm_acc[w][d] * f;
    L += sm_l[w] * f;
  }
  if (pool_mode == 2) {   // max over the warps' running maxima of gamma * xn; + beta
    float mx = -INFINITY;
    for (int w = 0; w < 8; ++w)
      if (sm_l[w] > 0.f) mx = fmaxf(mx, sm_acc[w][d]);
    out[2 + d] = mx + __ldg(bta + d);
    if (d == 0) { out[0] = 0.f; out[1] = mx == -INFINITY ? 0.f : 1.f; }
    return;
  }
  out[2 + d] = a * __ldg(g + d) + L * __ldg(bta + d);   // ln_f affine applied to the (unnormalised) weighted sum
  if (d == 0) { out[0] = M; out[1] = L; }
}

// Attention weights export (components/hyena.py:129-130, `save_attention=True`): w[b,t] = softmax_t(score[b,:]).
// One block per read; the scores are those of the last forward.
__global__ void __launch_bounds__(256) attention_softmax_kernel(const float* __restrict__ score, float* __restrict__ out, int T) {
  __shared__ float red[8];
  const float* s = score + (long long)blockIdx.x * T;
  float* o = out + (long long)blockIdx.x * T;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int t = threadIdx.x; t < T; t += 256) m = fmaxf(m, s[t]);
  for (int k = 16; k; k >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, k));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float l = 0.f;
  for (int t = threadIdx.x; t < T; t += 256) l += expf(s[t] - m);
  for (int k = 16; k; k >>= 1) l += __shfl_xor_sync(0xffffffffu, l, k);
  if (lane == 0) red[warp] = l;
  __syncthreads();
  l = 0.f;
  for (int w = 0; w < 8; ++w) l += red[w];
  const float inv = 1.0f / l;
  for (int t = threadIdx.x; t < T; t += 256) o[t] = expf(s[t] - m) * inv;
}

// Classifier head (components/hyena.py:55-74,142-146,149-180): merge pooling slices, then
//   Lin(256,512) GELU Lin(512,512) GELU [Lin(512,512) GELU Lin(512,512)] + skip, Lin(512,2);
//   label = argmax(logits) with ties -> 0 (chimeralm/models/callbacks.py:107).
struct HeadParams {
  const float* part; int n_split;
  const float *w0, *b0, *w1, *b1, *wr0, *br0, *wr1, *br1, *wo, *bo;
  float* logits;      // [B,2]
  uint8_t* labels;    // [B] (may be null)
  float* pooled_out;  // [B,256] (may be null; debug / attention export)
};

__device__ __forceinline__ float gelu_erf_h(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }

// ---------------------------------------------------------------------------------------------
// Head as a chain of small kernels in which every weight row is read ONCE for the whole batch (a one-CTA-per-read
// head streamed all 4 MB of head weights per read; at batch 32 that is 128 MB through L2).  Unfused path; the fused
// head_fused_kernel below is what clm_forward launches.
__global__ void __launch_bounds__(256) pool_merge_kernel(const float* __restrict__ part, int n_split, float* __restrict__ pooled,
                                                         int pool_mode) {
  constexpr int D = 256;
  const int b = blockIdx.x, d = threadIdx.x;
  const float* pp = part + (long long)b * n_split * (2 + D);
  if (pool_mode == 2) {   // max pooling: the partials hold per-slice maxima
    float mx = -INFINITY;
    for (int s = 0; s < n_split; ++s)
      if (pp[s * (2 + D) + 1] > 0.f) mx = fmaxf(mx, pp[s * (2 + D) + 2 + d]);
    pooled[(long long)b * D + d] = mx;
    return;
  }
  float M = -INFINITY;
  for (int s = 0; s < n_split; ++s) M = fmaxf(M, pp[s * (2 + D)]);
  float a = 0.f, L = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float ms = pp[s * (2 + D)];
    const float f = (ms == -INFINITY) ? 0.f : __expf(ms - M);
    a += pp[s * (2 + D) + 2 + d] * f;
    L += pp[s * (2 + D) + 1] * f;
  }
  pooled[(long long)b * D + d] = a / L;
}

// y[b,o] = act(bias[o] + sum_i W[o,i] x[b,i]) (+ skip[b,o]); one warp per output neuron o, looping over the batch.
// FINAL: OUT == 2 -> also writes labels (argmax, ties -> 0; chimeralm/models/callbacks.py:107).
template <int IN, bool GELU, bool SKIP, bool FINAL>
__global__ void __launch_bounds__(256) head_layer_kernel(const float* __restrict__ W, const float* __restrict__ bias,
                                                         const float* __restrict__ x, const float* __restrict__ skip,
                                                         float* __restrict__ y, uint8_t* __restrict__ labels, int B, int OUT) {
  const int o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= OUT) return;
  constexpr int V = IN / 128;   // float4 per lane
  float4 w[V];
#pragma unroll
  for (int v = 0; v < V; ++v) w[v] = __ldg(reinterpret_cast<const float4*>(W + (long long)o * IN) + lane + 32 * v);
  const float bo = __ldg(bias + o);
  const int b_lo = blockIdx.y * 32, b_hi = min(B, b_lo + 32);   // grid.y tiles the batch
#pragma unroll 4
  for (int b = b_lo; b < b_hi; ++b) {
    float a = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float4 xv = __ldg(reinterpret_cast<const float4*>(x + (long long)b * IN) + lane + 32 * v);
      a += w[v].x * xv.x + w[v].y * xv.y + w[v].z * xv.z + w[v].w * xv.w;
    }
    for (int s = 16; s > 0; s >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s);
    if (lane == 0) {
      a += bo;
      if (GELU) a = gelu_erf_h(a);
      if (SKIP) a += skip[(long long)b * OUT + o];
      y[(long long)b * OUT + o] = a;
    }
  }
  if (FINAL) {   // OUT == 2: both logits of a read are written by warps 0 and 1 of block 0
    __syncthreads();
    if (labels && o == 0)
      for (int b = b_lo + lane; b < b_hi; b += 32) labels[b] = (y[b * 2 + 1] > y[b * 2]) ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// The whole head in ONE cooperative launch (pooling merge + 5 layers were 6 launch-latency-bound kernels, 0.11 ms/step):
// 64 CTAs, each weight row still read once for the whole batch, a grid-wide barrier between layers.  The barrier is a
// monotonically increasing counter (never reset; the host passes the value it has before this launch), the launch is
// cooperative so that all CTAs are guaranteed to be resident.  Activations written by other CTAs are read with plain
// (coherent) loads after the barrier's fences, never through the read-only path.
struct HeadFusedParams {
  const float* part; int n_split;
  const float *w0, *b0, *w1, *b1, *wr0, *br0, *wr1, *br1, *wo, *bo;
  float *pooled, *h0, *h1, *h2, *h3;   // [B,256], [B,512] x4
  float* logits;      // [B,2]
  uint8_t* labels;    // [B] (may be null)
  unsigned int* counter;
  unsigned int base;  // counter value before this launch
  int B;
  int pool_mode;      // 2 = max pooling (the partials are maxima); otherwise softmax-weighted sums
  int* err;           // status word of the forward in flight (may be null)
  int* status_out;    // mapped host slot that receives (status_tag | *err) when the forward is complete (may be null)
  int status_tag;     // forward sequence number << 8
};

// Last thing a forward does: hand the status word to the host (mapped pinned memory) and clear it for the next forward.
__device__ __forceinline__ void publish_status(int* err, int* status_out, int tag) {
  if (!err) return;
  const int v = atomicExch(err, 0);
  if (status_out) {
    *reinterpret_cast<volatile int*>(status_out) = tag | (v & 0xff);
    __threadfence_system();
  }
}

__global__ void publish_status_kernel(int* err, int* status_out, int tag) { publish_status(err, status_out, tag); }

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    // bounded like the mbarrier waits: a host/device counter mismatch is a reported launch failure, never a hung GPU
    unsigned int polls = 0;
    while ((int)(*reinterpret_cast<volatile unsigned int*>(counter) - target) < 0)
      if (++polls > (1u << 28)) __trap();
    __threadfence();
  }
  __syncthreads();
}

constexpr int HEAD_BT = 32;   // reads per shared-memory tile of the layer input (32 x 512 fp32 = 64 KB)

// One layer for output neuron o (one warp), all reads: the layer input is staged in shared memory tile by tile (every warp of
// the CTA needs all of it; from global memory each warp would pull B x IN floats through L2 again).
// weight row and bias of output neuron o: loaded BEFORE the grid barrier in front of the layer (they do not depend on it), so
// the L2 round trip overlaps the barrier instead of following it
template <int IN>
__device__ __forceinline__ void head_load_w(const float* __restrict__ W, const float* __restrict__ bias, int o, int OUT,
                                            float4 (&w)[4], float& bo) {
  const int lane = threadIdx.x & 31;
  bo = 0.f;
  if (o < OUT) {
#pragma unroll
    for (int v = 0; v < IN / 128; ++v) w[v] = __ldg(reinterpret_cast<const float4*>(W + (long long)o * IN) + lane + 32 * v);
    bo = __ldg(bias + o);
  }
}

template <int IN, bool GELU, bool SKIP>
__device__ __forceinline__ uint32_t head_fused_layer(const float4 (&w)[4], float bo, const float* x,
                                                 const float* skip, float* y, int o, int OUT, int b_lo, int B, float* xs,
                                                 uint64_t* bar, uint32_t phase) {
  const int lane = threadIdx.x & 31;
  constexpr int V = IN / 128;
  for (int b0 = b_lo; b0 < B; b0 += HEAD_BT) {   // this CTA's reads [b_lo, B)
    const int nb = min(HEAD_BT, B - b0);
    __syncthreads();   // previous tile fully consumed
    // One bulk copy (<= 64 KB) by the TMA engine instead of 64 dependent 16-byte loads per thread: the staging was ~5 us of
    // each layer's ~15.  The rows were written by other CTAs before the grid barrier (generic proxy, fenced there); the
    // proxy fence orders this thread's acquire of that barrier before the async-proxy read.
    if (threadIdx.x == 0) {
      ptx::fence_proxy_async_all();
      ptx::mbar_expect_tx(bar, (uint32_t)(nb * IN * 4));
      ptx::bulk_load_1d(xs, x + (long long)b0 * IN, (uint32_t)(nb * IN * 4), bar);
    }
    ptx::mbar_wait(bar, phase & 1);
    ++phase;
    if (o < OUT) {
      static_assert(HEAD_BT == 32, "one read of the tile per lane in the layer epilogue");
      float mine = 0.f;   // lane b keeps read b's dot product: bias, GELU, skip and the store then run once, 32 reads wide
#pragma unroll 4
      for (int b = 0; b < nb; ++b) {
        float a = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
          const float4 xv = reinterpret_cast<const float4*>(xs + b * IN)[lane + 32 * v];
          a += w[v].x * xv.x + w[v].y * xv.y + w[v].z * xv.z + w[v].w * xv.w;
        }
        for (int s2 = 16; s2 > 0; s2 >>= 1) a += __shfl_xor_sync(0xffffffffu, a, s2);
        if (b == lane) mine = a;
      }
      if (lane < nb) {
        float a = mine + bo;
        if (GELU) a = gelu_erf_h(a);
        if (SKIP) a += skip[(long long)(b0 + lane) * OUT + o];
        y[(long long)(b0 + lane) * OUT + o] = a;
      }
    }
  }
  return phase;
}

__global__ void __launch_bounds__(256) head_fused_kernel(HeadFusedParams p) {
  constexpr int D = 256, H = 512;
  extern __shared__ __align__(16) float head_xs[];   // HEAD_BT x 512 fp32
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // grid = (H / 8 neuron groups, read slices): every weight row is still read once per slice (<= 4 slices), and a batch of
  // 255 short reads no longer walks eight 32-read tiles through one CTA per neuron group (0.41 ms at B = 255,
  // profiles/r2_length_sweep.txt)
  const unsigned int G = gridDim.x * gridDim.y;
  const unsigned int cta = blockIdx.y * gridDim.x + blockIdx.x;
  const int per = ((p.B + (int)gridDim.y - 1) / (int)gridDim.y + HEAD_BT - 1) / HEAD_BT * HEAD_BT;   // reads per slice, whole tiles
  const int b_lo = (int)blockIdx.y * per, b_hi = min(p.B, b_lo + per);
  __shared__ uint64_t head_bar;
  uint32_t phase = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&head_bar, 1);
    ptx::fence_mbar_init();
  }
  // phase 0: merge the pooling partials of each read (pool_merge_kernel).  The per-slice maxima and sums are staged in
  // shared memory first, so that the long loop over slices carries only independent loads (the sums are accumulated in the
  // same order as before: bit-identical results)
  for (int b = cta; b < p.B; b += G) {
    const int d = threadIdx.x;
    const float* pp = p.part + (long long)b * p.n_split * (2 + D);
    if (p.pool_mode == 2) {   // max pooling: the partials hold per-tile maxima
      float mx = -INFINITY;
#pragma unroll 8
      for (int s = 0; s < p.n_split; ++s) {
        const float cnt = pp[s * (2 + D) + 1], v = pp[s * (2 + D) + 2 + d];
        if (cnt > 0.f) mx = fmaxf(mx, v);
      }
      p.pooled[(long long)b * D + d] = mx;
      continue;
    }
    float* s_m = head_xs;                 // [n_split] slice maxima, then the factors exp(m_s - M)
    float* s_l = head_xs + p.n_split;     // [n_split] slice sums
    __syncthreads();
    for (int s = threadIdx.x; s < p.n_split; s += blockDim.x) {
      s_m[s] = pp[s * (2 + D)];
      s_l[s] = pp[s * (2 + D) + 1];
    }
    __syncthreads();
    float M = -INFINITY;
    for (int s = 0; s < p.n_split; ++s) M = fmaxf(M, s_m[s]);
    __syncthreads();
    for (int s = threadIdx.x; s < p.n_split; s += blockDim.x) {
      const float ms = s_m[s];
      s_m[s] = (ms == -INFINITY) ? 0.f : __expf(ms - M);
    }
    __syncthreads();
    float a = 0.f, L = 0.f;
#pragma unroll 8
    for (int s = 0; s < p.n_split; ++s) {
      const float f = s_m[s];
      a += pp[s * (2 + D) + 2 + d] * f;
      L += s_l[s] * f;
    }
    p.pooled[(long long)b * D + d] = a / L;
  }
  const int o = blockIdx.x * 8 + warp;   // gridDim.x * 8 == H
  float4 w[4];
  float bo;
  head_load_w<D>(p.w0, p.b0, o, H, w, bo);
  grid_barrier(p.counter, p.base + 1 * G);
  phase = head_fused_layer<D, true, false>(w, bo, p.pooled, nullptr, p.h0, o, H, b_lo, b_hi, head_xs, &head_bar, phase);
  head_load_w<H>(p.w1, p.b1, o, H, w, bo);
  grid_barrier(p.counter, p.base + 2 * G);
  phase = head_fused_layer<H, true, false>(w, bo, p.h0, nullptr, p.h1, o, H, b_lo, b_hi, head_xs, &head_bar, phase);
  head_load_w<H>(p.wr0, p.br0, o, H, w, bo);
  grid_barrier(p.counter, p.base + 3 * G);
  phase = head_fused_layer<H, true, false>(w, bo, p.h1, nullptr, p.h2, o, H, b_lo, b_hi, head_xs, &head_bar, phase);
  head_load_w<H>(p.wr1, p.br1, o, H, w, bo);
  grid_barrier(p.counter, p.base + 4 * G);
  phase = head_fused_layer<H, false, true>(w, bo, p.h2, p.h1, p.h3, o, H, b_lo, b_hi, head_xs, &head_bar, phase);
  if (blockIdx.x == 0) head_load_w<H>(p.wo, p.bo, warp, 2, w, bo);
  grid_barrier(p.counter, p.base + 5 * G);
  if (blockIdx.x == 0) {   // output layer: one CTA per read slice
    phase = head_fused_layer<H, false, false>(w, bo, p.h3, nullptr, p.logits, warp, 2, b_lo, b_hi, head_xs, &head_bar, phase);
    __syncthreads();
    if (p.labels && warp == 0)
      for (int b = b_lo + lane; b < b_hi; b += 32) p.labels[b] = (p.logits[b * 2 + 1] > p.logits[b * 2]) ? 1 : 0;   // ties -> 0
  }
  // The status word is handed to the host by the last CTA to get here (one more arrival on the barrier counter: the host
  // accounts for 6 G per launch), after every slice has written its logits.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(p.counter, 1u);
    if (prev == p.base + 6 * G - 1) publish_status(p.err, p.status_out, p.status_tag);
  }
}

}  // namespace clm
