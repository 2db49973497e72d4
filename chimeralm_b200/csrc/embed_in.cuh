// First half of block 0 without a GEMM (SURVEY.md A.2 + A.6 head + A.3; reference call sites
// chimeralm/models/components/hyena.py:249-255 -> HF HyenaEmbeddings / HyenaBlock / HyenaOperator.in_proj + short_filter).
//
// The input of block 0 is the embedding row of a token id, so everything up to and including in_proj is a function of
// the id alone: u[t] = W_in' . LN1(E[id[t]]) + b_in' takes one of `vocab_rows` (16) values per channel.  The table
// U[channel][id] is built once at clm_finalize - from the SAME bf16 operands block_in_kernel feeds the tensor cores
// (normalised embedding rows, folded weights) with fp32 accumulation, so both paths agree to fp32 rounding - and block 0's
// first half becomes
//
//   uc[c,t] = w[c,0] U[c][id[t-2]] + w[c,1] U[c][id[t-1]] + w[c,2] U[c][id[t]] + cb[c]     (u = 0 for t < 0)
//   x0, x1, v = the three channel groups;   outputs x0 and v * x1 channel-major [B][256][Tp]
//
// i.e. a table lookup, 9 FMAs and a product per (channel, token): no 256 x 768 GEMM over 262 k tokens, no xn read, and the
// embedding kernel no longer writes xn.  Per token it reads 1 id byte and writes 1 KB - HBM-bound on the stores.
//
// Mapping: a thread owns 8 consecutive tokens (one 16-byte store per output row) and keeps their 10 ids in registers while
// it walks the block's 64 channels; the block's slice of U sits in shared memory as [group][channel][id], so the lanes of a
// warp (same channel, different ids) hit consecutive words - no bank conflicts.  Bounds: 3.75 LDS and 10 FP32 ops per
// (channel, token) against 4 B of stores.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "gemm_tcgen05.cuh"

namespace clm {

namespace ei {
// CG = 64 channels per block: K2 then needs 131 k threads, fewer than the 148 x 7 x 128 = 133 k that are resident at once - one
// even wave.  With CG = 32 the 262 k threads ran as 1.54 waves of 9 blocks per SM (measured 0.076 ms against 0.045 of stores).
constexpr int D = 256, NV = 16, CG = 64, THREADS = 128, TOK = 8, BLOCK_TOK = THREADS * TOK;
}

// U[ch][v] = b'[ch] + sum_k bf16(W'[ch][k]) * xn_bf16[v][k]     (ch < 768, v < rows <= 16; one warp per channel)
__global__ void __launch_bounds__(32) embed_in_table_kernel(const float* __restrict__ Wf, const float* __restrict__ bf,
                                                            const __nv_bfloat16* __restrict__ En, float* __restrict__ U,
                                                            int rows) {
  const int ch = blockIdx.x, lane = threadIdx.x;
  float w[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) w[i] = __bfloat162float(__float2bfloat16(Wf[(long long)ch * ei::D + lane + 32 * i]));
  for (int v = 0; v < ei::NV; ++v) {
    float acc = 0.f;
    if (v < rows) {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(w[i], __bfloat162float(En[v * ei::D + lane + 32 * i]), acc);
    }
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) U[ch * ei::NV + v] = v < rows ? acc + bf[ch] : 0.f;
  }
}

struct EmbedInParams {
  const void* ids;        // [B][T] token ids (IdT)
  const float* U;         // [768][16] table (embed_in_table_kernel)
  const float* cw;        // [768][3] short filter taps
  const float* cb;        // [768]    short filter bias
  const float* vx_scale;  // [256] power-of-two factor on v * x1 (fp16 rows of the tensor-core conv) or nullptr
  __nv_bfloat16* x0;      // [B][256][Tp]
  void* vx;               // [B][256][Tp] bf16, or fp16 when vx_f16
  int B, T, Tp, vx_f16, rows;
  int* err;               // bit 0 raised for an id outside [0, rows) (the embedding kernel's check, when that kernel is not run)
};

// E [rows][256] -> the first 32-row group of an R32 residual buffer (rows >= `rows` zero): block_mlp's table for block 0
__global__ void __launch_bounds__(256) embed_r32_table_kernel(const float* __restrict__ E, float* __restrict__ out, int rows) {
  const int v = blockIdx.x, col = threadIdx.x;
  out[ptx::r32_off(v, col)] = v < rows ? E[v * ei::D + col] : 0.f;
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// VX_F16: v * x1 as fp16 (operand rows of the tensor-core conv) instead of bf16
template <typename IdT, bool VX_F16>
__global__ void __launch_bounds__(ei::THREADS) embed_in_kernel(EmbedInParams p) {
  using namespace ei;
  // [group][channel][id], 17 words per channel: slot 16 holds 0 = "before the read" (the causal zero padding), so the
  // lookup needs no select; lanes of a warp read <= 17 consecutive words of one channel: no bank conflicts
  constexpr int NS = NV + 1;
  ptx::griddep_launch();
  __shared__ float s_u[3][CG][NS];
  __shared__ __align__(16) float4 s_c[3][CG];   // (w0, w1, w2, cb) per group and channel; the v group carries vx_scale
  const int b = blockIdx.z, c0 = blockIdx.y * CG;
  for (int i = threadIdx.x; i < 3 * CG * NS; i += THREADS) {
    const int g = i / (CG * NS), rem = i - g * (CG * NS), cl = rem / NS, v = rem - cl * NS;
    s_u[g][cl][v] = v < NV ? __ldg(p.U + (g * D + c0 + cl) * NV + v) : 0.f;
  }
  for (int i = threadIdx.x; i < 3 * CG; i += THREADS) {
    const int g = i / CG, ch = g * D + c0 + (i - g * CG);
    const float a = (g == 2 && p.vx_scale) ? __ldg(p.vx_scale + c0 + (i - g * CG)) : 1.f;
    s_c[g][i - g * CG] = make_float4(__ldg(p.cw + ch * 3) * a, __ldg(p.cw + ch * 3 + 1) * a, __ldg(p.cw + ch * 3 + 2) * a,
                                     __ldg(p.cb + ch) * a);
  }
  ptx::griddep_wait();   // the ids come from the encode kernel (the tables above are constants)
  const int t0 = blockIdx.x * BLOCK_TOK + threadIdx.x * TOK;
  // shared-memory byte address of s_u[0][cl][id] for tokens t0 - 2 .. t0 + 7, bumped by one channel per iteration (the group
  // offset is an immediate): one LDS per lookup and 10 adds per channel instead of per-lookup address arithmetic
  uint32_t ia[TOK + 2];
  const IdT* row = reinterpret_cast<const IdT*>(p.ids) + (long long)b * p.T;
  const uint32_t su0 = (uint32_t)__cvta_generic_to_shared(&s_u[0][0][0]);
#pragma unroll
  for (int k = 0; k < TOK + 2; ++k) {
    const int t = t0 - 2 + k;
    long long v = 0;                                     // past the end of the read: any valid row (masked below)
    if (t < 0) v = NV;                                   // before the read: the zero slot
    else if (t < p.T) {
      v = (long long)row[t];
      if (v < 0 || v >= p.rows) {                        // nn.Embedding would raise: flag it, substitute row 0
        if (p.err && blockIdx.y == 0 && k >= 2) atomicOr(p.err, 1);
        v = 0;
      }
    }
    ia[k] = su0 + (uint32_t)v * 4u;
  }
  __syncthreads();
  if (t0 >= p.Tp) return;
  const bool full = t0 + TOK <= p.T;   // all 8 tokens inside the read (everything but a read's last thread or two)
  __nv_bfloat16* x0p = p.x0 + ((size_t)b * D + c0) * p.Tp + t0;
  uint16_t* vxp = reinterpret_cast<uint16_t*>(p.vx) + ((size_t)b * D + c0) * p.Tp + t0;
#pragma unroll 2
  for (int cl = 0; cl < CG; ++cl) {
    float o[3][TOK];
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const float4 k4 = s_c[g][cl];
      float u[TOK + 2];
#pragma unroll
      for (int k = 0; k < TOK + 2; ++k) u[k] = lds_f32(ia[k] + uint32_t(g * CG * NS * 4));
#pragma unroll
      for (int j = 0; j < TOK; ++j) o[g][j] = fmaf(k4.x, u[j], fmaf(k4.y, u[j + 1], fmaf(k4.z, u[j + 2], k4.w)));
    }
#pragma unroll
    for (int k = 0; k < TOK + 2; ++k) ia[k] += NS * 4;
    float m[TOK];
#pragma unroll
    for (int j = 0; j < TOK; ++j) m[j] = o[2][j] * o[1][j];
    if (!full) {   // whole 128-token rows of the tensor-core conv: ZERO past the end of the read
#pragma unroll
      for (int j = 0; j < TOK; ++j)
        if (t0 + j >= p.T) { m[j] = 0.f; o[0][j] = 0.f; }
    }
    *reinterpret_cast<uint4*>(x0p) = make_uint4(pack_bf16(o[0][0], o[0][1]), pack_bf16(o[0][2], o[0][3]),
                                                pack_bf16(o[0][4], o[0][5]), pack_bf16(o[0][6], o[0][7]));
    if (VX_F16)
      *reinterpret_cast<uint4*>(vxp) = make_uint4(pack_f16(m[0], m[1]), pack_f16(m[2], m[3]), pack_f16(m[4], m[5]), pack_f16(m[6], m[7]));
    else
      *reinterpret_cast<uint4*>(vxp) = make_uint4(pack_bf16(m[0], m[1]), pack_bf16(m[2], m[3]), pack_bf16(m[4], m[5]), pack_bf16(m[6], m[7]));
    x0p += p.Tp;
    vxp += p.Tp;
  }
}

}  // namespace clm
