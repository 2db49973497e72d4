// Probe: per-SM cost of moving one 128 KB fp32 residual tile (128 rows x 256 columns, R32 blocked layout, contiguous)
// in and out of an SM while every other SM does the same - the tile boundary of block_mlp_kernel.
//   mode 0  st.global.v4 from 256 threads (a warp writes 512 contiguous bytes per instruction)      [what E3 does today]
//   mode 1  st.shared.v4 into a staging buffer + cp.async.bulk shared -> global, 32 KB pieces, 2 buffers
//   mode 2  ld.global.v4 into registers, 32 per thread                                              [what E1 does today]
//   mode 3  cp.async.bulk global -> shared (32 KB pieces, 2 buffers, mbarrier) + ld.shared.v4
// Prints cycles per tile (median over CTAs).  Design input for DESIGN.md 4.2.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o resid_io_probe resid_io_probe.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "../../chimeralm_b200/csrc/ptx.cuh"
using namespace clm;

constexpr int TILE_BYTES = 128 * 1024, PIECE = 32 * 1024;

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ptx::smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ptx::smem_u32(sdst)),
               "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(float* buf, int tiles_per_cta, long long* out_cycles, float* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * PIECE);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_mbar_init();
  }
  __syncthreads();
  float acc = 0.f;
  const long long t0 = clock64();
  for (int t = 0; t < tiles_per_cta; ++t) {
    float* tile = buf + ((size_t)t * gridDim.x + blockIdx.x) * (TILE_BYTES / 4);
    if (MODE == 0) {
      // thread (row r = tid & 127, column half hf = tid >> 7): 32 float4, column group c4 at ((r / 32) * 64 + c4) * 128 + (r % 32) * 4
      const int r = tid & 127, hf = tid >> 7;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        *reinterpret_cast<float4*>(tile + ((r >> 5) * 64 + hf * 32 + j) * 128 + (r & 31) * 4) = make_float4(t, j, r, hf);
    } else if (MODE == 1) {
      const int r = tid & 127, hf = tid >> 7;
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        uint8_t* st = smem + (ci & 1) * PIECE;
        if (ci >= 2) {
          if (tid == 0) ptx::tma_store_wait_read<1>();
          __syncthreads();
        }
        // piece ci = column groups [8 ci, 8 ci + 8) of both halves: segment (g, hf) = 8 groups x 512 B = 4 KB contiguous
#pragma unroll
        for (int j = 0; j < 8; ++j)
          *reinterpret_cast<float4*>(st + (((r >> 5) * 2 + hf) * 8 + j) * 512 + (r & 31) * 16) = make_float4(t, j, r, hf);
        ptx::fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
          for (int seg = 0; seg < 8; ++seg)
            bulk_store(tile + ((seg >> 1) * 64 + (seg & 1) * 32 + 8 * ci) * 128, st + seg * 4096, 4096);
          ptx::tma_store_commit();
        }
      }
      if (tid == 0) ptx::tma_store_wait_read<0>();
      __syncthreads();
    } else if (MODE == 2) {
      const int r = tid & 127, hf = tid >> 7;
      float4 v[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = *reinterpret_cast<const float4*>(tile + ((r >> 5) * 64 + hf * 32 + j) * 128 + (r & 31) * 4);
#pragma unroll
      for (int j = 0; j < 32; ++j) acc += v[j].x + v[j].y + v[j].z + v[j].w;
    } else {
      const int r = tid & 127, hf = tid >> 7;
      // pieces 0, 1 are in flight from the previous iteration (or the prologue); consume, then refill with pieces 2, 3 / the next tile
      if (t == 0 && tid == 0)
        for (int pc = 0; pc < 2; ++pc) {
          ptx::mbar_expect_tx(&bars[pc], PIECE);
          for (int seg = 0; seg < 8; ++seg)
            bulk_load(smem + pc * PIECE + seg * 4096, tile + ((seg >> 1) * 64 + (seg & 1) * 32 + 8 * pc) * 128, 4096, &bars[pc]);
        }
#pragma unroll 1
      for (int ci = 0; ci < 4; ++ci) {
        const int bsel = ci & 1;
        ptx::mbar_wait(&bars[bsel], ((t * 4 + ci) >> 1) & 1);
        const uint8_t* st = smem + bsel * PIECE;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(st + (((r >> 5) * 2 + hf) * 8 + j) * 512 + (r & 31) * 16);
          acc += v.x + v.y + v.z + v.w;
        }
        __syncthreads();
        const int nci = ci + 2;   // refill this buffer: piece ci + 2 of this tile, or piece ci - 2 of the next
        const bool nxt = nci >= 4;
        if (tid == 0 && (!nxt || t + 1 < tiles_per_cta)) {
          float* src = nxt ? buf + ((size_t)(t + 1) * gridDim.x + blockIdx.x) * (TILE_BYTES / 4) : tile;
          const int pc = nci & 3;
          ptx::mbar_expect_tx(&bars[bsel], PIECE);
          for (int seg = 0; seg < 8; ++seg)
            bulk_load(smem + bsel * PIECE + seg * 4096, src + ((seg >> 1) * 64 + (seg & 1) * 32 + 8 * pc) * 128, 4096, &bars[bsel]);
        }
      }
    }
  }
  if (MODE == 1 && tid == 0) ptx::tma_store_wait<0>();
  __syncthreads();
  if (tid == 0) out_cycles[blockIdx.x] = (clock64() - t0) / tiles_per_cta;
  if (acc == 123.456f) sink[0] = acc;
  (void)warp; (void)lane;
}

template <int MODE>
void run(const char* name, float* buf, int grid, int tiles, long long* d_cyc, float* sink) {
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * PIECE + 64);
  std::vector<long long> h(grid);
  for (int rep = 0; rep < 3; ++rep) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<MODE><<<grid, 256, 2 * PIECE + 64>>>(buf, tiles, d_cyc, sink);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    cudaMemcpy(h.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double gb = (double)grid * tiles * TILE_BYTES / 1e9;
    printf("%-34s grid %3d: %6lld cycles/tile (median; max %lld)  %7.1f GB/s chip  %5.1f B/clk/SM  [%s]\n", name, grid, h[grid / 2],
           h[grid - 1], gb / (ms / 1e3), (double)TILE_BYTES / h[grid / 2], cudaGetErrorString(e));
  }
}

int main() {
  const int tiles = 14;
  float* buf; cudaMalloc(&buf, (size_t)148 * tiles * TILE_BYTES);
  cudaMemset(buf, 0, (size_t)148 * tiles * TILE_BYTES);
  long long* d_cyc; cudaMalloc(&d_cyc, 148 * 8);
  float* sink; cudaMalloc(&sink, 4);
  for (int grid : {148, 37}) {
    run<0>("st.global.v4", buf, grid, tiles, d_cyc, sink);
    run<1>("st.shared + bulk store (32 KB x2)", buf, grid, tiles, d_cyc, sink);
    run<2>("ld.global.v4 -> regs", buf, grid, tiles, d_cyc, sink);
    run<3>("bulk load (32 KB x2) + ld.shared", buf, grid, tiles, d_cyc, sink);
  }
  return 0;
}
