// Probe: how fast can every SM stream the SAME L2-resident weight matrix into shared memory with
// TMA, as a function of ring depth and slot size?  (Design input for the fused block kernels.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_stream_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../chimeralm_b200/csrc/ptx.cuh"
using namespace clm;

template <int BOXES_PER_SLOT>
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int nslot, int n_loads, int rows,
                                               long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int SLOT = BOXES_PER_SLOT * 16384;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 196608);
  uint64_t* empty = full + 16;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslot; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int row = (blockIdx.x * 128) % rows;
    for (int i = 0; i < n_loads; ++i) {
      const int s = i % nslot; const uint32_t ph = (i / nslot) & 1;
      ptx::mbar_wait(&empty[s], ph ^ 1);
      ptx::mbar_expect_tx(&full[s], SLOT);
      for (int b = 0; b < BOXES_PER_SLOT; ++b) {
        ptx::tma_load_2d(smem + s * SLOT + b * 16384, &tm, &full[s], ((i * BOXES_PER_SLOT + b) & 3) * 64, row);
        if ((((i * BOXES_PER_SLOT + b) & 3)) == 3) { row += 128; if (row >= rows) row = 0; }
      }
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < n_loads; ++i) {
      const int s = i % nslot; const uint32_t ph = (i / nslot) & 1;
      ptx::mbar_wait(&full[s], ph);
      ptx::mbar_arrive(&empty[s]);
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int rows = 2304, cols = 256;  // 1.15 MB of bf16 "weights"
  void* w; cudaMalloc(&w, (size_t)rows * cols * 2); cudaMemset(w, 0, (size_t)rows * cols * 2);
  long long* d_cyc; cudaMalloc(&d_cyc, 148 * 8);
  void* fn; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  CUtensorMap tm;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
  cuuint32_t box[2] = {64, 128}, es[2] = {1, 1};
  ((EncodeTiledFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  const int smem = 12 * 32768 / 2 + 1024;  // up to 12 x 16 KB
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 512);
  cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 512);
  (void)smem;
  const size_t total_bytes = 32ull << 20;  // per CTA
  for (int grid : {148, 74, 16}) {
    for (int bps : {1, 2}) {
      for (int nslot : {1, 2, 3, 4, 6, 8, 12}) {
        if (bps * nslot * 16384 > 200 * 1024) continue;
        const int n_loads = (int)(total_bytes / (bps * 16384));
        for (int rep = 0; rep < 2; ++rep) {
          if (bps == 1) probe<1><<<grid, 64, 196608 + 512>>>(tm, nslot, n_loads, rows, d_cyc);
          else probe<2><<<grid, 64, 196608 + 512>>>(tm, nslot, n_loads, rows, d_cyc);
          { cudaError_t le = cudaGetLastError(); if (le != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(le)); return 1; } }
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        }
        std::vector<long long> c(grid); cudaMemcpy(c.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
        printf("grid=%3d slot=%2dKB nslot=%2d inflight=%3dKB : %.1f B/clk/SM\n", grid, bps * 16, nslot, bps * 16 * nslot,
               (double)total_bytes / mx);
      }
    }
  }
  return 0;
}
