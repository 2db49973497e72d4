// Probe: does TMA MULTICAST relieve the L2 -> SM stream that bounds block_in?  Every CTA streams the same L2-resident
// "weights" into shared memory (32 KB slots, ring of 3); with a cluster of CS CTAs each CTA issues 1 / CS of every slot's
// boxes with a multicast mask over the cluster, so L2 is read once per cluster and slot.  Reported: bytes LANDING per SM and
// clock with all SMs active.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_multicast_probe tma_multicast_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../chimeralm_b200/csrc/ptx.cuh"
using namespace clm;

__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

template <int CS>   // cluster size; slot = 2 boxes of 16 KB
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int nslot, int n_loads, int rows,
                                               long long* out_cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int SLOT = 32768, BOXES = 2;   // 2 boxes of 16 KB (64 columns x 128 rows) per slot
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 196608);
  uint64_t* empty = full + 16;
  const uint32_t rank = CS > 1 ? ptx::cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < nslot; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], CS); }
    ptx::fence_mbar_init();
  }
  if (CS > 1) ptx::cluster_sync(); else __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int row = ((blockIdx.x / CS) * 128) % rows;
    for (int i = 0; i < n_loads; ++i) {
      const int s = i % nslot; const uint32_t ph = (i / nslot) & 1;
      if (CS > 1) ptx::mbar_wait_cluster(&empty[s], ph ^ 1); else ptx::mbar_wait(&empty[s], ph ^ 1);
      ptx::mbar_expect_tx(&full[s], SLOT);
      for (int b = rank; b < BOXES; b += CS) {
        void* dst = smem + s * SLOT + b * 16384;
        const int c0 = ((2 * i + b) & 3) * 64, c1 = row;
        if (CS > 1) tma_load_2d_mc(dst, &tm, &full[s], c0, c1, (uint16_t)((1u << CS) - 1));
        else ptx::tma_load_2d(dst, &tm, &full[s], c0, c1);
      }
      if (i & 1) { row += 128; if (row >= rows) row = 0; }
    }
  } else if (threadIdx.x == 32) {
    for (int i = 0; i < n_loads; ++i) {
      const int s = i % nslot; const uint32_t ph = (i / nslot) & 1;
      if (CS > 1) ptx::mbar_wait_cluster(&full[s], ph); else ptx::mbar_wait(&full[s], ph);
      if (CS > 1) {
        for (uint32_t r = 0; r < CS; ++r) ptx::mbar_arrive_cluster(ptx::mapa(ptx::smem_u32(&empty[s]), r));
      } else {
        ptx::mbar_arrive(&empty[s]);
      }
    }
  }
  if (CS > 1) ptx::cluster_sync(); else __syncthreads();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CS>
double run(const CUtensorMap& tm, int grid, int nslot, int rows, long long* d_cyc, size_t total_bytes) {
  cudaFuncSetAttribute(probe<CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608 + 512);
  cudaFuncSetAttribute(probe<CS>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  const int n_loads = (int)(total_bytes / 32768);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(64); cfg.dynamicSmemBytes = 196608 + 512;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  for (int rep = 0; rep < 2; ++rep) {
    cudaError_t le = cudaLaunchKernelEx(&cfg, probe<CS>, tm, nslot, n_loads, rows, d_cyc);
    if (le != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(le)); exit(1); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  }
  std::vector<long long> c(grid); cudaMemcpy(c.data(), d_cyc, grid * 8, cudaMemcpyDeviceToHost);
  long long mx = 0; for (auto v : c) mx = v > mx ? v : mx;
  return (double)total_bytes / mx;
}

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
  const size_t total_bytes = 32ull << 20;  // landing per CTA
  for (int grid : {148, 144, 72, 36}) {
    for (int nslot : {3, 6}) {
      printf("grid=%3d ring=%d x 32 KB : no cluster %.1f", grid, nslot, run<1>(tm, grid, nslot, rows, d_cyc, total_bytes));
      if (grid % 2 == 0) printf(" | cluster 2 multicast %.1f", run<2>(tm, grid, nslot, rows, d_cyc, total_bytes));
      
      printf("  B/clk/SM landing\n");
    }
  }
  return 0;
}
