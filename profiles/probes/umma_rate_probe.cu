// Probe: cycles per tcgen05.mma (cta_group::1, bf16 -> fp32, M = 128) when ONE thread issues them back to
// back on static shared-memory operands - the tensor pipe's real pace for the shapes the fused kernels use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o umma_probe umma_rate_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../../chimeralm_b200/csrc/ptx.cuh"
using namespace clm;

// mode 0: SS, A and B K-major SW128 in smem; mode 1: TS, A in TMEM; mode 2: SS with MN-major A
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) probe(int n_mma, int same_acc, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&tptr);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tb = tptr;
  if (threadIdx.x == 0) {
    const uint32_t sA = ptx::smem_u32(smem), sB = sA + 16384;
    constexpr uint32_t idesc = MODE == 2 ? ptx::idesc_bf16_f32_amn(128, N) : ptx::idesc_bf16_f32(128, N);
    const uint64_t da = MODE == 2 ? ptx::smem_desc_mn_sw128(sA, 8192, 1024) : ptx::smem_desc_k_sw128(sA);
    const uint64_t db = ptx::smem_desc_k_sw128(sB);
    long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const int k = i & 3;
      const uint32_t d = tb + (same_acc ? 0 : ((i >> 2) & 1) * 256);
      if (MODE == 3) {   // the fused block tail's fc1: 16 K-steps walk 128 TMEM columns of A and four 16 KB k-blocks of B
        const int kk = i & 15;
        ptx::umma_f16_ts(d, tb + 384 + kk * 8, ptx::smem_desc_k_sw128(sB + (kk >> 2) * 8192) + 2 * (kk & 3), idesc, i >= 16);
      } else if (MODE == 1) ptx::umma_f16_ts(d, tb + 384 + k * 8, db + 2 * k, idesc, i >= 8);
      else if (MODE == 2) ptx::umma_f16(d, da + 128 * k, db + 2 * k, idesc, i >= 8);
      else ptx::umma_f16(d, da + 2 * k, db + 2 * k, idesc, i >= 8);
    }
    long long t1 = clock64();
    ptx::umma_commit(&bar);
    ptx::mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tb);
}

template <int N, int MODE>
void run(const char* name, int grid, long long* d) {
  cudaFuncSetAttribute(probe<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int n = 2048;
  for (int same = 0; same < 2; ++same) {
    for (int rep = 0; rep < 2; ++rep) {
      probe<N, MODE><<<grid, 128, 64 * 1024>>>(n, same, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
    }
    std::vector<long long> h(grid * 2);
    cudaMemcpy(h.data(), d, grid * 16, cudaMemcpyDeviceToHost);
    long long iss = 0, tot = 0;
    for (int i = 0; i < grid; ++i) { iss = h[2 * i] > iss ? h[2 * i] : iss; tot = h[2 * i + 1] > tot ? h[2 * i + 1] : tot; }
    printf("%-28s grid=%3d %s: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (nominal %d)\n", name, grid,
           same ? "one accumulator " : "two accumulators", (double)iss / n, (double)tot / n, N / 2);
  }
}

int main() {
  long long* d; cudaMalloc(&d, 148 * 16);
  for (int grid : {1, 148}) {
    run<256, 0>("SS K-major   M128 N256", grid, d);
    run<128, 0>("SS K-major   M128 N128", grid, d);
    run<144, 0>("SS K-major   M128 N144", grid, d);
    run<128, 1>("TS (A tmem)  M128 N128", grid, d);
    run<256, 1>("TS (A tmem)  M128 N256", grid, d);
    run<256, 2>("SS A MN-major M128 N256", grid, d);
    run<128, 3>("TS 16 K-steps M128 N128", grid, d);
    run<64, 0>("SS K-major   M128 N64", grid, d);
    run<64, 1>("TS (A tmem)  M128 N64", grid, d);
    run<96, 1>("TS (A tmem)  M128 N96", grid, d);
    run<32, 1>("TS (A tmem)  M128 N32", grid, d);
  }
  return 0;
}
