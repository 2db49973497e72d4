// Probe: MUFU.TANH throughput per scheduler, fp32 vs packed f16x2 (two MUFU.TANH.F16 per instruction), 2 warps per scheduler
// like the GELU epilogue of the fused block tail.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_probe mufu_tanh_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
template <int MODE>
__global__ void __launch_bounds__(256, 1) probe(int iters, float seed, long long* out, float* sink) {
  float x[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { x[i] = seed * (threadIdx.x + i + 1) * 1e-3f; h[i] = 0x2c003800u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      else asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
}
int main() {
  long long* d; float* sink; cudaMalloc(&d, 148 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 2000;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) probe<0><<<148, 256>>>(iters, 1.0f, d, sink); else probe<1><<<148, 256>>>(iters, 1.0f, d, sink);
      cudaDeviceSynchronize();
    }
    long long h[148]; cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    // per scheduler: 2 warps x 8 instructions per iteration
    double cyc_per_warp_instr = (double)h[0] / (iters * 8.0 * 2.0);
    printf("%s: %.2f cycles per warp instruction per scheduler = %.2f cycles per 32 results\n", mode ? "tanh.approx.f16x2" : "tanh.approx.f32  ",
           cyc_per_warp_instr, mode ? cyc_per_warp_instr / 2 : cyc_per_warp_instr);
  }
  return 0;
}
