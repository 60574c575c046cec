// SFU (MUFU) throughput on one SM as a function of resident warps: the number the LSTM epilogue's pacing depends on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu && ./mufu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
  float a = threadIdx.x * 1e-3f, b = a + 1.f, c = a + 2.f, d = a + 3.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    if (MODE == 0) {   // ex2 only, 4 independent chains
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d));
    } else if (MODE == 1) {   // rcp only
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(c));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(d));
    } else {           // ex2 + rcp pairs (sigmoid shape)
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b));
      asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(b));
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {4, 8, 16, 32}) {
      long long h[148];
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, cyc);
        if (mode == 2) k<2><<<148, warps * 32>>>(out, iters, cyc);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double n = (double)iters * 4 * warps;      // warp-level MUFU instructions per SM
      printf("mode %d warps/SM %2d: %.2f cycles per warp-MUFU per SM  (%.2f lanes/clk/SM)\n", mode, warps, h[0] / n,
             32.0 * n / h[0]);
    }
  return 0;
}
