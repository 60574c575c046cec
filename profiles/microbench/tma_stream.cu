// L2 -> shared-memory bulk-copy stream per SM as a function of ring depth and chunk size (no MMA, no epilogue):
// does the operand stream of lstm_tc_kernel saturate because of bytes in flight (Little's law: 192 KB of ring /
// round-trip latency) or because of the chip's L2 bandwidth?  One persistent CTA per SM; thread 0 produces
// (mbarrier expect_tx + cp.async.bulk), thread 32 consumes (wait full, arrive empty), like the kernel's hand-shake
// minus the tcgen05.commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_stream tma_stream.cu && ./tma_stream
// Written in round 1 (GPU budget exhausted before it could run): first thing to run in round 2.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar) : "memory");
}

// src: `span` bytes that every CTA streams `nchunk` chunks from (CTA-dependent offset so that neighbours do not read
// the very same lines at the same time unless shared != 0, which mimics the weight chunks every SM re-reads)
__global__ void __launch_bounds__(64, 1) k(const unsigned char* src, size_t span, int nstage, int chunk, int nchunk, int shared,
                                           long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[32];
  const uint32_t full = smem_u32(bars), empty = smem_u32(bars + 16);
  if (threadIdx.x == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const size_t base = shared ? 0 : ((size_t)blockIdx.x * 1315423911ull) % (span - (size_t)chunk * 64);
  const long long t0 = clock64();
  if (threadIdx.x == 0) {
    for (int g = 0; g < nchunk; ++g) {
      const int s = g % nstage;
      mbar_wait(empty + 8 * s, ((g / nstage) & 1) ^ 1);
      mbar_expect_tx(full + 8 * s, chunk);
      const size_t off = (base + (size_t)(g % 64) * chunk) & ~(size_t)127;
      bulk_g2s(smem_u32(smem + (size_t)s * chunk), src + off, chunk, full + 8 * s);
    }
  } else if (threadIdx.x == 32) {
    for (int g = 0; g < nchunk; ++g) {
      const int s = g % nstage;
      mbar_wait(full + 8 * s, (g / nstage) & 1);
      mbar_arrive(empty + 8 * s);
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const size_t span = (size_t)64 << 20;            // 64 MB source: L2-resident after the first pass
  unsigned char* src;
  long long* cyc;
  cudaMalloc(&src, span);
  cudaMemset(src, 1, span);
  cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 196608);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int cfgs[][2] = {{4, 49152}, {3, 65536}, {6, 32768}, {8, 24576}, {12, 16384}, {2, 49152}, {1, 49152}, {4, 16384}};
  for (int shared = 0; shared < 2; ++shared)
    for (auto& c : cfgs) {
      const int nstage = c[0], chunk = c[1], nchunk = 4096;
      float ms = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<<<148, 64, (size_t)nstage * chunk>>>(src, span, nstage, chunk, nchunk, shared, cyc);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      const double bytes = 148.0 * nchunk * chunk;
      printf("%s source, %2d stages x %5d B: %7.1f GB/s per SM, %6.2f TB/s chip, %.3f us per chunk\n",
             shared ? "shared  " : "distinct", nstage, chunk, bytes / 148 / (ms * 1e-3) / 1e9, bytes / (ms * 1e-3) / 1e12,
             ms * 1e3 / nchunk);
    }
  return 0;
}
