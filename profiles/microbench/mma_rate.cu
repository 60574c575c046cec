// Raw tcgen05.mma issue rate on B200 for the instruction streams policy_tc.cu uses (no TMA, no epilogue):
// is the tensor pipe itself (operand fetch from shared memory, accumulator dependency) the pacing stage?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu && ./mma_rate
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}
// mode 0: policy_tc stream (hi.hi, lo.hi, hi.lo per k-step, no-swizzle core-matrix layout, 4 stages of 48 KB)
// mode 1: one descriptor pair repeated (no-swizzle)
// mode 2: policy_tc stream, alternating between the two accumulators every MMA
// mode 3: one descriptor pair repeated, 128-byte swizzle layout
// mode 4: mode 0 with N = 128
__global__ void __launch_bounds__(64, 1) k(int mode, int iters, long long* cyc) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < 4 * 49152 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 32) {
    const int N = mode == 4 ? 128 : 256;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t base = smem_u32(smem);
    const long long t0 = clock64();
    int n = 0;
    for (int it = 0; it < iters; ++it) {
      const uint32_t a0 = base + (it & 3) * 49152, b0 = a0 + 16384;
      for (int ks = 0; ks < 2; ++ks) {
        if (mode == 0 || mode == 2 || mode == 4) {
          const uint64_t ah = make_desc(a0 + ks * 4096, 2048, 128, 0), al = make_desc(a0 + 8192 + ks * 4096, 2048, 128, 0);
          const uint64_t bh = make_desc(b0 + ks * 8192, 4096, 128, 0), bl = make_desc(b0 + 16384 + ks * 8192, 4096, 128, 0);
          const uint32_t d0 = tm, d1 = mode == 2 ? tm + 256 : tm;
          mma(d0, ah, bh, idesc, 1); mma(d1, al, bh, idesc, 1); mma(d0, ah, bl, idesc, 1);
        } else if (mode == 1) {
          const uint64_t ah = make_desc(base, 2048, 128, 0), bh = make_desc(base + 16384, 4096, 128, 0);
          mma(tm, ah, bh, idesc, 1); mma(tm, ah, bh, idesc, 1); mma(tm, ah, bh, idesc, 1);
        } else {   // 128B swizzle, K-major: rows of 128 B, 8-row groups 1024 B apart; K slice = 32 B inside the row
          const uint64_t ah = make_desc(base + ks * 32, 16, 1024, 2), bh = make_desc(base + 16384 + ks * 32, 16, 1024, 2);
          mma(tm, ah, bh, idesc, 1); mma(tm, ah, bh, idesc, 1); mma(tm, ah, bh, idesc, 1);
        }
        n += 3;
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    }
    const long long t1 = clock64();
    cyc[blockIdx.x * 2] = t1 - t0;
    cyc[blockIdx.x * 2 + 1] = n;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512) : "memory");
}
int main() {
  long long* cyc;
  cudaMalloc(&cyc, 148 * 16);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 49152);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int mode = 0; mode < 5; ++mode)
    for (int grid : {1, 148}) {
      long long h[296];
      float ms = 0;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<<<grid, 64, 4 * 49152>>>(mode, 4096, cyc);
        cudaEventRecord(e1);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d failed: %s\n", mode, cudaGetErrorString(cudaGetLastError())); return 1; }
        cudaEventElapsedTime(&ms, e0, e1);
      }
      cudaMemcpy(h, cyc, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
      printf("mode %d grid %3d: %.1f cycles/MMA (SM 0), kernel %.3f ms -> %.1f ns/MMA\n", mode, grid, (double)h[0] / h[1], ms,
             1e6 * ms / h[1]);
    }
  return 0;
}
