// Correctness probe for tcgen05.mma kind::f16 with MN-MAJOR operands in the no-swizzle (INTERLEAVE) core-matrix
// layout -- the form the weight-gradient GEMM of the BPTT kernels needs (contraction over agent rows, both operand
// images stored row-major-in-core-matrix for the forward/backward GEMMs that contract over features):
//   D[m][n] = sum_k A[m][k] * B[n][k],  A stored as [mgroup][kgroup][8 k][8 m] halfs, B as [ngroup][kgroup][8 k][8 n]
// i.e. core matrix = 8 K-rows of 16 contiguous bytes along MN.  Tries the candidate (LBO, SBO) assignments and
// instruction-descriptor major bits and prints which one reproduces the host result.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mn_major mn_major.cu && ./mn_major
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
               "l"(a), "l"(b), "r"(idesc), "r"(acc)
               : "memory");
}

// variant bit 0: swap LBO/SBO; bit 1: k-step advance uses the other stride
__global__ void __launch_bounds__(128, 1) probe(const __half* a_img, const __half* b_img, float* out, int variant) {
  __shared__ __align__(1024) __half sA[M * K];
  __shared__ __align__(1024) __half sB[N * K];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) sA[i] = a_img[i];
  for (int i = threadIdx.x; i < N * K; i += blockDim.x) sB[i] = b_img[i];
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;
  if (threadIdx.x == 0) {
    // D = f32, A = B = f16, a_major (bit 15) = b_major (bit 16) = 1 (MN-major), N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    // image: [mgroup][kgroup = K/8][8 k][8 m]: MN-group stride = (K/8)*128 B, K-group stride = 128 B
    const uint32_t mn_stride = (K / 8) * 128, k_stride = 128;
    const uint32_t lbo = (variant & 1) ? mn_stride : k_stride;     // variant 0: LBO = K-group stride, SBO = MN-group stride
    const uint32_t sbo = (variant & 1) ? k_stride : mn_stride;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * k_stride, lbo, sbo);
      const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * k_stride, lbo, sbo);
      mma(tm, da, db, idesc, ks != 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  uint32_t done = 0;
  int spins = 0;
  while (!done && spins < (1 << 22)) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
    ++spins;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tm + c0 + ((uint32_t)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(64) : "memory");
}

int main() {
  static __half hA[M * K], hB[N * K];
  static float fA[M][K], fB[N][K], want[M][N], got[M * N];
  srand(7);
  for (int m = 0; m < M; ++m)
    for (int k = 0; k < K; ++k) {
      const float v = (float)((rand() % 17) - 8) / 8.f;
      fA[m][k] = v;
      hA[(((m / 8) * (K / 8) + k / 8) * 8 + k % 8) * 8 + m % 8] = __float2half(v);
    }
  for (int n = 0; n < N; ++n)
    for (int k = 0; k < K; ++k) {
      const float v = (float)((rand() % 13) - 6) / 4.f;
      fB[n][k] = v;
      hB[(((n / 8) * (K / 8) + k / 8) * 8 + k % 8) * 8 + n % 8] = __float2half(v);
    }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += fA[m][k] * fB[n][k];
      want[m][n] = s;
    }
  __half *dA, *dB;
  float* dO;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dO, sizeof(got));
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
  int good = -1;
  for (int variant = 0; variant < 2; ++variant) {
    cudaMemset(dO, 0, sizeof(got));
    probe<<<1, 128>>>(dA, dB, dO, variant);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d: %s\n", variant, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(got, dO, sizeof(got), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < N; ++n) worst = fmax(worst, fabs((double)got[m * N + n] - want[m][n]));
    printf("variant %d (%s): max |D - host| = %g  %s\n", variant,
           variant == 0 ? "LBO = K-group stride 128 B, SBO = MN-group stride" : "LBO = MN-group stride, SBO = K-group stride 128 B",
           worst, worst < 1e-3 ? "MATCH" : "mismatch");
    if (worst < 1e-3) good = variant;
  }
  printf("MN-major no-swizzle: %s\n", good < 0 ? "NO VARIANT MATCHED" : (good == 0 ? "variant 0 is correct" : "variant 1 is correct"));
  return good < 0;
}
