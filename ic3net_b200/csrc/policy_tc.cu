// CommNet / IC3Net policy step on the 5th-generation tensor cores ("policy v2", H = 128).
//
// Math.  With S the gated hidden-state mean (comm.py:181-205), the LSTM pre-activations of
// comm.py:206-218 are ONE contraction over K = 384:
//   gates = W_ih (x + C S + c_b) + W_hh h + b_ih + b_hh
//         = [x | S | h] . [W_ih ; W_ih C ; W_hh]^T + (b_ih + b_hh + W_ih c_b)
// (W_ih C is formed once per weight update in float64.)  fp32 accuracy on fp16 tensor cores:
// a * 16 = a_hi + a_lo and w * 256 = w_hi + w_lo with fp16 halves (11 significant bits each,
// power-of-two pre-scaling keeps both halves in the fp16 normal range for |a| < 4094, |w| < 255);
//   D = a_hi w_hi + a_lo w_hi + a_hi w_lo   (3 tcgen05.mma kind::f16, fp32 accumulate in TMEM)
// drops only a_lo w_lo (2^-22 relative) and gates = D * 2^-12 + bias.
//
// Data movement.  Both operands are stored in global memory as ready-made shared-memory
// images in the no-swizzle K-major core-matrix layout (8 rows x 16 bytes per core matrix),
// so a stage of the pipeline is two plain cp.async.bulk copies (A: 16 KB, B: 32 KB) that
// complete on an mbarrier, and the MMA descriptors are fixed offsets into the stage:
//   prep kernel   x, h (fp32), gate masks  ->  A image [tile][12 chunks][hi,lo][4 kcore][16 rcore][8][8]
//   pack kernel   weights                  ->  B image [2 halves][12 chunks][hi,lo][4 kcore][32 ncore][8][8]
//   lstm kernel   one CTA per (128-row tile, 256-column half): warp 4 = bulk-copy producer +
//                 TMEM allocator, warp 5 = MMA issuer (one thread), warps 0-3 = epilogue
//                 (tcgen05.ld 32 lanes x 16 columns -> LSTM cell -> c', h'); 2 CTAs per SM
//                 (256 TMEM columns, 97 KB smem each) overlap each other's prologue/epilogue.
//   heads kernel  value / action heads + sampling from h' (warp per row).
#include <cuda_fp16.h>

#include "ic3_common.cuh"
#include "policy_heads.cuh"
#include "policy_internal.h"

namespace {

constexpr int TC_H = 128;
constexpr int TC_K = 384;               // [x | S | h]
constexpr int TC_KC = 32;               // K per pipeline stage
constexpr int TC_NCHUNK = TC_K / TC_KC; // 12
constexpr int TC_M = 128;               // rows per tile
constexpr int TC_NH = 256;              // gate columns per CTA (64 hidden units x i,f,g,o)
constexpr int A_CHUNK_BYTES = 2 * TC_M * TC_KC * 2;   // hi + lo = 16384
constexpr int B_CHUNK_BYTES = 2 * TC_NH * TC_KC * 2;  // 32768
constexpr int STAGE_BYTES = A_CHUNK_BYTES + B_CHUNK_BYTES;
constexpr int NSTAGE = 2;
constexpr int A_TILE_HALFS = TC_NCHUNK * A_CHUNK_BYTES / 2;   // 98304
constexpr float SCALE_A = 16.f, SCALE_B = 256.f, INV_SCALE = 1.f / 4096.f;
constexpr int TC_THREADS = 192;
constexpr uint32_t WATCHDOG_SPINS = 1u << 22;

// ---- image addressing (in halfs) ---------------------------------------------------------
__host__ __device__ __forceinline__ size_t a_img_off(int tile, int k, int r, int part) {
  const int c = k >> 5, kk = k & 31;
  return (size_t)tile * A_TILE_HALFS + (((((size_t)(c * 2 + part) * 4 + (kk >> 3)) * 16 + (r >> 3)) * 8 + (r & 7)) * 8) +
         (kk & 7);
}
__host__ __device__ __forceinline__ size_t b_img_off(int nh, int k, int n, int part) {
  const int c = k >> 5, kk = k & 31;
  return ((((((size_t)(nh * TC_NCHUNK + c) * 2 + part) * 4 + (kk >> 3)) * 32 + (n >> 3)) * 8 + (n & 7)) * 8) + (kk & 7);
}

__device__ __forceinline__ void split_f16(float v, float scale, __half& hi, __half& lo) {
  const float s = v * scale;            // power of two: exact
  hi = __float2half_rn(s);
  lo = __float2half_rn(s - __half2float(hi));
}

__device__ __forceinline__ void store_split4(__half* img, size_t off_hi, size_t off_lo, const float4& v, float scale) {
  __half h[4], l[4];
  split_f16(v.x, scale, h[0], l[0]);
  split_f16(v.y, scale, h[1], l[1]);
  split_f16(v.z, scale, h[2], l[2]);
  split_f16(v.w, scale, h[3], l[3]);
  uint2 ph, pl;
  ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
  ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
  pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
  pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
  *reinterpret_cast<uint2*>(img + off_hi) = ph;
  *reinterpret_cast<uint2*>(img + off_lo) = pl;
}

// ---- weight images (once per optimizer step) ------------------------------------------------
__global__ void pack_tc_kernel(ic3_policy_params p, __half* __restrict__ img, float* __restrict__ bias_cat) {
  const int H = TC_H;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;      // (col, k)
  if (idx < 4 * H * TC_K) {
    const int col = idx / TC_K, k = idx - col * TC_K;
    const int u = col >> 2, g = col & 3, row = g * H + u;       // column 4*u+gate <- LSTMCell row g*H+u
    double w;
    if (k < H) {
      w = p.w_ih[(size_t)row * H + k];
    } else if (k < 2 * H) {                                     // (W_ih . C)[row][k-H]
      double acc = 0.0;
      for (int m = 0; m < H; ++m) acc += (double)p.w_ih[(size_t)row * H + m] * (double)p.c_w[(size_t)m * H + (k - H)];
      w = acc;
    } else {
      w = p.w_hh[(size_t)row * H + (k - 2 * H)];
    }
    __half hi, lo;
    split_f16((float)w, SCALE_B, hi, lo);
    const int nh = col >> 8, n = col & 255;
    img[b_img_off(nh, k, n, 0)] = hi;
    img[b_img_off(nh, k, n, 1)] = lo;
  }
  if (idx < 4 * H) {
    const int u = idx >> 2, g = idx & 3, row = g * H + u;
    double acc = (double)p.b_ih[row] + (double)p.b_hh[row];
    for (int m = 0; m < H; ++m) acc += (double)p.w_ih[(size_t)row * H + m] * (double)p.c_b[m];
    bias_cat[idx] = (float)acc;
  }
}

// ---- operand A image: x | S | h (every step) ---------------------------------------------------
// One CTA per 128-row tile.  A warp item = 8 rows x 4 float4 columns, so every store instruction
// writes two complete 128-byte core matrices.  Environments may straddle tiles: the gated sum of
// comm.py:181-205 reads the other agents' rows straight from global memory (L1/L2 hits).
__global__ void __launch_bounds__(256) prep_kernel(ic3_policy_cfg cfg, ic3_policy_io io, __half* __restrict__ img) {
  __shared__ float s_gate[TC_M + 64];
  __shared__ float s_den[TC_M + 64];
  const int N = cfg.N;
  const long R = (long)cfg.B * N;
  const int tile = blockIdx.x;
  const long row0 = (long)tile * TC_M;
  for (int w = threadIdx.x; w < TC_M + 64; w += blockDim.x) {
    const long row = row0 - 32 + w;
    float g = 0.f, den = 1.f;
    if (row >= 0 && row < R) {
      const int e = (int)(row / N), i = (int)(row - (long)e * N);
      const bool fr = io.fresh && io.fresh[e];
      int n_alive = N, al = 1;
      if (io.alive && !fr) {                       // comm.py:102-104
        n_alive = 0;
        for (int j = 0; j < N; ++j) n_alive += io.alive[(size_t)e * N + j] != 0;
        al = io.alive[(size_t)e * N + i] != 0;
      }
      int cm = 1;
      if (cfg.hard_attn) cm = fr ? 0 : (io.comm_action[(size_t)e * N + i] != 0);   // comm.py:171-175
      g = (float)(al * cm);
      if (cfg.comm_avg && n_alive > 1) den = (float)(n_alive - 1);                  // comm.py:194-196
    }
    s_gate[w] = g;
    s_den[w] = den;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int item = warp; item < 128; item += 8) {
    const int rc = item & 15, qg = item >> 4;
    const int r = rc * 8 + (lane & 7), q = qg * 4 + (lane >> 3);
    const long row = row0 + r;
    float4 xv = zero4, hv = zero4, sv = zero4;
    if (row < R) {
      const int e = (int)(row / N);
      const bool fr = io.fresh && io.fresh[e];
      xv = __ldg(reinterpret_cast<const float4*>(io.x + (size_t)row * TC_H) + q);
      if (!fr) hv = __ldg(reinterpret_cast<const float4*>(io.h + (size_t)row * TC_H) + q);
      // episode start: every agent of the env has h = 0 (trainer.py:50-51), so S = 0 whatever the gates
      if (!fr && !cfg.comm_mask_zero && s_gate[r + 32] != 0.f) {
        const long base = (long)e * N;
        for (int j = 0; j < N; ++j) {
          const long rj = base + j;
          if (rj != row && s_gate[(int)(rj - row0) + 32] != 0.f) {
            const float4 o = __ldg(reinterpret_cast<const float4*>(io.h + (size_t)rj * TC_H) + q);
            sv.x += o.x; sv.y += o.y; sv.z += o.z; sv.w += o.w;
          }
        }
        const float d = s_den[r + 32];
        sv.x /= d; sv.y /= d; sv.z /= d; sv.w /= d;
      }
    }
    const int k = 4 * q;
    store_split4(img, a_img_off(tile, k, r, 0), a_img_off(tile, k, r, 1), xv, SCALE_A);
    store_split4(img, a_img_off(tile, TC_H + k, r, 0), a_img_off(tile, TC_H + k, r, 1), sv, SCALE_A);
    store_split4(img, a_img_off(tile, 2 * TC_H + k, r, 0), a_img_off(tile, 2 * TC_H + k, r, 1), hv, SCALE_A);
  }
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a mis-programmed pipeline must never hang the GPU; it raises the flag instead.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int32_t* err) {
  for (uint32_t spin = 0; spin < WATCHDOG_SPINS; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return true;
  }
  if (err) atomicOr(err, 0x100);
  return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, kind::f16 (fp16 inputs, fp32 accumulate), one CTA
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, no swizzle: core matrices of 8 rows x 16 B; LBO = byte distance between the two
// K-adjacent core matrices of one MMA, SBO = distance between 8-row groups (both >> 4).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;                 // descriptor version 1 (sm_100)
  return d;                               // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float sigmoid_(float v) { return 1.f / (1.f + expf(-v)); }

// ---- the tensor-core kernel -----------------------------------------------------------------------
__global__ void __launch_bounds__(TC_THREADS) lstm_tc_kernel(ic3_policy_cfg cfg, ic3_policy_io io,
                                                             const __half* __restrict__ a_img,
                                                             const __half* __restrict__ b_img,
                                                             const float* __restrict__ bias_cat) {
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + NSTAGE * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + NSTAGE), bar_tmem = smem_u32(bars + 2 * NSTAGE);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x >> 1, nh = blockIdx.x & 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {   // TMEM: 256 fp32 columns x 128 lanes for the accumulator tile
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TC_NH)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4 && lane == 0) {
    // ===== producer: two bulk copies per stage =====
    const unsigned char* a_src = reinterpret_cast<const unsigned char*>(a_img) + (size_t)tile * TC_NCHUNK * A_CHUNK_BYTES;
    const unsigned char* b_src = reinterpret_cast<const unsigned char*>(b_img) + (size_t)nh * TC_NCHUNK * B_CHUNK_BYTES;
    for (int c = 0; c < TC_NCHUNK; ++c) {
      const int s = c & (NSTAGE - 1);
      if (!mbar_wait(bar_empty + 8 * s, ((c / NSTAGE) & 1) ^ 1, io.err)) break;
      const uint32_t dst = smem_u32(smem + s * STAGE_BYTES);
      mbar_expect_tx(bar_full + 8 * s, STAGE_BYTES);
      bulk_g2s(dst, a_src + (size_t)c * A_CHUNK_BYTES, A_CHUNK_BYTES, bar_full + 8 * s);
      bulk_g2s(dst + A_CHUNK_BYTES, b_src + (size_t)c * B_CHUNK_BYTES, B_CHUNK_BYTES, bar_full + 8 * s);
    }
  } else if (warp == 5 && lane == 0) {
    // ===== MMA issuer: 3 x (128 x 256 x 16) per k-step, 72 instructions per tile =====
    // instruction descriptor: D = f32 (bits 4-5 = 1), A = B = f16 (0), K-major both, N >> 3 at 17, M >> 4 at 24
    const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_NH >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);
    bool ok = true;
    for (int c = 0; c < TC_NCHUNK && ok; ++c) {
      const int s = c & (NSTAGE - 1);
      ok = mbar_wait(bar_full + 8 * s, (c / NSTAGE) & 1, io.err);
      tc_fence_after();
      const uint32_t a0 = smem_u32(smem + s * STAGE_BYTES), b0 = a0 + A_CHUNK_BYTES;
#pragma unroll
      for (int ks = 0; ks < TC_KC / 16; ++ks) {
        // A: kcore block = 16 rcores x 128 B = 2048 B; B: 32 ncores x 128 B = 4096 B; lo half follows hi half
        const uint64_t da_hi = make_desc(a0 + ks * 4096, 2048, 128);
        const uint64_t da_lo = make_desc(a0 + A_CHUNK_BYTES / 2 + ks * 4096, 2048, 128);
        const uint64_t db_hi = make_desc(b0 + ks * 8192, 4096, 128);
        const uint64_t db_lo = make_desc(b0 + B_CHUNK_BYTES / 2 + ks * 8192, 4096, 128);
        tc_mma_f16(tmem_base, da_hi, db_hi, idesc, (c | ks) != 0);
        tc_mma_f16(tmem_base, da_lo, db_hi, idesc, 1);
        tc_mma_f16(tmem_base, da_hi, db_lo, idesc, 1);
      }
      tc_commit(bar_empty + 8 * s);      // frees the stage when these MMAs have read it
    }
    tc_commit(bar_tmem);                 // accumulator complete
  } else if (warp < 4) {
    // ===== epilogue: thread = one row of the tile (TMEM lane), 4 hidden units per tcgen05.ld =====
    const bool ok = mbar_wait(bar_tmem, 0, io.err);
    tc_fence_after();
    const int r = warp * 32 + lane;
    const long row = (long)tile * TC_M + r;
    const long R = (long)cfg.B * cfg.N;
    const bool valid = ok && row < R;
    bool fr = false;
    if (valid && io.fresh) fr = io.fresh[row / cfg.N] != 0;
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int cg = 0; cg < TC_NH / 16; ++cg) {
      uint32_t v[16];
      tmem_ld16(tlane + cg * 16, v);
      if (valid) {
        const int u0 = nh * (TC_NH / 4) + cg * 4;
        float4 cold = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!fr) cold = *reinterpret_cast<const float4*>(io.c + (size_t)row * TC_H + u0);
        float cn[4], hn[4];
        const float co[4] = {cold.x, cold.y, cold.z, cold.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias_cat) + (u0 + j));
          const float gi = sigmoid_(fmaf(__uint_as_float(v[4 * j + 0]), INV_SCALE, b.x));
          const float gf = sigmoid_(fmaf(__uint_as_float(v[4 * j + 1]), INV_SCALE, b.y));
          const float gg = tanhf(fmaf(__uint_as_float(v[4 * j + 2]), INV_SCALE, b.z));
          const float go = sigmoid_(fmaf(__uint_as_float(v[4 * j + 3]), INV_SCALE, b.w));
          cn[j] = gf * co[j] + gi * gg;
          hn[j] = go * tanhf(cn[j]);
        }
        *reinterpret_cast<float4*>(io.c_out + (size_t)row * TC_H + u0) = make_float4(cn[0], cn[1], cn[2], cn[3]);
        *reinterpret_cast<float4*>(io.h_out + (size_t)row * TC_H + u0) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TC_NH) : "memory");
  }
}

// ---- heads + sampling from h' (comm.py:228-239, action_utils.py:32-36) -----------------------------
__global__ void __launch_bounds__(256) heads_kernel(ic3_policy_cfg cfg, ic3_policy_packed w, ic3_policy_io io) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * 8 + warp;
  if (row >= (long)cfg.B * cfg.N) return;
  float hv[TC_H / 32];
#pragma unroll
  for (int m = 0; m < TC_H / 32; ++m) hv[m] = io.h_out[(size_t)row * TC_H + lane + 32 * m];
  const int e = (int)(row / cfg.N), i = (int)(row - (long)e * cfg.N);
  heads_for_row<TC_H>(cfg, w.head_w, w.head_b, hv, (size_t)row, e, i, lane, io.tick, io.draws, io.value, io.logp,
                      io.action);
}

}  // namespace

uint64_t ic3_tc_workspace_bytes(const ic3_policy_cfg* cfg) {
  if (!cfg || cfg->H != TC_H) return 0;
  const long R = (long)cfg->B * cfg->N;
  const long ntiles = (R + TC_M - 1) / TC_M;
  return (uint64_t)ntiles * TC_NCHUNK * A_CHUNK_BYTES;
}

int ic3_tc_pack(const ic3_policy_cfg* cfg, const ic3_policy_params* p, const ic3_policy_packed* out, cudaStream_t s) {
  if (cfg->H != TC_H) return IC3_E_UNSUPPORTED;
  const int total = 4 * TC_H * TC_K;
  pack_tc_kernel<<<(total + 255) / 256, 256, 0, s>>>(*p, reinterpret_cast<__half*>(out->lstm_img), out->bias_cat);
  IC3_LAUNCH_CHECK();
  return IC3_OK;
}

int ic3_tc_policy_step(const ic3_policy_cfg* cfg, const ic3_policy_packed* w, const ic3_policy_io* io, cudaStream_t s) {
  if (cfg->H != TC_H) return IC3_E_UNSUPPORTED;
  if (!io->workspace || !w->lstm_img || !w->bias_cat) return IC3_E_NULL;
  const long R = (long)cfg->B * cfg->N;
  const int ntiles = (int)((R + TC_M - 1) / TC_M);
  __half* img = reinterpret_cast<__half*>(io->workspace);
  prep_kernel<<<ntiles, 256, 0, s>>>(*cfg, *io, img);
  IC3_LAUNCH_CHECK();
  const size_t smem = NSTAGE * STAGE_BYTES + 128;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    configured = true;
  }
  lstm_tc_kernel<<<2 * ntiles, TC_THREADS, smem, s>>>(*cfg, *io, img, reinterpret_cast<const __half*>(w->lstm_img),
                                                      w->bias_cat);
  IC3_LAUNCH_CHECK();
  heads_kernel<<<(int)((R + 7) / 8), 256, 0, s>>>(*cfg, *w, *io);
  IC3_LAUNCH_CHECK();
  return IC3_OK;
}
