// Reverse return scan of Trainer.compute_grad (reference trainer.py:160-173) for B env slots in
// parallel: one warp per env slot, lane = agent, sequential over the T lock-step records.
//   coop[i]  = r[i] + gamma * coop[i+1]  * episode_mask[i]
//   ncoop[i] = r[i] + gamma * ncoop[i+1] * episode_mask[i] * episode_mini_mask[i]
//   R[i]     = mean_ratio * mean_agents(coop[i]) + (1 - mean_ratio) * ncoop[i]
// float64 accumulators like the reference (default tensor type double, main.py:20).
#include "ic3_common.cuh"

namespace {

__global__ void returns_scan_kernel(int T, int B, int N, double gamma, double mean_ratio,
                                    const float* __restrict__ reward, const uint8_t* __restrict__ emask,
                                    const uint8_t* __restrict__ mini, float* __restrict__ returns) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  double coop = 0.0, ncoop = 0.0;
  for (int t = T - 1; t >= 0; --t) {
    const size_t i = ((size_t)t * B + b) * N + lane;
    double r = 0.0, mi = 1.0;
    if (lane < N) {
      r = (double)reward[i];
      mi = (double)mini[i];
    }
    const double em = (double)emask[(size_t)t * B + b];
    coop = r + gamma * coop * em;
    ncoop = r + gamma * ncoop * em * mi;
    double s = lane < N ? coop : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(IC3_FULL_MASK, s, o);
    if (lane < N) returns[i] = (float)(mean_ratio * (s / (double)N) + (1.0 - mean_ratio) * ncoop);
  }
}

}  // namespace

extern "C" int ic3_returns_scan(int32_t T, int32_t B, int32_t N, float gamma, float mean_ratio, const float* reward,
                                const uint8_t* episode_mask, const uint8_t* mini_mask, float* returns, void* stream) {
  if (!reward || !episode_mask || !mini_mask || !returns) return IC3_E_NULL;
  if (T <= 0 || B <= 0 || N <= 0 || N > IC3_MAX_AGENTS) return IC3_E_RANGE;
  const int wpb = 4;
  returns_scan_kernel<<<(B + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(T, B, N, (double)gamma,
                                                                                   (double)mean_ratio, reward,
                                                                                   episode_mask, mini_mask, returns);
  IC3_LAUNCH_CHECK();
  return IC3_OK;
}
