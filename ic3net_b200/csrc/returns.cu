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

// Batch statistics of Trainer.run_batch (trainer.py:73-75,86-88,109-110,124-125,235) summed over the env slots of this
// GPU into ONE float64 vector, so the host needs a single device->host copy per update (and the data-parallel
// trainer can all-reduce the vector before it ever reaches the host):
//   out = [num_episodes, num_steps, success, err flags, reward[N], comm_action[N]]
__global__ void stat_reduce_kernel(int B, int N, const int32_t* __restrict__ episodes, const int32_t* __restrict__ steps,
                                   const int32_t* __restrict__ success, const int32_t* __restrict__ err,
                                   const float* __restrict__ reward, const float* __restrict__ comm,
                                   double* __restrict__ out) {
  __shared__ double s_acc[4 + 2 * IC3_MAX_AGENTS];
  const int nacc = 4 + 2 * N;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) s_acc[i] = 0.0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * (blockDim.x >> 5);
  double e = 0.0, st = 0.0, su = 0.0, rw = 0.0, cm = 0.0;
  for (int b = warp_global; b < B; b += nwarps) {       // warp per env slot, lane = agent
    if (lane == 0) {
      e += (double)episodes[b];
      st += (double)steps[b];
      su += (double)success[b];
    }
    if (lane < N) {
      rw += (double)reward[(size_t)b * N + lane];
      if (comm) cm += (double)comm[(size_t)b * N + lane];
    }
  }
  if (lane == 0) {
    atomicAdd(&s_acc[0], e);
    atomicAdd(&s_acc[1], st);
    atomicAdd(&s_acc[2], su);
  }
  if (lane < N) {
    atomicAdd(&s_acc[4 + lane], rw);
    atomicAdd(&s_acc[4 + N + lane], cm);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nacc; i += blockDim.x)
    if (i != 3 && s_acc[i] != 0.0) atomicAdd(out + i, s_acc[i]);
  if (blockIdx.x == 0 && threadIdx.x == 0 && err) out[3] = (double)err[0];
}

}  // namespace

extern "C" int ic3_stat_reduce(int32_t B, int32_t N, const int32_t* stat_episodes, const int32_t* stat_steps,
                               const int32_t* stat_success, const int32_t* err, const float* stat_reward,
                               const float* stat_comm, double* out, void* stream) {
  if (!stat_episodes || !stat_steps || !stat_success || !stat_reward || !out) return IC3_E_NULL;
  if (B <= 0 || N <= 0 || N > IC3_MAX_AGENTS) return IC3_E_RANGE;
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(double) * (4 + 2 * N), s);
  if (e != cudaSuccess) return (int)e;
  const int grid = B < 8 * 64 ? (B + 7) / 8 : 64;
  stat_reduce_kernel<<<grid, 256, 0, s>>>(B, N, stat_episodes, stat_steps, stat_success, err, stat_reward, stat_comm, out);
  IC3_LAUNCH_CHECK();
  return IC3_OK;
}

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
