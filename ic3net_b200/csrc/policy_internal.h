// Internal (non-ABI) entry points of the tcgen05 policy path, called from policy.cu.
#pragma once
#include <cuda_runtime.h>

#include "../../include/ic3net_b200.h"

uint64_t ic3_tc_workspace_bytes(const ic3_policy_cfg* cfg);
int ic3_tc_pack(const ic3_policy_cfg* cfg, const ic3_policy_params* p, const ic3_policy_packed* out, cudaStream_t s);
int ic3_tc_policy_step(const ic3_policy_cfg* cfg, const ic3_policy_packed* w, const ic3_policy_io* io, cudaStream_t s);
// layout hint of ic3_policy_cfg vs the environment (IC3_OK when no hint is given)
int ic3_pp_layout_check(const ic3_pp_cfg* env, const ic3_policy_cfg* cfg);
int ic3_tj_layout_check(const ic3_tj_cfg* env, const ic3_policy_cfg* cfg);
