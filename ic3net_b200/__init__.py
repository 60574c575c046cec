"""ic3net_b200 -- B200-native rollout hot path of IC3Net behind the reference's Python surface.

Host side (this package) mirrors the reference modules of the path:
  predator_prey_env / traffic_junction_env / traffic_helper   <- ic3net_envs/*
  env_wrappers.GymWrapper, data.init                         <- env_wrappers.py, data.py
  comm.CommNetMLP, action_utils                              <- comm.py, action_utils.py
  trainer.Trainer, multi_gpu.MultiGPUTrainer                 <- trainer.py, multi_processing.py
Device side: csrc/*.cu (hand-written sm_100a kernels) behind the C ABI of include/ic3net_b200.h,
loaded with ctypes (ic3net_b200/_lib.py).  No CPU fallback exists.
"""
__version__ = "0.1"
