"""
ppo_and_friends_b200 — B200-native (sm_100a) implementation of the PPO-AF post-rollout update path:
GAE / reward-to-go segmented scan, RunningMeanStd normalisation, and the fused PPO minibatch update
of separate actor / critic MLPs, behind the reference's policy / dataset / trainer surface.

  csrc/            hand-written CUDA kernels + the C ABI (include/ppoaf_b200.h) -> libppoaf_b200.so
  _lib.py, ops.py  ctypes binding and torch-tensor front-ends (no CPU fallback)
  utils/, networks/, policies/, ppo.py   host-side mirror of the reference modules of the same name
  synthetic.py     synthetic rollouts + the rollout-replay driver (the caller of the path)
"""
__version__ = "0.1.0"
