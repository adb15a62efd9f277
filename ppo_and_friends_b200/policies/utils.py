"""
`generate_policy` for the B200 path: the hook the reference trainer calls for every entry of `policy_settings`
(reference policies/utils.py:11-66, ppo.py:336-345).  Same signature and error behaviour; it builds the device-resident
PPOPolicy for the combinations the CUDA path covers and refuses the rest loudly (there is no CPU fallback).
"""
from .ppo_policy import PPOPolicy
from ..utils.mpi_utils import abort


def generate_policy(policy_name, policy_class, actor_observation_space, critic_observation_space, action_space,
                    test_mode, envs_per_proc, **kw_args):
    name = getattr(policy_class, "__name__", None)
    if policy_class is not None and name != "PPOPolicy":
        abort("ERROR: policy_class is of unsupported type, {}. Supported types on the B200 path are "
              "[PPOPolicy, None] (MATPolicy is outside the accelerated path).".format(policy_class))
    return PPOPolicy(name=policy_name, action_space=action_space, actor_observation_space=actor_observation_space,
                     critic_observation_space=critic_observation_space, test_mode=test_mode, envs_per_proc=envs_per_proc,
                     **kw_args)
