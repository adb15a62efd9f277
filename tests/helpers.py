"""Shared builders for the GPU parity tests (the same golden fixtures that pin the oracle)."""
import numpy as np
import torch

from ppo_and_friends_b200.spaces import Box, Discrete

ACT_MODULES = {"leaky_relu": torch.nn.LeakyReLU, "tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}


def make_policy(ro, act="leaky_relu", actor_hidden=32, critic_hidden=32, depth=3, dist_range=1.0, device="cuda",
                **policy_kw):
    from ppo_and_friends_b200.policies.ppo_policy import PPOPolicy
    if ro.n_discrete:
        action_space = Discrete(ro.n_discrete)
    else:
        action_space = Box(-dist_range, dist_range, (ro.act_dim,))
    pol = PPOPolicy("pol", action_space, Box(-np.inf, np.inf, (ro.obs_dim,)), Box(-np.inf, np.inf, (ro.critic_obs_dim,)),
                    envs_per_proc=ro.E,
                    actor_kw_args=dict(activation=ACT_MODULES[act](), hidden_size=actor_hidden, hidden_depth=depth),
                    critic_kw_args=dict(activation=ACT_MODULES[act](), hidden_size=critic_hidden, hidden_depth=depth),
                    **policy_kw)
    for a in ro.agents:
        pol.register_agent(a)
    # the reference's agent order is a set-union order; keep the rollout's order for deterministic columns
    pol.agent_ids = np.array(ro.agents)
    pol.finalize({"global status": {"iteration": 0, "timesteps": 0}}, torch.device(device))
    return pol


def policy_kwargs_from_golden(g, prefix=""):
    kw = {}
    if (prefix + "have_bootstrap_clip") in g:
        if bool(g[prefix + "have_bootstrap_clip"]):
            bc = g[prefix + "bootstrap_clip"]
            kw["bootstrap_clip"] = (float(bc[0]), float(bc[1]))
        else:
            kw["bootstrap_clip"] = None
    for k in ("use_gae", "dynamic_bs_clip"):
        if (prefix + k) in g:
            kw[k] = bool(g[prefix + k])
    for k in ("gamma", "lambd"):
        if (prefix + k) in g:
            kw[k] = float(g[prefix + k])
    return kw


def run_device_rollout(pol, ro, tensor_bootstrap=False):
    from ppo_and_friends_b200.synthetic import replay_rollout
    pol.initialize_dataset()
    pol.initialize_episodes(ro.E, {"global status": {"iteration": 0, "timesteps": 0}})
    wrap = (lambda x: torch.tensor(x, requires_grad=True)) if tensor_bootstrap else None
    replay_rollout(lambda a: pol, ro, to_bootstrap=wrap)
    pol.finalize_dataset()
    return pol.dataset


def rel_err(a, b, floor):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))) if a.size else 0.0
