"""
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of rollout-time inference:
PPOPolicy.get_rollout_actions (reference policies/ppo_policy.py:729-794) and get_critic_values (:1057-1071), with the
sampling protocol of the torch CPU generator that the reference runs on:

  Gaussian     Normal(mean, std).sample() == torch.normal(mean, std): out.normal_(0, 1) of the broadcast shape, then
               out.mul_(std).add_(mean)  (aten normal_out_impl)       -> draw_gaussian_noise()
               refine_sample: tanh, then ((s + 1) / 2) * (max - min) + min when the range is not [-1, 1]
               (networks/distributions.py:560-610, 633-655); log-prob :518-558
  Categorical  Categorical(probs).sample() == multinomial(probs, 1, True): q.exponential_(1); argmax(probs / q)
               (aten multinomial fast path for one sample)            -> draw_categorical_noise()
               log-prob :223-249 through probs_to_logits

Pinned against tests/golden/act_*.npz (outputs of the unmodified reference).
"""
import numpy as np
import torch

from .update import OracleMLP, categorical_from_probs, gaussian_std, gaussian_tanh_log_prob


def draw_gaussian_noise(n_rows, act_dim):
    """The N(0,1) draws Normal(mean[n, d], std[d]).sample() consumes from the global CPU generator."""
    return torch.empty(n_rows, act_dim, dtype=torch.float32).normal_(0, 1)


def draw_categorical_noise(n_rows, n_cat):
    """The Exp(1) draws Categorical(probs[n, c]).sample() consumes from the global CPU generator."""
    return torch.empty(n_rows, n_cat, dtype=torch.float32).exponential_(1)


def rollout_actions(actor_params, activation, obs, discrete, dist_min=-1.0, dist_max=1.0, min_std=0.01, noise=None):
    """Returns (raw_action, action, log_prob) as numpy arrays with the reference's shapes."""
    log_std = actor_params.get("distribution.log_std")
    actor = OracleMLP({k: v for k, v in actor_params.items() if k != "distribution.log_std"}, activation)
    x = torch.tensor(np.asarray(obs), dtype=torch.float32)
    with torch.no_grad():
        pred = actor.forward(x)
        if discrete:
            probs = torch.softmax(pred, dim=-1)                               # distributions.py:1045
            p, logits = categorical_from_probs(probs)
            q = draw_categorical_noise(*p.shape) if noise is None else noise
            sample = torch.argmax(p / q, dim=-1)
            lp = logits.gather(-1, sample.unsqueeze(-1))                      # [n, 1]
            a = sample.unsqueeze(-1)
            return a.numpy(), a.numpy(), lp.numpy()
        std = gaussian_std(torch.tensor(np.asarray(log_std), dtype=torch.float32), min_std)
        eps = draw_gaussian_noise(*pred.shape) if noise is None else noise
        raw = eps.mul(std).add(pred)
        act = torch.tanh(raw)
        lo = torch.as_tensor(np.asarray(dist_min, dtype=np.float32))
        hi = torch.as_tensor(np.asarray(dist_max, dtype=np.float32))
        if bool((lo != -1.0).any()) or bool((hi != 1.0).any()):
            act = ((act + 1.0) / 2.0) * (hi - lo) + lo
        lp = gaussian_tanh_log_prob(pred, std, raw)
        return raw.numpy(), act.numpy(), lp.numpy()


def critic_values(critic_params, activation, critic_obs):
    critic = OracleMLP(critic_params, activation)
    with torch.no_grad():
        return critic.forward(torch.tensor(np.asarray(critic_obs), dtype=torch.float32)).numpy()
