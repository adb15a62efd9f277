"""
Synthetic rollouts and the rollout-replay driver.

The hot path this package implements starts at the calls `PPO.rollout` makes into
the policy (`add_episode_info` / `end_episodes` / `finalize_dataset`).  Environment
stepping stays on the host and is out of scope, so tests and benchmarks need a
stand-in *caller* that issues exactly the call sequence the reference trainer
issues.  `replay_rollout` is that caller: it restates the bookkeeping of the
reference rollout loop (reference ppo.py:1646-1938) over pre-generated arrays:

  * per step: `ep_ts += 1`, `episode_lengths += 1`, `total += E`      (ppo.py:1648-1654)
  * truncation beats termination when both are set                   (ppo.py:1446-1455)
  * `add_episode_info` once per agent, in agent order                 (ppo.py:1729-1752)
  * terminated envs are closed first, terminal=1, zero bootstraps     (ppo.py:1806-1819)
  * then "maxed" envs: `ep_ts >= max_ts_per_ep`, all envs at rollout end, plus
    truncated envs, minus terminated ones, `np.unique`-sorted; closed with
    terminal=0 and the FULL [E] next-value arrays (the policy indexes them by
    position, reference policies/ppo_policy.py:684-689)               (ppo.py:1866-1938)

The same driver feeds the unmodified reference (when generating golden vectors),
the CPU oracle, and the CUDA path, so all three see identical inputs.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np


@dataclass
class SyntheticRollout:
    """Time-major synthetic rollout: every array is [T, E, ...] per agent."""
    T: int
    E: int
    agents: List[str]
    obs_dim: int
    critic_obs_dim: int
    act_dim: int                 # stored action width (1 for Discrete)
    n_discrete: int              # 0 => continuous (Gaussian-tanh head)
    max_ts_per_ep: int
    obs: Dict[str, np.ndarray] = field(default_factory=dict)
    next_obs: Dict[str, np.ndarray] = field(default_factory=dict)
    critic_obs: Dict[str, np.ndarray] = field(default_factory=dict)
    raw_actions: Dict[str, np.ndarray] = field(default_factory=dict)
    actions: Dict[str, np.ndarray] = field(default_factory=dict)
    values: Dict[str, np.ndarray] = field(default_factory=dict)        # [T, E]
    log_probs: Dict[str, np.ndarray] = field(default_factory=dict)     # [T, E]
    rewards: Dict[str, np.ndarray] = field(default_factory=dict)       # [T, E]
    next_values: Dict[str, np.ndarray] = field(default_factory=dict)   # [T, E] critic(next obs)
    terminated: Optional[np.ndarray] = None                            # [T, E] bool (env level)
    truncated: Optional[np.ndarray] = None                             # [T, E] bool (env level)

    @property
    def timesteps(self):
        return self.T * self.E


def make_rollout(seed, T, E, agents=("agent0",), obs_dim=8, critic_obs_dim=None,
                 act_dim=2, n_discrete=0, max_ts_per_ep=32, p_term=0.01, p_trunc=0.01,
                 obs_scale=True, shared_critic_obs=False):
    """
    Seeded synthetic rollout with the statistics SURVEY.md §8(d) names: rewards and
    values ~ N(0,1) fp32, `terminated`/`truncated` ~ Bernoulli (mutually exclusive
    draws), observations ~ N(mu_d, sigma_d^2) with per-feature mu_d in [-3,3] and
    sigma_d in [0.1,10] (or unit normal when obs_scale=False).  values/log_probs are
    random here; callers that need "old" values/log-probs consistent with a network
    overwrite them (see `fill_policy_outputs` in the tests and bench).
    """
    rng = np.random.default_rng(seed)
    agents = list(agents)
    Dc = obs_dim if critic_obs_dim is None else critic_obs_dim
    ro = SyntheticRollout(T=T, E=E, agents=agents, obs_dim=obs_dim, critic_obs_dim=Dc,
                          act_dim=(1 if n_discrete else act_dim), n_discrete=n_discrete,
                          max_ts_per_ep=max_ts_per_ep)
    u = rng.random((T, E))
    ro.terminated = u < p_term
    ro.truncated = (u >= p_term) & (u < p_term + p_trunc)
    if obs_scale:
        mu = rng.uniform(-3.0, 3.0, size=obs_dim).astype(np.float32)
        sd = rng.uniform(0.1, 10.0, size=obs_dim).astype(np.float32)
    else:
        mu = np.zeros(obs_dim, np.float32)
        sd = np.ones(obs_dim, np.float32)
    for a in agents:
        ro.obs[a] = (rng.standard_normal((T, E, obs_dim), dtype=np.float32) * sd + mu)
        ro.next_obs[a] = (rng.standard_normal((T, E, obs_dim), dtype=np.float32) * sd + mu)
    for a in agents:
        if shared_critic_obs:
            # MAPPO "policy" critic view: concat of all agents' obs in agent order
            # (reference environments/ppo_env_wrappers.py:701-730).
            ro.critic_obs[a] = np.concatenate([ro.obs[b] for b in agents], axis=-1)
            assert ro.critic_obs[a].shape[-1] == Dc
        elif Dc == obs_dim:
            ro.critic_obs[a] = ro.obs[a].copy()
        else:
            ro.critic_obs[a] = rng.standard_normal((T, E, Dc), dtype=np.float32)
        if n_discrete:
            act = rng.integers(0, n_discrete, size=(T, E, 1)).astype(np.int64)
            ro.raw_actions[a] = act
            ro.actions[a] = act.copy()
        else:
            raw = rng.standard_normal((T, E, act_dim), dtype=np.float32)
            ro.raw_actions[a] = raw
            ro.actions[a] = np.tanh(raw)
        ro.values[a] = rng.standard_normal((T, E), dtype=np.float32)
        ro.next_values[a] = rng.standard_normal((T, E), dtype=np.float32)
        ro.log_probs[a] = (-1.0 - rng.random((T, E), dtype=np.float32))
        ro.rewards[a] = rng.standard_normal((T, E), dtype=np.float32)
    return ro


def replay_rollout(policy_for_agent, rollout, to_bootstrap=None):
    """
    Drive `add_episode_info` / `end_episodes` exactly as the reference trainer does.

    policy_for_agent: callable agent_id -> policy object (reference PPOPolicy, the
        oracle's policy stand-in, or this package's PPOPolicy).
    to_bootstrap: optional callable wrapping the [E] next-value array handed to
        `end_episodes` (the reference hands a torch tensor for non-terminal ends and
        numpy zeros for terminal ends); default passes numpy.

    Returns the list of (step, kind, env_idxs) closure events for inspection.
    """
    T, E = rollout.T, rollout.E
    ts_per_rollout = T * E
    episode_lengths = np.zeros(E, dtype=np.int32)
    ep_ts = np.zeros(E, dtype=np.int32)
    total = 0
    events = []
    wrap = to_bootstrap if to_bootstrap is not None else (lambda x: x)

    for t in range(T):
        ep_ts += 1
        total += E
        episode_lengths += 1

        truncated = rollout.truncated[t]
        terminated = rollout.terminated[t].copy()
        where_truncated = np.where(truncated)[0]
        if terminated[where_truncated].any():
            terminated[where_truncated] = False
        have_truncated = where_truncated.size > 0

        where_term = np.where(terminated)[0]
        where_not_term = np.where(~terminated)[0]
        term_count = where_term.size

        for a in rollout.agents:
            policy_for_agent(a).add_episode_info(
                agent_id=a,
                critic_observations=rollout.critic_obs[a][t],
                observations=rollout.obs[a][t],
                next_observations=rollout.next_obs[a][t],
                raw_actions=rollout.raw_actions[a][t],
                actions=rollout.actions[a][t],
                values=rollout.values[a][t],
                log_probs=rollout.log_probs[a][t],
                rewards=rollout.rewards[a][t],
                where_done=where_term)

        if term_count > 0:
            for a in rollout.agents:
                policy_for_agent(a).end_episodes(
                    agent_id=a,
                    env_idxs=where_term,
                    episode_lengths=episode_lengths,
                    terminal=np.ones(term_count).astype(bool),
                    ending_values=np.zeros(term_count),
                    ending_rewards=np.zeros(term_count))
            events.append((t, "terminal", where_term.copy()))
            episode_lengths[where_term] = 0
            ep_ts[where_term] = 0

        ep_max_reached = bool((ep_ts == rollout.max_ts_per_ep).any() and where_not_term.size > 0)

        if ep_max_reached or total >= ts_per_rollout or have_truncated:
            if total >= ts_per_rollout:
                where_maxed = np.arange(E)
            else:
                where_maxed = np.where(ep_ts >= rollout.max_ts_per_ep)[0]
            where_maxed = np.setdiff1d(where_maxed, where_term)
            where_maxed = np.concatenate((where_maxed, where_truncated))
            where_maxed = np.unique(where_maxed)
            maxed_count = where_maxed.size
            if maxed_count > 0:
                for a in rollout.agents:
                    nv = wrap(rollout.next_values[a][t])
                    policy_for_agent(a).end_episodes(
                        agent_id=a,
                        env_idxs=where_maxed,
                        episode_lengths=episode_lengths,
                        terminal=np.zeros(maxed_count).astype(bool),
                        ending_values=nv,
                        ending_rewards=nv)
                events.append((t, "maxed", where_maxed.copy()))
            ep_ts[where_maxed] = 0
    return events
