"""
CPU restatement (numpy, fp64 state) of the reference's normaliser / clipper wrapper stack
(environments/filter_wrappers.py).  TEST INFRASTRUCTURE ONLY.

  ObservationNormalizer._filter_*  :155-196  update(obs[agent]) then (obs - mean) / sqrt(var + eps)
  ObservationClipper               :617-671  np.clip
  RewardNormalizer.step            :393-447  per env, IN ORDER: running_reward[e] = running_reward[e] * gamma + r[e] and
                                             running_stats.update(running_reward) with the partially updated vector (Q9);
                                             running_reward[done] = 0; reward / sqrt(var + eps)
  RewardClipper.step               :690-719  np.clip
  RunningMeanStd.update            utils/stats.py:29-94 (oracle/stats.py)
"""
import numpy as np

from .stats import OracleRunningMeanStd


class OracleFilterStack:
    def __init__(self, agents, obs_dim, critic_dim, n_envs, gamma=0.99, eps=1e-8, obs_clip=(-10.0, 10.0),
                 reward_clip=(-10.0, 10.0)):
        self.agents = tuple(agents)
        self.actor = {a: OracleRunningMeanStd(shape=(obs_dim,)) for a in self.agents}
        self.critic = {a: OracleRunningMeanStd(shape=(critic_dim,)) for a in self.agents}
        self.reward = {a: OracleRunningMeanStd(shape=()) for a in self.agents}
        self.running_reward = {a: np.zeros(n_envs) for a in self.agents}
        self.gamma, self.eps, self.obs_clip, self.reward_clip = gamma, eps, obs_clip, reward_clip

    def _obs(self, table, obs):
        out = {}
        for a in obs:
            table[a].update(obs[a])
            x = (obs[a] - table[a].mean) / np.sqrt(table[a].variance + self.eps)
            out[a] = np.clip(x, *self.obs_clip)
        return out

    def filter_obs(self, obs, critic_obs):
        return self._obs(self.actor, obs), self._obs(self.critic, critic_obs)

    def filter_reward(self, reward, terminated, truncated):
        out = {}
        for a in reward:
            rr = self.running_reward[a]
            for e in range(rr.shape[0]):
                rr[e] = rr[e] * self.gamma + reward[a][e]
                self.reward[a].update(rr)
            rr[np.logical_or(terminated[a], truncated[a])] = 0.0
            out[a] = np.clip(reward[a] / np.sqrt(self.reward[a].variance + self.eps), *self.reward_clip)
        return out
