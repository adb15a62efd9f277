"""
The normaliser / clipper wrapper stack with device-resident state (reference environments/filter_wrappers.py:
ObservationNormalizer :113-340, RewardNormalizer :343-476, GenericClipper :548-616, ObservationClipper :617-671,
RewardClipper :674-719) — SURVEY.md §8f row 2.

Same class names, constructor arguments and `step` / `reset` / `save_info` / `load_info` behaviour as the reference
wrappers, over the same duck-typed environment interface (`step(action) -> (obs, critic_obs, reward, terminated,
truncated, info)`, `reset() -> (obs, critic_obs)`, dictionaries keyed by agent id, arrays of shape [E, ...]).  The running
statistics live on the GPU (utils/stats.py), every per-step array is a CUDA tensor once it has crossed the bus, and the
arithmetic is libppoaf_b200.so: Welford / Chan batch moments + merge, normalise + clip in one pass, and the reward
normaliser's sequential-in-time update as one small kernel per step.  Across ranks the (mean, M2, n) triples
are exchanged instead of the raw batches (reference utils/stats.py:47-50 all-gathers pickled arrays on every step).
"""
import ctypes as C
import os
import pickle

import numpy as np
import torch

from .. import ops
from .._lib import check, load, ptr, stream_ptr
from ..utils import mpi_utils
from ..utils.stats import RunningMeanStd


def _callable(v):
    return v if callable(v) else (lambda: v)


def _dev(x, device, dtype=torch.float32):
    t = x if torch.is_tensor(x) else torch.as_tensor(np.ascontiguousarray(x))
    return t.to(device=device, dtype=dtype, non_blocking=True).contiguous()


class IdentityWrapper(object):
    """The slice of the reference's IdentityWrapper (environments/ppo_env_wrappers.py:24-147) this stack relies on."""

    def __init__(self, env, test_mode=False, device=None, **kw_args):
        self.env = env
        self.test_mode = test_mode
        self.device = torch.device("cuda") if device is None else torch.device(device)
        self.observation_space = env.observation_space
        self.critic_observation_space = env.critic_observation_space
        self.action_space = getattr(env, "action_space", None)
        self.agent_ids = tuple(getattr(env, "agent_ids", tuple(self.observation_space.keys())))
        self.finalized = False

    def get_batch_size(self):
        return self.env.get_batch_size()

    def step(self, action):
        return self.env.step(action)

    def reset(self):
        return self.env.reset()

    def finalize(self, status_dict):
        self.finalized = True
        if hasattr(self.env, "finalize"):
            self.env.finalize(status_dict)

    def save_info(self, path):
        if hasattr(self.env, "save_info"):
            self.env.save_info(path)

    def load_info(self, path):
        if hasattr(self.env, "load_info"):
            self.env.load_info(path)


class ObservationFilter(IdentityWrapper):
    """filter_wrappers.py:21-110."""

    def _apply_filters(self, local_obs, critic_obs):
        return self._filter_local_observation(local_obs), self._filter_critic_observation(critic_obs)

    def step(self, action):
        obs, critic_obs, reward, terminated, truncated, info = self.env.step(action)
        obs, critic_obs = self._apply_filters(obs, critic_obs)
        return obs, critic_obs, reward, terminated, truncated, info

    def reset(self):
        obs, critic_obs = self.env.reset()
        return self._apply_filters(obs, critic_obs)


class ObservationNormalizer(ObservationFilter):
    """Running-statistics observation normalisation (:113-340); `clip_range` fuses an ObservationClipper into the same
    pass (one read + one write of the observation instead of two of each)."""

    def __init__(self, env, update_stats=True, epsilon=1e-8, clip_range=None, **kw_args):
        super().__init__(env, **kw_args)
        self.actor_running_stats = {a: RunningMeanStd(shape=self.env.observation_space[a].shape, device=self.device)
                                    for a in self.env.observation_space}
        self.critic_running_stats = {a: RunningMeanStd(shape=self.env.critic_observation_space[a].shape, device=self.device)
                                     for a in self.env.critic_observation_space}
        self.update_stats = update_stats
        self.epsilon = epsilon
        self.clip_range = None if clip_range is None else (_callable(clip_range[0]), _callable(clip_range[1]))

    def _normalize(self, stats, x):
        x = _dev(x, self.device)
        lo, hi = (1.0, -1.0) if self.clip_range is None else (self.clip_range[0](), self.clip_range[1]())
        flat = x.reshape(-1, stats.dim)
        return ops.normalize_clip(flat, stats.state, stats.dim, self.epsilon, lo, hi).reshape(x.shape)

    def _filter(self, table, obs):
        out = {}
        for agent_id in obs:
            x = _dev(obs[agent_id], self.device)
            if self.update_stats:
                table[agent_id].update(x.reshape(-1, table[agent_id].dim))
            out[agent_id] = self._normalize(table[agent_id], x)
        return out

    def _filter_local_observation(self, obs):
        return self._filter(self.actor_running_stats, obs)

    def _filter_critic_observation(self, critic_obs):
        return self._filter(self.critic_running_stats, critic_obs)

    def local_normalize(self, obs):
        return {a: self._normalize(self.actor_running_stats[a], obs[a]) for a in obs}

    def critic_normalize(self, obs):
        return {a: self._normalize(self.critic_running_stats[a], obs[a]) for a in obs}

    # file names of the reference (:282-340)
    def save_info(self, path):
        if not self.test_mode:
            r = mpi_utils.get_rank()
            for f_name, stats in (("ActorRunningObsStats_{}.pickle".format(r), self.actor_running_stats),
                                  ("CriticRunningObsStats_{}.pickle".format(r), self.critic_running_stats)):
                with open(os.path.join(path, f_name), "wb") as fh:
                    pickle.dump(stats, fh)
        super().save_info(path)

    def load_info(self, path):
        r = 0 if self.test_mode else mpi_utils.get_rank()
        for attr, stem in (("actor_running_stats", "ActorRunningObsStats"), ("critic_running_stats", "CriticRunningObsStats")):
            f = os.path.join(path, "{}_{}.pickle".format(stem, r))
            if not os.path.exists(f):
                f = os.path.join(path, "{}_0.pickle".format(stem))
            with open(f, "rb") as fh:
                setattr(self, attr, pickle.load(fh))
        super().load_info(path)


class GenericClipper(IdentityWrapper):
    """:548-616."""

    def __init__(self, env, clip_range=(-10., 10.), **kw_args):
        super().__init__(env, **kw_args)
        self.clip_range = (_callable(clip_range[0]), _callable(clip_range[1]))

    def get_clip_range(self):
        return (self.clip_range[0](), self.clip_range[1]())

    def _clip(self, val):
        lo, hi = self.get_clip_range()
        return torch.clamp(_dev(val, self.device), lo, hi)

    def _apply_agent_clipping(self, agent_dict):
        return {a: self._clip(agent_dict[a]) for a in agent_dict}


class ObservationClipper(GenericClipper, ObservationFilter):
    """:617-671."""

    def _filter_critic_observation(self, obs):
        return self._apply_agent_clipping(obs)

    def _filter_local_observation(self, obs):
        return self._apply_agent_clipping(obs)

    def step(self, action):
        return ObservationFilter.step(self, action)

    def reset(self):
        return ObservationFilter.reset(self)


class RewardNormalizer(IdentityWrapper):
    """:343-476.  `clip_range` fuses a RewardClipper into the scaling kernel."""

    def __init__(self, env, update_stats=True, epsilon=1e-8, gamma=0.99, clip_range=None, **kw_args):
        super().__init__(env, **kw_args)
        self.running_stats = {a: RunningMeanStd(shape=(), device=self.device) for a in self.agent_ids}
        self.update_stats = update_stats
        self.epsilon = epsilon
        self.gamma = gamma
        self.batch_size = self.get_batch_size()
        self.running_reward = {a: torch.zeros(self.batch_size, dtype=torch.float64, device=self.device) for a in self.agent_ids}
        self._triples = torch.empty((self.batch_size, 3), dtype=torch.float64, device=self.device)
        self._scratch_stats = torch.empty((self.batch_size, 2), dtype=torch.float32, device=self.device)
        self.clip_range = None if clip_range is None else (_callable(clip_range[0]), _callable(clip_range[1]))

    def step(self, action):
        obs, critic_obs, reward, terminated, truncated, info = self.env.step(action)
        lib = load()
        out = {}
        for agent_id in reward:
            r = _dev(reward[agent_id], self.device).reshape(-1)
            done = torch.logical_or(_dev(terminated[agent_id], self.device, torch.bool).reshape(-1),
                                    _dev(truncated[agent_id], self.device, torch.bool).reshape(-1)).to(torch.uint8)
            if self.update_stats:
                check(lib.ppoaf_reward_norm_triples(ptr(r), ptr(done), ptr(self.running_reward[agent_id]), self.batch_size,
                                                    float(self.gamma), ptr(self._triples), stream_ptr()),
                      "ppoaf_reward_norm_triples")
                triples = self._triples
                n_ranks = 1
                if mpi_utils.get_num_procs() > 1:
                    triples = mpi_utils.all_gather_cat(self._triples).contiguous()         # [R, E, 3]
                    n_ranks = triples.shape[0]
                check(lib.ppoaf_value_stats_sequence(ptr(self.running_stats[agent_id].state), ptr(triples), n_ranks,
                                                     self.batch_size, float(self.epsilon), ptr(self._scratch_stats),
                                                     stream_ptr()), "ppoaf_value_stats_sequence")
            else:
                self.running_reward[agent_id].masked_fill_(done.bool(), 0.0)
            if isinstance(info, dict) and agent_id in info:          # "natural reward" book-keeping (:431-442)
                ai = info[agent_id]
                if isinstance(ai, (list, tuple, np.ndarray)):
                    for b_idx in range(len(ai)):
                        if isinstance(ai[b_idx], dict) and "natural reward" not in ai[b_idx]:
                            raw = reward[agent_id]
                            ai[b_idx]["natural reward"] = (raw.reshape(-1)[b_idx].item() if torch.is_tensor(raw)
                                                           else np.asarray(raw).reshape(-1)[b_idx].copy())
            out[agent_id] = self.normalize(agent_id, r).reshape(np.shape(reward[agent_id]) if not torch.is_tensor(reward[agent_id])
                                                                else reward[agent_id].shape)
        return obs, critic_obs, out, terminated, truncated, info

    def normalize(self, agent_id, agent_reward):
        r = _dev(agent_reward, self.device).reshape(-1)
        out = torch.empty_like(r)
        lo, hi = (1.0, -1.0) if self.clip_range is None else (self.clip_range[0](), self.clip_range[1]())
        check(load().ppoaf_reward_scale_clip(ptr(r), ptr(self.running_stats[agent_id].state), float(self.epsilon), float(lo),
                                             float(hi), ptr(out), r.numel(), stream_ptr()), "ppoaf_reward_scale_clip")
        return out

    def save_info(self, path):
        if not self.test_mode:
            with open(os.path.join(path, "RunningRewardsStats_{}.pickle".format(mpi_utils.get_rank())), "wb") as fh:
                pickle.dump(self.running_stats, fh)
        super().save_info(path)

    def load_info(self, path):
        r = 0 if self.test_mode else mpi_utils.get_rank()
        f = os.path.join(path, "RunningRewardsStats_{}.pickle".format(r))
        if not os.path.exists(f):
            f = os.path.join(path, "RunningRewardsStats_0.pickle")
        with open(f, "rb") as fh:
            self.running_stats = pickle.load(fh)
        super().load_info(path)


class RewardClipper(GenericClipper):
    """:674-719."""

    def step(self, actions):
        obs, critic_obs, reward, terminated, truncated, info = self.env.step(actions)
        return obs, critic_obs, self._apply_agent_clipping(reward), terminated, truncated, info
