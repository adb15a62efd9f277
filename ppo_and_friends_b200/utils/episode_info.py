"""
Device-resident rollout storage and dataset (reference utils/episode_info.py).

The reference keeps one Python `EpisodeInfo` object per (env, agent) trajectory segment, appends
Python lists per step, runs per-element Python loops for the discounted sums at `end_episode`,
and concatenates everything at `PPODataset.build` (utils/episode_info.py:138-465, 647-914).
Here the same information lives in

  * `RolloutRing`  — one packed, time-major ring [T, E*A, row] (pinned host staging + device copy,
                     one async H2D per `add_episode_info` call) and
  * a segment table — (column, first ring step, length, terminal, bootstrap value, clipped
                     bootstrap reward) appended by `end_episodes` in completion order, which IS the
                     reference's dataset order (combine_episodes, :44-135),

and `PPODataset.build` turns them into the flat arrays with three kinds of kernels: the flat map
(segment table -> source row per element), strided row gathers (de-interleave the ring into
`observations`, `critic_observations`, ...), and ONE segmented reverse scan that produces both
`advantages` and `rewards_to_go` (replacing EpisodeInfo.end_episode / compute_discounted_sums /
_compute_gae_advantages, :223-301, 401-465).
"""
import numpy as np
import torch

from .. import ops
from .mpi_utils import abort

_F32 = 4


def _align4(n):
    return (n + 3) // 4 * 4


class RolloutRing:
    """Packed time-major ring: row(t, col) = all per-timestep fields of one (step, env, agent)."""

    FIELDS = ("critic_obs", "obs", "next_obs", "raw_action", "action", "value", "log_prob", "reward")

    def __init__(self, device, n_envs, n_agents, obs_dim, critic_obs_dim, act_dim, discrete, capacity_steps=64):
        self.device = torch.device(device)
        self.E, self.A = int(n_envs), int(n_agents)
        self.C = self.E * self.A
        self.discrete = bool(discrete)
        act_words = act_dim * (2 if discrete else 1)          # int64 actions occupy two fp32 words each
        widths = dict(critic_obs=critic_obs_dim, obs=obs_dim, next_obs=obs_dim, raw_action=act_words,
                      action=act_words, value=1, log_prob=1, reward=1)
        self.widths, self.offsets = widths, {}
        off = 0
        for f in ("critic_obs", "obs", "next_obs", "raw_action", "action"):   # 16-byte aligned field starts
            self.offsets[f] = off
            off += _align4(widths[f])
        for f in ("value", "log_prob", "reward"):                              # the scalars share one slot
            self.offsets[f] = off
            off += 1
        self.row_words = _align4(off)
        self.act_dim = act_dim
        self.capacity = 0
        self.host = None
        self.dev = None
        self.copy_stream = torch.cuda.Stream(device=self.device) if self.device.type == "cuda" else None
        self._grow(int(capacity_steps))
        self.reset()

    def _grow(self, capacity):
        host = torch.zeros((capacity, self.C, self.row_words), dtype=torch.float32,
                           pin_memory=self.device.type == "cuda")
        dev = torch.zeros((capacity, self.C, self.row_words), dtype=torch.float32, device=self.device)
        if self.host is not None:
            if self.copy_stream is not None:
                self.copy_stream.synchronize()
            host[:self.capacity].copy_(self.host)
            dev[:self.capacity].copy_(self.dev)
        if self.copy_stream is not None:
            # the zero-fill / carry-over above ran on the current stream; the per-step H2D copies run on copy_stream and
            # must not be overtaken by them
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        self.host, self.dev, self.capacity = host, dev, capacity
        self.host_np = self.host.numpy()

    def reset(self):
        self.steps = np.zeros(self.A, dtype=np.int64)     # add_step calls so far, per agent

    def add_step(self, agent_idx, critic_obs, obs, next_obs, raw_actions, actions, values, log_probs, rewards):
        t = int(self.steps[agent_idx])
        if t >= self.capacity:
            self._grow(max(2 * self.capacity, t + 1))
        c0, c1 = agent_idx * self.E, (agent_idx + 1) * self.E
        slab = self.host_np[t, c0:c1]
        o, w = self.offsets, self.widths
        slab[:, o["critic_obs"]:o["critic_obs"] + w["critic_obs"]] = np.asarray(critic_obs).reshape(self.E, -1)
        slab[:, o["obs"]:o["obs"] + w["obs"]] = np.asarray(obs).reshape(self.E, -1)
        slab[:, o["next_obs"]:o["next_obs"] + w["next_obs"]] = np.asarray(next_obs).reshape(self.E, -1)
        if self.discrete:
            ra = np.ascontiguousarray(np.asarray(raw_actions).reshape(self.E, -1).astype(np.int64))
            ac = np.ascontiguousarray(np.asarray(actions).reshape(self.E, -1).astype(np.int64))
            slab[:, o["raw_action"]:o["raw_action"] + w["raw_action"]] = ra.view(np.float32)
            slab[:, o["action"]:o["action"] + w["action"]] = ac.view(np.float32)
        else:
            slab[:, o["raw_action"]:o["raw_action"] + w["raw_action"]] = _np(raw_actions).reshape(self.E, -1)
            slab[:, o["action"]:o["action"] + w["action"]] = _np(actions).reshape(self.E, -1)
        slab[:, o["value"]] = _np(values).reshape(self.E)
        slab[:, o["log_prob"]] = _np(log_probs).reshape(self.E)
        slab[:, o["reward"]] = _np(rewards).reshape(self.E)
        if self.copy_stream is not None:
            with torch.cuda.stream(self.copy_stream):
                self.dev[t, c0:c1].copy_(self.host[t, c0:c1], non_blocking=True)
        else:
            self.dev[t, c0:c1].copy_(self.host[t, c0:c1])
        self.steps[agent_idx] = t + 1
        return t

    def wait_copies(self):
        if self.copy_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)

    def segment_rewards(self, col, t0, length):
        return self.host_np[t0:t0 + length, col, self.offsets["reward"]]


def _np(x):
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


class PPODataset(object):
    """
    Device dataset with the reference PPODataset's surface (utils/episode_info.py:647-987):
    `build()`, `recalculate_advantages()`, `__len__`, `__getitem__` (13-tuple), and the attributes
    `actions, raw_actions, critic_observations, observations, next_observations, rewards_to_go,
    log_probs, ep_lens, advantages, values` as CUDA tensors in the reference's flat order.
    Segments are added with `add_segment` (by PPOPolicy.end_episodes) instead of `add_episode`.
    """

    def __init__(self, device, action_dtype, sequence_length=1, ring=None, use_gae=True, gamma=0.99, lambd=0.95):
        if sequence_length != 1:
            abort("ERROR: the B200 path supports sequence_length == 1 only (no LSTM datasets).")
        self.device = torch.device(device)
        self.action_dtype = action_dtype
        self.sequence_length = 1
        self.ring = ring
        self.use_gae, self.gamma, self.lambd = use_gae, gamma, lambd
        self.is_built = False
        self.shared = False
        self.build_hidden_states = False
        self._seg = dict(col=[], t0=[], length=[], terminal=[], v_boot=[], r_boot=[], start_ts=[], end_ts=[])
        self.total_timestates = 0
        self.actions = self.raw_actions = self.critic_observations = self.observations = None
        self.next_observations = self.rewards_to_go = self.log_probs = self.ep_lens = None
        self.advantages = self.values = None

    # -- A2: one closed segment ---------------------------------------------------------------------
    def add_segment(self, col, t0, length, terminal, ending_value, clipped_ending_reward, starting_ts, ending_ts):
        if int(length) < 1:
            # the reference never closes an empty episode (end_episodes runs after add_episode_info of the same step); an
            # empty segment would shift the per-segment bootstrap values of every later segment in the scan
            abort("ERROR: attempting to close an episode segment of length {}".format(length))
        s = self._seg
        s["col"].append(col); s["t0"].append(t0); s["length"].append(length); s["terminal"].append(bool(terminal))
        s["v_boot"].append(ending_value); s["r_boot"].append(clipped_ending_reward)
        s["start_ts"].append(starting_ts); s["end_ts"].append(ending_ts)

    def add_segments(self, cols, t0s, lengths, terminals, ending_values, clipped_ending_rewards, starting_ts, ending_ts):
        """Several segments closed by one `end_episodes` call (same meaning as add_segment, array arguments)."""
        lengths = np.asarray(lengths)
        if lengths.size and int(lengths.min()) < 1:
            abort("ERROR: attempting to close an episode segment of length {}".format(int(lengths.min())))
        s = self._seg
        s["col"].extend(np.asarray(cols).tolist()); s["t0"].extend(np.asarray(t0s).tolist())
        s["length"].extend(lengths.tolist()); s["terminal"].extend(np.asarray(terminals, dtype=bool).tolist())
        s["v_boot"].extend(np.asarray(ending_values, dtype=np.float64).tolist())
        s["r_boot"].extend(np.asarray(clipped_ending_rewards, dtype=np.float64).tolist())
        s["start_ts"].extend(np.asarray(starting_ts).tolist()); s["end_ts"].extend(np.asarray(ending_ts).tolist())

    @property
    def num_segments(self):
        return len(self._seg["col"])

    # -- A5/A6 + A3/A4 ------------------------------------------------------------------------------
    def build(self):
        if self.is_built:
            abort("ERROR: attempting to build a batch, but it's already been built! Bailing...")
        s = self._seg
        n_seg = len(s["col"])
        lens = np.asarray(s["length"], dtype=np.int64)
        ts_lens = np.asarray(s["end_ts"], dtype=np.int64) - np.asarray(s["start_ts"], dtype=np.int64)
        self.ep_lens = ts_lens                                   # EpisodeInfo.length = ending_ts - starting_ts (:439)
        self.total_timestates = int(ts_lens.sum())
        n = int(lens.sum())
        if self.total_timestates != n:
            abort("ERROR: expected the total timestates to match the total number of observations, "
                  "but got {} vs {}".format(self.total_timestates, n))
        self.seg_terminal = np.asarray(s["terminal"], dtype=bool)
        off = np.zeros(n_seg + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        dev = self.device
        i32 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.int32)).to(dev, non_blocking=True)
        self.seg_off = torch.as_tensor(off).to(dev, non_blocking=True)
        self.v_boot = torch.as_tensor(np.asarray(s["v_boot"], dtype=np.float64).astype(np.float32)).to(dev, non_blocking=True)
        self.r_boot = torch.as_tensor(np.asarray(s["r_boot"], dtype=np.float64).astype(np.float32)).to(dev, non_blocking=True)
        term_u8 = torch.as_tensor(self.seg_terminal.astype(np.uint8)).to(dev, non_blocking=True)
        ring = self.ring
        ring.wait_copies()
        self.src_row, self.seg_flag = ops.build_flat_map(i32(s["col"]), i32(s["t0"]), i32(lens), self.seg_off,
                                                         term_u8, ring.C, n)
        stride = ring.row_words * _F32
        flat_ring = ring.dev.view(-1, ring.row_words)

        def field(name, width_words, out_dtype=torch.float32, out_cols=None):
            cols = width_words if out_cols is None else out_cols
            return ops.gather_rows(flat_ring, self.src_row, row_bytes=width_words * _F32, src_stride_bytes=stride,
                                   src_offset_bytes=ring.offsets[name] * _F32, n_rows=n, out_shape=(n, cols),
                                   out_dtype=out_dtype)

        w = ring.widths
        self.critic_observations = field("critic_obs", w["critic_obs"])
        self.observations = field("obs", w["obs"])
        self.next_observations = field("next_obs", w["next_obs"])
        if ring.discrete:
            self.raw_actions = field("raw_action", w["raw_action"], torch.int64, ring.act_dim)
            self.actions = field("action", w["action"], torch.int64, ring.act_dim)
        else:
            self.raw_actions = field("raw_action", w["raw_action"])
            self.actions = field("action", w["action"])
        self.values = field("value", 1).reshape(n)
        self.log_probs = field("log_prob", 1).reshape(n)
        self.rewards = field("reward", 1).reshape(n)
        self.advantages, self.rewards_to_go = ops.gae_rtg_segscan(
            self.rewards, self.values, self.seg_flag, self.seg_off, self.v_boot, self.r_boot, self.gamma, self.lambd,
            self.use_gae)
        self.is_built = True

    # -- P7 -----------------------------------------------------------------------------------------
    def recalculate_advantages(self):
        """Re-run the scan with the current `values` (utils/episode_info.py:721-743)."""
        if not self.is_built:
            from .mpi_utils import rank_print
            rank_print("WARNING: recalculate_advantages was called before the dataset has been built. Ignoring call.")
            return
        scratch_rtg = torch.empty_like(self.rewards_to_go)      # the reference rebuilds advantages only
        ops.gae_rtg_segscan(self.rewards, self.values, self.seg_flag, self.seg_off, self.v_boot, self.r_boot,
                            self.gamma, self.lambd, self.use_gae, adv_out=self.advantages, rtg_out=scratch_rtg)
        if not self.use_gae:
            # non-GAE advantages are rewards_to_go - values with the ORIGINAL rewards_to_go, which the scan reproduces
            pass

    def __len__(self):
        return self.total_timestates

    def __getitem__(self, idx):
        empty = torch.zeros((), dtype=torch.uint8)
        return (self.critic_observations[idx], self.observations[idx], self.next_observations[idx],
                self.raw_actions[idx], self.actions[idx], self.advantages[idx], self.log_probs[idx],
                self.rewards_to_go[idx], empty, empty, empty, empty, idx)
