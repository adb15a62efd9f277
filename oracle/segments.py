"""
Oracle (test infrastructure, see oracle/__init__.py): episode-segment bookkeeping,
reward-to-go and GAE, and the flattened dataset order.

Restates, array-style, what the reference does with per-env Python objects:
  EpisodeInfo.add_info / end_episode            utils/episode_info.py:303-399, 419-465
  compute_discounted_sums                        utils/episode_info.py:223-262
  _compute_gae_advantages / standard advantages  utils/episode_info.py:264-301, 401-417
  PPOPolicy.initialize_episodes / add_episode_info / end_episodes / get_bs_clip_range
                                                 policies/ppo_policy.py:474-504, 545-712, 1086-1112
  combine_episodes / PPODataset.build            utils/episode_info.py:44-135, 745-914

dtype notes: the GAE scan runs in float64 on deltas formed from
float32-rounded values; the reward-to-go scan sees float32-rounded rewards and
accumulates in float64 under the numpy the reference pins (<1.24) but in float32
under numpy>=2 (what the build container runs).  `rtg_accum` selects which;
"float64" is the semantics of record, "float32" reproduces the reference as run
here bit-for-bit (used to pin this file against tests/golden).
"""
import numpy as np


def discounted_sums(x, gamma, accum="float64"):
    """DS_t = x_t + gamma * DS_{t+1}, reverse sequential scan (episode_info.py:254-260)."""
    n = len(x)
    out = np.zeros(n, dtype=np.float64)
    if accum == "float64":
        acc = 0.0
        g = float(gamma)
        for i in range(n - 1, -1, -1):
            acc = float(x[i]) + g * acc
            out[i] = acc
    elif accum == "float32":
        acc = np.float32(0.0)
        g = np.float32(gamma)
        for i in range(n - 1, -1, -1):
            acc = np.float32(np.float32(x[i]) + np.float32(g * acc))
            out[i] = acc
    else:
        raise ValueError(accum)
    return out


def segment_returns(rewards, values, ending_value, ending_reward, gamma, lambd, use_gae,
                    bootstrap_clip, rtg_accum="float64"):
    """
    One closed segment -> (rewards_to_go float64[L], advantages float64[L], values float32[L]).
    Follows EpisodeInfo.end_episode (episode_info.py:419-465): the bootstrap reward is
    clipped (also for terminal segments), the bootstrap value is not.
    """
    rewards = [float(r) for r in rewards]
    if bootstrap_clip is not None:
        ending_reward = float(np.clip(ending_reward, bootstrap_clip[0], bootstrap_clip[1]))
    padded_rewards = np.array(rewards + [ending_reward], dtype=np.float32)
    rtg = discounted_sums(padded_rewards, gamma, rtg_accum)[:-1]
    vals32 = np.array(values, dtype=np.float64).astype(np.float32)
    if use_gae:
        padded = np.concatenate((vals32.astype(np.float64), [float(ending_value)])).astype(np.float32)
        gv = (np.float32(gamma) * padded[1:]).astype(np.float32)          # float32 product (:289)
        deltas = np.array(rewards, dtype=np.float64) + gv.astype(np.float64) - padded[:-1].astype(np.float64)
        adv = discounted_sums(deltas, gamma * lambd, "float64")
    else:
        adv = rtg - vals32.astype(np.float64)
    return rtg, adv, vals32


class OracleRolloutPolicy:
    """
    Stand-in for the bookkeeping half of the reference PPOPolicy + PPODataset: accepts
    the same calls the trainer makes during a rollout and produces the flat dataset.
    """

    def __init__(self, agent_ids, use_gae=True, gamma=0.99, lambd=0.95,
                 bootstrap_clip=(-100.0, 100.0), dynamic_bs_clip=False, discrete=False,
                 rtg_accum="float64"):
        self.agent_ids = list(agent_ids)
        self.use_gae, self.gamma, self.lambd = use_gae, gamma, lambd
        self.bootstrap_clip = None if bootstrap_clip is None else (float(bootstrap_clip[0]), float(bootstrap_clip[1]))
        self.dynamic_bs_clip = dynamic_bs_clip
        self.discrete = discrete
        self.rtg_accum = rtg_accum
        self.frozen = False

    # -- lifecycle (ppo_policy.py:474-526) -------------------------------------------
    def initialize_dataset(self):
        self.closed = []          # segments in completion order (= add_episode order)

    def initialize_episodes(self, env_batch_size, status_dict=None):
        self.E = env_batch_size
        self.open = {a: [self._new_segment(0, self.bootstrap_clip) for _ in range(env_batch_size)]
                     for a in self.agent_ids}

    @staticmethod
    def _new_segment(starting_ts, clip):
        return dict(starting_ts=int(starting_ts), clip=clip, rows=[])

    # -- per step (ppo_policy.py:545-651) ----------------------------------------------
    def add_episode_info(self, agent_id, critic_observations, observations, next_observations,
                         raw_actions, actions, values, log_probs, rewards, where_done):
        for e in range(self.E):
            self.open[agent_id][e]["rows"].append(dict(
                critic_obs=np.asarray(critic_observations[e]),
                obs=np.asarray(observations[e]),
                next_obs=np.asarray(next_observations[e]),
                raw_action=np.asarray(raw_actions[e]).squeeze() if np.ndim(raw_actions[e]) > 1 else np.asarray(raw_actions[e]),
                action=np.asarray(actions[e]).squeeze() if np.ndim(actions[e]) > 1 else np.asarray(actions[e]),
                value=float(np.asarray(values[e]).item()),
                log_prob=float(np.asarray(log_probs[e]).reshape(-1)[0]),
                reward=float(np.asarray(rewards[e]).item())))

    # -- segment closure (ppo_policy.py:653-712) -----------------------------------------
    def end_episodes(self, agent_id, env_idxs, episode_lengths, terminal, ending_values, ending_rewards):
        if self.frozen:
            return
        for pos, env_i in enumerate(env_idxs):
            seg = self.open[agent_id][env_i]
            ending_ts = int(episode_lengths[env_i])
            # bootstrap arrays are indexed by POSITION in env_idxs, not by env
            ev = float(_item(ending_values[pos]))
            er = float(_item(ending_rewards[pos]))
            rewards = [r["reward"] for r in seg["rows"]]
            rtg, adv, vals32 = segment_returns(
                rewards, [r["value"] for r in seg["rows"]], ev, er, self.gamma, self.lambd,
                self.use_gae, seg["clip"], self.rtg_accum)
            seg.update(length=ending_ts - seg["starting_ts"], ending_ts=ending_ts,
                       terminal=bool(terminal[pos]), ending_value=ev, rtg=rtg, adv=adv, values32=vals32,
                       agent=agent_id, env=int(env_i))
            self.closed.append(seg)
            if self.bootstrap_clip is None:
                clip = None
            elif self.dynamic_bs_clip:
                clip = (min(rewards), max(rewards))                     # ppo_policy.py:1104-1106
            else:
                clip = self.bootstrap_clip
            start = 0 if terminal[pos] else ending_ts
            self.open[agent_id][env_i] = self._new_segment(start, clip)

    # -- flatten (episode_info.py:44-135, 745-914) ---------------------------------------
    def finalize_dataset(self):
        segs = self.closed
        rows = [r for s in segs for r in s["rows"]]
        act_dtype = np.int64 if self.discrete else np.float32
        out = dict(
            ep_lens=np.array([s["length"] for s in segs], dtype=np.int64),
            seg_terminal=np.array([s["terminal"] for s in segs], dtype=bool),
            seg_start_ts=np.array([s["starting_ts"] for s in segs], dtype=np.int64),
            seg_end_ts=np.array([s["ending_ts"] for s in segs], dtype=np.int64),
            seg_end_value=np.array([s["ending_value"] for s in segs], dtype=np.float64),
            seg_agent=np.array([s["agent"] for s in segs]),
            seg_env=np.array([s["env"] for s in segs], dtype=np.int64),
            advantages_f64=np.concatenate([s["adv"] for s in segs]),
            rewards_to_go_f64=np.concatenate([s["rtg"] for s in segs]),
            values=np.concatenate([s["values32"] for s in segs]).astype(np.float32),
            log_probs=np.array([r["log_prob"] for r in rows], dtype=np.float32),
            observations=np.array([r["obs"] for r in rows], dtype=np.float32),
            next_observations=np.array([r["next_obs"] for r in rows], dtype=np.float32),
            critic_observations=np.array([r["critic_obs"] for r in rows], dtype=np.float32),
            actions=np.array([r["action"] for r in rows]).astype(act_dtype),
            raw_actions=np.array([r["raw_action"] for r in rows]).astype(act_dtype),
        )
        out["advantages"] = out["advantages_f64"].astype(np.float32)
        out["rewards_to_go"] = out["rewards_to_go_f64"].astype(np.float32)
        for k in ("actions", "raw_actions"):
            if out[k].ndim <= 1:
                out[k] = out[k][:, None]
        total = int(out["ep_lens"].sum())
        assert total == len(rows), (total, len(rows))
        self.dataset = out
        return out


def _item(x):
    if hasattr(x, "detach"):
        x = x.detach().cpu().numpy()
    return np.asarray(x).item()


def flat_segscan_reference(rewards, values, seg_lens, v_boot, r_boot_clipped, gamma, lambd, use_gae=True):
    """
    The same arithmetic as `segment_returns`, stated over the FLATTENED buffer the CUDA
    scan consumes: rewards/values fp32 [N] in dataset order, `seg_lens` int [n_seg],
    per-segment seeds.  float64 accumulation.  Returns (adv f64[N], rtg f64[N]).
    """
    N = int(np.sum(seg_lens))
    adv = np.zeros(N)
    rtg = np.zeros(N)
    off = 0
    for s, L in enumerate(seg_lens):
        L = int(L)
        r = rewards[off:off + L]
        v = values[off:off + L]
        g, a, _ = segment_returns(r, v, v_boot[s], r_boot_clipped[s], gamma, lambd, use_gae, None)
        rtg[off:off + L] = g
        adv[off:off + L] = a
        off += L
    return adv, rtg
