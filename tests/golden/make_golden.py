"""
Generate the golden fixtures under tests/golden/*.npz by running the UNMODIFIED
reference (/root/reference, imported through ref_harness.py) on seeded synthetic
inputs.  Run in the build container only:

    python tests/golden/make_golden.py

The reference's own test-suite holds no golden vectors or known-answer tests for
this path, so these fixtures — outputs of the reference code
itself on recorded inputs — are what pins the oracle (oracle/) and, through it, the
CUDA path.  Everything a test needs (inputs and outputs) is stored in the .npz, so
nothing reads /root/reference at test time.

Reference entry points exercised (file:line in /root/reference):
  policies/ppo_policy.py:474-526,545-719  initialize_*/add_episode_info/end_episodes/finalize_dataset
  utils/episode_info.py:223-301,401-465   EpisodeInfo scans           :745-914 PPODataset.build
  utils/stats.py:29-94                    RunningMeanStd              utils/misc.py:84-128 normaliser
  ppo.py:2274-2485                        PPO._ppo_batch_train (via object.__new__(PPO))
  policies/ppo_policy.py:891-952,1012-1055 evaluate / update_weights
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import ref_harness  # noqa: E402

ref_harness.install()

import gymnasium.spaces as sp  # noqa: E402  (stub)
from ppo_and_friends.policies.ppo_policy import PPOPolicy  # noqa: E402
from ppo_and_friends.ppo import PPO  # noqa: E402
from ppo_and_friends.networks.ppo_networks.feed_forward import FeedForwardNetwork  # noqa: E402
from ppo_and_friends.utils.stats import RunningMeanStd  # noqa: E402
from ppo_and_friends.utils.misc import RunningStatNormalizer  # noqa: E402
from torch.utils.data import DataLoader  # noqa: E402

from ppo_and_friends_b200.synthetic import make_rollout, replay_rollout  # noqa: E402

ACTS = {"leaky_relu": torch.nn.LeakyReLU, "tanh": torch.nn.Tanh, "relu": torch.nn.ReLU}


def build_policy(ro, act="leaky_relu", actor_hidden=32, critic_hidden=32, depth=3,
                 dist_range=1.0, **policy_kw):
    if ro.n_discrete:
        action_space = sp.Discrete(ro.n_discrete)
    else:
        action_space = sp.Box(-dist_range, dist_range, (ro.act_dim,))
    pol = PPOPolicy(
        "pol", action_space,
        sp.Box(-np.inf, np.inf, (ro.obs_dim,)),
        sp.Box(-np.inf, np.inf, (ro.critic_obs_dim,)),
        envs_per_proc=ro.E,
        ac_network=FeedForwardNetwork,
        actor_kw_args=dict(activation=ACTS[act](), hidden_size=actor_hidden, hidden_depth=depth),
        critic_kw_args=dict(activation=ACTS[act](), hidden_size=critic_hidden, hidden_depth=depth),
        **policy_kw)
    for a in ro.agents:
        pol.register_agent(a)
    status = {"global status": {"iteration": 0, "timesteps": 0}}
    pol.finalize(status, torch.device("cpu"))
    return pol


def fill_policy_outputs(pol, ro, seed):
    """Overwrite values / raw_actions / actions / log_probs / next_values with what the
    reference networks produce on the synthetic observations (so ratios start near 1)."""
    torch.manual_seed(seed)
    with torch.no_grad():
        for a in ro.agents:
            T, E = ro.T, ro.E
            obs = torch.tensor(ro.obs[a].reshape(T * E, -1))
            cobs = torch.tensor(ro.critic_obs[a].reshape(T * E, -1))
            ro.values[a] = pol.critic(cobs).reshape(T, E).numpy().copy()
            nxt = torch.tensor(np.roll(ro.critic_obs[a], -1, axis=0).reshape(T * E, -1))
            ro.next_values[a] = pol.critic(nxt).reshape(T, E).numpy().copy()
            pred = pol.actor(obs)
            dist = pol.actor.distribution.get_distribution(pred)
            action, raw = pol.actor.distribution.sample_distribution(dist)
            lp = pol.actor.distribution.get_log_probs(dist, raw)
            ro.log_probs[a] = lp.reshape(T, E).numpy().astype(np.float32).copy()
            if ro.n_discrete:
                ro.raw_actions[a] = raw.reshape(T, E, 1).numpy().copy()
                ro.actions[a] = action.reshape(T, E, 1).numpy().copy()
            else:
                ro.raw_actions[a] = raw.reshape(T, E, -1).numpy().copy()
                ro.actions[a] = action.reshape(T, E, -1).numpy().copy()


def rollout_inputs(ro, prefix="in_"):
    d = dict(T=ro.T, E=ro.E, agents=np.array(ro.agents), obs_dim=ro.obs_dim,
             critic_obs_dim=ro.critic_obs_dim, act_dim=ro.act_dim, n_discrete=ro.n_discrete,
             max_ts_per_ep=ro.max_ts_per_ep, terminated=ro.terminated, truncated=ro.truncated)
    for a in ro.agents:
        for name in ("obs", "next_obs", "critic_obs", "raw_actions", "actions", "values",
                     "log_probs", "rewards", "next_values"):
            d[f"{name}/{a}"] = getattr(ro, name)[a]
    return {prefix + k: v for k, v in d.items()}


def run_reference_rollout(pol, ro, tensor_bootstrap=True):
    pol.initialize_dataset()
    pol.initialize_episodes(ro.E, {"global status": {"iteration": 0, "timesteps": 0}})
    wrap = (lambda x: torch.tensor(x)) if tensor_bootstrap else None
    # The reference receives log_probs as torch tensors ([E] or [E,1]).
    lp_backup = dict(ro.log_probs)
    for a in ro.agents:
        ro.log_probs[a] = torch.tensor(lp_backup[a])
    events = replay_rollout(lambda a: pol, ro, to_bootstrap=wrap)
    ro.log_probs = lp_backup
    # float64 per-episode results before PPODataset.build casts to fp32
    adv64 = np.concatenate([np.asarray(ep.advantages, dtype=np.float64) for ep in pol.dataset.episodes])
    rtg_asrun = np.concatenate([np.asarray(ep.rewards_to_go) for ep in pol.dataset.episodes])
    terminal = np.array([bool(ep.terminal) for ep in pol.dataset.episodes])
    start_ts = np.array([ep.starting_ts for ep in pol.dataset.episodes], dtype=np.int64)
    end_ts = np.array([ep.ending_ts for ep in pol.dataset.episodes], dtype=np.int64)
    end_val = np.array([ep.ending_value for ep in pol.dataset.episodes], dtype=np.float64)
    pol.finalize_dataset()
    ds = pol.dataset
    out = dict(
        ep_lens=np.asarray(ds.ep_lens, dtype=np.int64),
        seg_terminal=terminal, seg_start_ts=start_ts, seg_end_ts=end_ts, seg_end_value=end_val,
        advantages=ds.advantages.numpy(), rewards_to_go=ds.rewards_to_go.numpy(),
        advantages_f64=adv64, rewards_to_go_asrun=rtg_asrun.astype(np.float64),
        rewards_to_go_asrun_dtype=np.array(str(rtg_asrun.dtype)),
        values=ds.values.numpy(), log_probs=ds.log_probs.numpy(),
        observations=ds.observations.numpy(), next_observations=ds.next_observations.numpy(),
        critic_observations=ds.critic_observations.numpy(),
        actions=ds.actions.numpy(), raw_actions=ds.raw_actions.numpy(),
        n_events=len(events))
    return out


def gen_segments(only=None):
    cases = {
        "seg_single": dict(ro=dict(seed=11, T=48, E=4, obs_dim=3, act_dim=2, max_ts_per_ep=8,
                                   p_term=0.05, p_trunc=0.04), pol={}),
        "seg_multi": dict(ro=dict(seed=12, T=40, E=3, agents=("a0", "a1", "a2"), obs_dim=4,
                                  critic_obs_dim=12, n_discrete=5, max_ts_per_ep=16,
                                  p_term=0.06, p_trunc=0.03, shared_critic_obs=True), pol={}),
        "seg_bsclip": dict(ro=dict(seed=13, T=33, E=5, obs_dim=2, act_dim=1, max_ts_per_ep=7,
                                   p_term=0.08, p_trunc=0.05), pol=dict(bootstrap_clip=(0.01, 0.5))),
        "seg_nogae": dict(ro=dict(seed=14, T=30, E=2, obs_dim=2, act_dim=1, max_ts_per_ep=9,
                                  p_term=0.05, p_trunc=0.05), pol=dict(use_gae=False)),
        "seg_dynclip": dict(ro=dict(seed=15, T=36, E=3, obs_dim=2, act_dim=1, max_ts_per_ep=6,
                                    p_term=0.05, p_trunc=0.05), pol=dict(dynamic_bs_clip=True)),
        "seg_noclip": dict(ro=dict(seed=16, T=20, E=2, obs_dim=2, act_dim=1, max_ts_per_ep=64,
                                   p_term=0.1, p_trunc=0.1), pol=dict(bootstrap_clip=None)),
        "seg_long": dict(ro=dict(seed=17, T=300, E=2, obs_dim=1, act_dim=1, max_ts_per_ep=300,
                                 p_term=0.004, p_trunc=0.0), pol=dict(gamma=0.995, lambd=0.97)),
        # edge shapes: every step ends its episode (all segments have length 1) ...
        "seg_len1": dict(ro=dict(seed=18, T=12, E=3, obs_dim=2, act_dim=1, max_ts_per_ep=50,
                                 p_term=1.0, p_trunc=0.0), pol={}),
        # ... nothing ever ends: one open segment per env, closed (and bootstrapped) at the end of the rollout ...
        "seg_open": dict(ro=dict(seed=19, T=25, E=4, obs_dim=2, act_dim=2, max_ts_per_ep=1000,
                                 p_term=0.0, p_trunc=0.0), pol={}),
        # ... and every step is cut by max_ts_per_ep = 1 (truncation with bootstrapping on every step)
        "seg_max1": dict(ro=dict(seed=20, T=10, E=2, obs_dim=2, act_dim=1, max_ts_per_ep=1,
                                 p_term=0.0, p_trunc=0.0), pol={}),
    }
    for name, c in cases.items():
        if only is not None and name not in only:
            continue
        ro = make_rollout(**c["ro"])
        # widen rewards/bootstraps so clipping actually bites in the clip cases
        for a in ro.agents:
            ro.next_values[a] = ro.next_values[a] * 3.0
        pol = build_policy(ro, **c["pol"])
        out = run_reference_rollout(pol, ro)
        meta = dict(use_gae=bool(pol.use_gae), gamma=pol.gamma, lambd=pol.lambd,
                    dynamic_bs_clip=bool(pol.dynamic_bs_clip),
                    have_bootstrap_clip=bool(pol.have_bootstrap_clip),
                    bootstrap_clip=np.array([pol.bootstrap_clip[0](), pol.bootstrap_clip[1]()])
                    if pol.have_bootstrap_clip else np.array([np.nan, np.nan]))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rollout_inputs(ro), **out, **meta)
        print(name, "N =", len(out["advantages"]), "n_seg =", len(out["ep_lens"]),
              "rtg dtype as run:", out["rewards_to_go_asrun_dtype"])


def gen_stats():
    rng = np.random.default_rng(21)
    out = {}
    # vector stats, float64 batches (observation normaliser style, filter_wrappers.py:155-258)
    rms = RunningMeanStd(shape=(5,))
    batches = [rng.normal(2.0, 3.0, size=(n, 5)) for n in (7, 1, 16)]
    for i, b in enumerate(batches):
        rms.update(b)
        out[f"vec64_batch{i}"] = b
        out[f"vec64_mean{i}"] = np.asarray(rms.mean)
        out[f"vec64_var{i}"] = np.asarray(rms.variance)
        out[f"vec64_count{i}"] = np.float64(rms.count)
    # vector stats, float32 batches
    rms = RunningMeanStd(shape=(5,))
    for i, b in enumerate(batches):
        b32 = b.astype(np.float32)
        rms.update(b32)
        out[f"vec32_mean{i}"] = np.asarray(rms.mean)
        out[f"vec32_var{i}"] = np.asarray(rms.variance)
        out[f"vec32_count{i}"] = np.float64(rms.count)
        out[f"vec32_dtype{i}"] = np.array(str(np.asarray(rms.mean).dtype))
    # scalar stats through RunningStatNormalizer (value normaliser, ppo.py:2299-2303)
    norm = RunningStatNormalizer("vn", torch.device("cpu"))
    for i, n in enumerate((64, 64, 1, 50)):
        x = torch.tensor(rng.normal(-1.0, 4.0, size=n).astype(np.float32))
        y = norm.normalize(x)
        out[f"sc_in{i}"] = x.numpy()
        out[f"sc_out{i}"] = y.numpy()
        out[f"sc_mean{i}"] = np.asarray(norm.running_stats.mean)
        out[f"sc_var{i}"] = np.asarray(norm.running_stats.variance)
        out[f"sc_count{i}"] = np.float64(norm.running_stats.count)
    z = torch.tensor(rng.normal(size=9).astype(np.float32))
    out["sc_denorm_in"] = z.numpy()
    out["sc_denorm_out"] = norm.denormalize(z).numpy()
    out["sc_noupdate_out"] = norm.normalize(z, update_stats=False).numpy()
    np.savez_compressed(os.path.join(HERE, "stats.npz"), **out)
    print("stats ok")


class _RecordingDataset(torch.utils.data.Dataset):
    """Forwards to the reference PPODataset and records the indices the DataLoader draws."""

    def __init__(self, ds):
        self._ds = ds
        self.drawn = []

    def __len__(self):
        return len(self._ds)

    def __getitem__(self, idx):
        self.drawn.append(int(idx))
        return self._ds[idx]

    @property
    def values(self):
        return self._ds.values

    def recalculate_advantages(self):
        self._ds.recalculate_advantages()


def snapshot_net(prefix, net, optim):
    out = {}
    for k, v in net.state_dict().items():
        out[f"{prefix}/param/{k}"] = v.detach().numpy().copy()
    st = optim.state_dict()["state"]
    names = [k for k, _ in net.named_parameters()]
    for i, k in enumerate(names):
        if i in st:
            out[f"{prefix}/exp_avg/{k}"] = st[i]["exp_avg"].numpy().copy()
            out[f"{prefix}/exp_avg_sq/{k}"] = st[i]["exp_avg_sq"].numpy().copy()
            out[f"{prefix}/step/{k}"] = np.float64(float(st[i]["step"]))
    return out


def gen_updates():
    cases = {
        "upd_gauss": dict(
            ro=dict(seed=31, T=64, E=4, obs_dim=8, act_dim=2, max_ts_per_ep=16, p_term=0.03,
                    p_trunc=0.02, obs_scale=False),
            net=dict(act="leaky_relu", actor_hidden=32, critic_hidden=48),
            pol=dict(lr=3e-4), B=64, epochs=2, ppo={}),
        "upd_cat": dict(
            ro=dict(seed=32, T=32, E=2, agents=("a0", "a1", "a2"), obs_dim=6, critic_obs_dim=18,
                    n_discrete=5, max_ts_per_ep=12, p_term=0.04, p_trunc=0.02, obs_scale=False,
                    shared_critic_obs=True),
            net=dict(act="tanh", actor_hidden=32, critic_hidden=40),
            pol=dict(lr=1e-3), B=50, epochs=2, ppo={}),
        "upd_opts": dict(
            ro=dict(seed=33, T=40, E=3, obs_dim=5, act_dim=3, max_ts_per_ep=10, p_term=0.03,
                    p_trunc=0.03, obs_scale=False),
            net=dict(act="relu", actor_hidden=24, critic_hidden=24, depth=2, dist_range=0.4),
            pol=dict(lr=5e-4, use_huber_loss=True, kl_loss_weight=0.1, vf_clip=0.5,
                     entropy_weight=0.0, surr_clip=0.1, gradient_clip=0.3),
            B=32, epochs=1, ppo={}),
        "upd_skip1": dict(
            ro=dict(seed=34, T=13, E=5, obs_dim=4, act_dim=2, max_ts_per_ep=6, p_term=0.05,
                    p_trunc=0.02, obs_scale=False),
            net=dict(act="tanh", actor_hidden=16, critic_hidden=16),
            pol=dict(lr=1e-3), B=64, epochs=1, ppo=dict(normalize_adv=True)),
        "upd_nonorm": dict(
            ro=dict(seed=35, T=24, E=4, obs_dim=4, n_discrete=3, max_ts_per_ep=8, p_term=0.05,
                    p_trunc=0.02, obs_scale=False),
            net=dict(act="leaky_relu", actor_hidden=16, critic_hidden=16),
            pol=dict(lr=1e-3, gradient_clip=None), B=32, epochs=1,
            ppo=dict(normalize_adv=False, normalize_values=False)),
    }
    for name, c in cases.items():
        ro = make_rollout(**c["ro"])
        torch.manual_seed(1000 + c["ro"]["seed"])
        pol = build_policy(ro, **c["net"], **c["pol"])
        fill_policy_outputs(pol, ro, seed=2000 + c["ro"]["seed"])
        seg = run_reference_rollout(pol, ro)
        pid = "pol"
        ppo = object.__new__(PPO)
        ppo.policies = {pid: pol}
        ppo.normalize_values = c["ppo"].get("normalize_values", True)
        ppo.normalize_adv = c["ppo"].get("normalize_adv", True)
        ppo.value_normalizers = {pid: RunningStatNormalizer(pid + "-value_normalizer", torch.device("cpu"))}
        ppo.status_dict = {pid: {}, "global status": {"iteration": 0, "timesteps": 0}}
        ppo.user_huber_loss = pol.use_huber_loss  # attribute the reference forgot
        out = {}
        out.update(snapshot_net("init/actor", pol.actor, pol.actor_optim))
        out.update(snapshot_net("init/critic", pol.critic, pol.critic_optim))
        rec = _RecordingDataset(pol.dataset)
        loader = DataLoader(rec, batch_size=c["B"], shuffle=True)
        perm_seed = 3000 + c["ro"]["seed"]
        torch.manual_seed(perm_seed)
        pol.train()
        for ep in range(c["epochs"]):
            rec.drawn = []
            ppo._ppo_batch_train(loader, pid)
            out[f"ep{ep}/batch_idxs"] = np.array(rec.drawn, dtype=np.int64)
            sd = ppo.status_dict[pid]
            out[f"ep{ep}/status"] = np.array([sd["actor loss"], sd["critic loss"], sd["kl avg"],
                                              sd["weighted entropy"]], dtype=np.float64)
            out.update(snapshot_net(f"ep{ep}/actor", pol.actor, pol.actor_optim))
            out.update(snapshot_net(f"ep{ep}/critic", pol.critic, pol.critic_optim))
            rs = ppo.value_normalizers[pid].running_stats
            out[f"ep{ep}/vn"] = np.array([float(rs.mean), float(rs.variance), float(rs.count)])
            out[f"ep{ep}/dataset_values"] = pol.dataset.values.numpy().copy()
        hp = dict(B=c["B"], epochs=c["epochs"], perm_seed=perm_seed, lr=pol.lr(),
                  entropy_weight=pol.entropy_weight(), surr_clip=pol.surr_clip,
                  vf_clip=np.nan if pol.vf_clip is None else pol.vf_clip,
                  gradient_clip=np.nan if pol.gradient_clip is None else pol.gradient_clip,
                  kl_loss_weight=pol.kl_loss_weight, use_huber_loss=bool(pol.use_huber_loss),
                  normalize_adv=bool(ppo.normalize_adv), normalize_values=bool(ppo.normalize_values),
                  activation=np.array(c["net"]["act"]), gamma=pol.gamma, lambd=pol.lambd,
                  dist_range=c["net"].get("dist_range", 1.0),
                  actor_hidden=c["net"]["actor_hidden"], critic_hidden=c["net"]["critic_hidden"],
                  depth=c["net"].get("depth", 3))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **rollout_inputs(ro),
                            **{"ds_" + k: v for k, v in seg.items()}, **out,
                            **{"hp_" + k: v for k, v in hp.items()})
        print(name, "N =", len(seg["advantages"]), {k: out[k] for k in out if k.endswith("status")})


def gen_rollout_actions():
    """Rollout-time inference: PPOPolicy.get_rollout_actions / get_critic_values of the unmodified
    reference (policies/ppo_policy.py:729-794, 1057-1071) on seeded observations, with the global CPU generator seeded
    right before the call (the sampling protocol the device path has to reproduce)."""
    cases = {
        "act_gauss": dict(ro=dict(seed=41, T=4, E=8, obs_dim=11, act_dim=3, obs_scale=False),
                          net=dict(act="tanh", actor_hidden=32, critic_hidden=24)),
        "act_gauss_range": dict(ro=dict(seed=42, T=4, E=5, obs_dim=6, act_dim=2, obs_scale=False),      # < 16 samples
                                net=dict(act="leaky_relu", actor_hidden=16, critic_hidden=16, dist_range=0.4)),
        "act_cat": dict(ro=dict(seed=43, T=4, E=16, obs_dim=7, n_discrete=5, obs_scale=False),
                        net=dict(act="relu", actor_hidden=32, critic_hidden=32, depth=2)),
    }
    for name, c in cases.items():
        ro = make_rollout(**c["ro"])
        torch.manual_seed(1000 + c["ro"]["seed"])
        pol = build_policy(ro, **c["net"])
        out = {}
        out.update(snapshot_net("init/actor", pol.actor, pol.actor_optim))
        out.update(snapshot_net("init/critic", pol.critic, pol.critic_optim))
        a = ro.agents[0]
        for step in range(2):                       # two consecutive calls: the generator state carries over
            obs = ro.obs[a][step].astype(np.float32)
            if step == 0:
                torch.manual_seed(5000 + c["ro"]["seed"])
            raw, act, lp = pol.get_rollout_actions(obs)
            with torch.no_grad():
                val = pol.get_critic_values(torch.tensor(ro.critic_obs[a][step].astype(np.float32)))
            out[f"s{step}/obs"] = obs
            out[f"s{step}/critic_obs"] = ro.critic_obs[a][step].astype(np.float32)
            out[f"s{step}/raw_action"] = np.asarray(raw).copy()
            out[f"s{step}/action"] = np.asarray(act).copy()
            out[f"s{step}/log_prob"] = lp.numpy().copy()
            out[f"s{step}/value"] = val.numpy().copy()
        hp = dict(sample_seed=5000 + c["ro"]["seed"], activation=np.array(c["net"]["act"]),
                  dist_range=c["net"].get("dist_range", 1.0), actor_hidden=c["net"]["actor_hidden"],
                  critic_hidden=c["net"]["critic_hidden"], depth=c["net"].get("depth", 3),
                  n_discrete=ro.n_discrete, act_dim=ro.act_dim, obs_dim=ro.obs_dim, E=ro.E)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out, **{"hp_" + k: v for k, v in hp.items()})
        print(name, {k: out[k].shape for k in out if k.startswith("s0/")})


def gen_baseline_shapes():
    """Full updates at the network / minibatch shapes of BASELINE.json's configs, through the
    unmodified reference: C1 CartPole (baselines/gymnasium/cart_pole.py:30-65), C3 LunarLanderContinuous
    (baselines/gymnasium/lunar_lander_continuous.py:33-68), C4 Humanoid (baselines/gymnasium/humanoid.py:46-76, both
    nets 256-wide as in the bench) and C5 MPE simple_spread MAPPO (baselines/pettingzoo/mpe_simple_spread.py:29-104).
    The synthetic observations are NOT stored (they are regenerated from `in_make_rollout_json` by make_rollout, which is
    deterministic); what the reference's networks produced on them (values, actions, log-probs) is stored.  Parameters
    are stored after the LAST epoch only, the status scalars and drawn indices for every epoch."""
    import json
    cases = {
        "shape_c1": dict(ro=dict(seed=51, T=256, E=1, obs_dim=4, n_discrete=2, max_ts_per_ep=32, p_term=0.03, p_trunc=0.0,
                                 obs_scale=False),
                         net=dict(act="leaky_relu", actor_hidden=128, critic_hidden=128), pol=dict(lr=2e-3), B=256, epochs=10),
        "shape_c3": dict(ro=dict(seed=53, T=32, E=64, obs_dim=8, act_dim=2, max_ts_per_ep=32, p_term=0.01, p_trunc=0.01,
                                 obs_scale=False),
                         net=dict(act="leaky_relu", actor_hidden=64, critic_hidden=256), pol=dict(lr=3e-4), B=512, epochs=2),
        "shape_c4": dict(ro=dict(seed=54, T=16, E=64, obs_dim=376, act_dim=17, max_ts_per_ep=16, p_term=0.01, p_trunc=0.01,
                                 obs_scale=False),
                         net=dict(act="tanh", actor_hidden=256, critic_hidden=256, dist_range=0.4), pol=dict(lr=1e-4),
                         B=512, epochs=2),
        "shape_c5": dict(ro=dict(seed=55, T=32, E=8, agents=("a0", "a1", "a2"), obs_dim=18, critic_obs_dim=54, n_discrete=5,
                                 max_ts_per_ep=64, p_term=0.01, p_trunc=0.01, obs_scale=False, shared_critic_obs=True),
                         net=dict(act="leaky_relu", actor_hidden=128, critic_hidden=256), pol=dict(lr=3e-4), B=128, epochs=2),
    }
    for name, c in cases.items():
        ro = make_rollout(**c["ro"])
        torch.manual_seed(1000 + c["ro"]["seed"])
        pol = build_policy(ro, **c["net"], **c["pol"])
        fill_policy_outputs(pol, ro, seed=2000 + c["ro"]["seed"])
        seg = run_reference_rollout(pol, ro)
        pid = "pol"
        ppo = object.__new__(PPO)
        ppo.policies = {pid: pol}
        ppo.normalize_values = True
        ppo.normalize_adv = True
        ppo.value_normalizers = {pid: RunningStatNormalizer(pid + "-value_normalizer", torch.device("cpu"))}
        ppo.status_dict = {pid: {}, "global status": {"iteration": 0, "timesteps": 0}}
        ppo.user_huber_loss = pol.use_huber_loss
        out = {}
        for prefix, net in (("init/actor", pol.actor), ("init/critic", pol.critic)):
            for k, v in net.state_dict().items():
                out[f"{prefix}/param/{k}"] = v.detach().numpy().copy()
        rec = _RecordingDataset(pol.dataset)
        loader = DataLoader(rec, batch_size=c["B"], shuffle=True)
        perm_seed = 3000 + c["ro"]["seed"]
        torch.manual_seed(perm_seed)
        pol.train()
        for ep in range(c["epochs"]):
            rec.drawn = []
            ppo._ppo_batch_train(loader, pid)
            out[f"ep{ep}/batch_idxs"] = np.array(rec.drawn, dtype=np.int64)
            sd = ppo.status_dict[pid]
            out[f"ep{ep}/status"] = np.array([sd["actor loss"], sd["critic loss"], sd["kl avg"],
                                              sd["weighted entropy"]], dtype=np.float64)
            rs = ppo.value_normalizers[pid].running_stats
            out[f"ep{ep}/vn"] = np.array([float(rs.mean), float(rs.variance), float(rs.count)])
        last = c["epochs"] - 1
        for prefix, net in ((f"ep{last}/actor", pol.actor), (f"ep{last}/critic", pol.critic)):
            for k, v in net.state_dict().items():
                out[f"{prefix}/param/{k}"] = v.detach().numpy().copy()
        out[f"ep{last}/dataset_values"] = pol.dataset.values.numpy().copy()
        hp = dict(B=c["B"], epochs=c["epochs"], perm_seed=perm_seed, lr=pol.lr(),
                  entropy_weight=pol.entropy_weight(), surr_clip=pol.surr_clip,
                  vf_clip=np.nan if pol.vf_clip is None else pol.vf_clip,
                  gradient_clip=np.nan if pol.gradient_clip is None else pol.gradient_clip,
                  kl_loss_weight=pol.kl_loss_weight, use_huber_loss=bool(pol.use_huber_loss),
                  normalize_adv=True, normalize_values=True,
                  activation=np.array(c["net"]["act"]), gamma=pol.gamma, lambd=pol.lambd,
                  dist_range=c["net"].get("dist_range", 1.0),
                  actor_hidden=c["net"]["actor_hidden"], critic_hidden=c["net"]["critic_hidden"], depth=3)
        ins = rollout_inputs(ro)
        for k in list(ins):                                   # regenerated from the seed at test time
            if k.split("/")[0] in ("in_obs", "in_next_obs", "in_critic_obs", "in_rewards"):
                del ins[k]
        ro_kw = dict(c["ro"])
        if "agents" in ro_kw:
            ro_kw["agents"] = list(ro_kw["agents"])
        ins["in_make_rollout_json"] = np.array(json.dumps(ro_kw))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **ins,
                            ds_advantages=seg["advantages"], ds_rewards_to_go=seg["rewards_to_go"], ds_ep_lens=seg["ep_lens"],
                            **out, **{"hp_" + k: v for k, v in hp.items()})
        print(name, "N =", len(seg["advantages"]), {k: out[k] for k in out if k.endswith("status")})


class ReplayEnv:
    """A vectorised multi-agent environment stand-in (the interface of the reference's VectorizedEnv as its filter
    wrappers see it, environments/ppo_env_wrappers.py:24-147) that replays pre-generated raw step results."""

    def __init__(self, data):
        self.data = data
        self.agent_ids = tuple(data["agents"])
        self.observation_space = {a: sp.Box(-np.inf, np.inf, (data["obs_dim"],)) for a in self.agent_ids}
        self.critic_observation_space = {a: sp.Box(-np.inf, np.inf, (data["critic_dim"],)) for a in self.agent_ids}
        self.action_space = {a: sp.Discrete(2) for a in self.agent_ids}
        self.null_actions = {a: 0 for a in self.agent_ids}
        self.t = 0

    def get_batch_size(self):
        return self.data["E"]

    def reset(self):
        self.t = 0
        return ({a: self.data[f"obs/{a}"][0].copy() for a in self.agent_ids},
                {a: self.data[f"critic_obs/{a}"][0].copy() for a in self.agent_ids})

    def step(self, action):
        self.t += 1
        t, d, E = self.t, self.data, self.data["E"]
        info = {a: np.array([{} for _ in range(E)]) for a in self.agent_ids}
        return ({a: d[f"obs/{a}"][t].copy() for a in self.agent_ids},
                {a: d[f"critic_obs/{a}"][t].copy() for a in self.agent_ids},
                {a: d[f"reward/{a}"][t].copy() for a in self.agent_ids},
                {a: d[f"terminated/{a}"][t].copy() for a in self.agent_ids},
                {a: d[f"truncated/{a}"][t].copy() for a in self.agent_ids}, info)


def make_filter_data(seed, T, E, agents, obs_dim, critic_dim):
    rng = np.random.default_rng(seed)
    d = dict(agents=list(agents), E=E, T=T, obs_dim=obs_dim, critic_dim=critic_dim)
    for i, a in enumerate(agents):
        scale = rng.uniform(0.1, 20.0, obs_dim).astype(np.float32)
        shift = rng.uniform(-5.0, 5.0, obs_dim).astype(np.float32)
        d[f"obs/{a}"] = (rng.standard_normal((T + 1, E, obs_dim)).astype(np.float32) * scale + shift).astype(np.float32)
        d[f"critic_obs/{a}"] = (rng.standard_normal((T + 1, E, critic_dim)) * 3.0 + i).astype(np.float32)
        d[f"reward/{a}"] = (rng.standard_normal((T + 1, E)) * 4.0 + 1.0).astype(np.float32)
        d[f"terminated/{a}"] = rng.random((T + 1, E)) < 0.08
        d[f"truncated/{a}"] = rng.random((T + 1, E)) < 0.05
    return d


def gen_filters():
    """The reference's wrapper stack in its own order (environments/wrapper_utils.py:82-112): ObservationNormalizer ->
    ObservationClipper -> RewardNormalizer -> RewardClipper, driven for T steps over a replayed vectorised environment."""
    from ppo_and_friends.environments.filter_wrappers import (ObservationClipper, ObservationNormalizer, RewardClipper,
                                                                RewardNormalizer)
    cases = {"filt_multi": dict(seed=61, T=12, E=6, agents=("a0", "a1"), obs_dim=5, critic_dim=10),
             "filt_single": dict(seed=62, T=20, E=16, agents=("agent0",), obs_dim=24, critic_dim=24)}
    for name, c in cases.items():
        data = make_filter_data(**c)
        env = ReplayEnv(data)
        on = ObservationNormalizer(env)
        oc = ObservationClipper(on, clip_range=(-3.0, 3.0))
        rn = RewardNormalizer(oc, gamma=0.97)
        rc = RewardClipper(rn, clip_range=(-2.0, 2.0))
        out = {}
        obs, cobs = rc.reset()
        for a in env.agent_ids:
            out[f"t0/obs/{a}"], out[f"t0/critic_obs/{a}"] = np.asarray(obs[a]), np.asarray(cobs[a])
        for t in range(1, c["T"] + 1):
            obs, cobs, rew, term, trunc, info = rc.step(None)
            for a in env.agent_ids:
                out[f"t{t}/obs/{a}"], out[f"t{t}/critic_obs/{a}"] = np.asarray(obs[a]), np.asarray(cobs[a])
                out[f"t{t}/reward/{a}"] = np.asarray(rew[a])
                out[f"t{t}/natural/{a}"] = np.array([i["natural reward"] for i in info[a]])
        for a in env.agent_ids:
            for tag, rs in (("actor", on.actor_running_stats[a]), ("critic", on.critic_running_stats[a]),
                            ("reward", rn.running_stats[a])):
                out[f"final/{tag}/{a}/mean"] = np.asarray(rs.mean, dtype=np.float64)
                out[f"final/{tag}/{a}/variance"] = np.asarray(rs.variance, dtype=np.float64)
                out[f"final/{tag}/{a}/count"] = np.float64(rs.count)
            out[f"final/running_reward/{a}"] = np.asarray(rn.running_reward[a], dtype=np.float64)
        ins = {("in_" + k): (np.array(v) if isinstance(v, list) else v) for k, v in data.items()}
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **ins, **out, hp_gamma=0.97, hp_obs_clip=np.array([-3.0, 3.0]),
                            hp_reward_clip=np.array([-2.0, 2.0]))
        print(name, {k: v for k, v in out.items() if k.startswith("final/reward") and k.endswith("variance")})


if __name__ == "__main__":
    import sys
    if len(sys.argv) > 1 and sys.argv[1] == "actions":       # only the rollout-action fixtures (added later)
        gen_rollout_actions()
    elif len(sys.argv) > 1 and sys.argv[1] == "filters":     # only the normaliser / clipper stack fixtures (round 2)
        gen_filters()
    elif len(sys.argv) > 1 and sys.argv[1] == "shapes":      # only the BASELINE-shape update fixtures (added in round 2)
        gen_baseline_shapes()
    elif len(sys.argv) > 1 and sys.argv[1] == "edges":       # only the edge-shape segment fixtures (added later)
        gen_segments(only=("seg_len1", "seg_open", "seg_max1"))
    else:
        gen_segments()
        gen_stats()
        gen_updates()
        gen_rollout_actions()
        gen_baseline_shapes()
        gen_filters()
