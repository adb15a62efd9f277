"""
The normaliser / clipper wrapper stack against goldens recorded from the UNMODIFIED reference
wrappers (tests/golden/make_golden.py::gen_filters: ObservationNormalizer -> ObservationClipper -> RewardNormalizer ->
RewardClipper over a replayed vectorised environment).  CPU: the numpy oracle reproduces the reference step by step
(incl. the RewardNormalizer's E sequential statistic updates per step).  GPU: the device-resident stack.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden

CASES = ["filt_multi", "filt_single"]


class _Replay:
    def __init__(self, g):
        self.g = g
        self.agent_ids = tuple(str(a) for a in g["in_agents"])
        self.E, self.T = int(g["in_E"]), int(g["in_T"])

        class _S:
            def __init__(self, shape):
                self.shape = shape
        self.observation_space = {a: _S((int(g["in_obs_dim"]),)) for a in self.agent_ids}
        self.critic_observation_space = {a: _S((int(g["in_critic_dim"]),)) for a in self.agent_ids}
        self.t = 0

    def get_batch_size(self):
        return self.E

    def raw(self, t):
        g = self.g
        return ({a: g[f"in_obs/{a}"][t].copy() for a in self.agent_ids}, {a: g[f"in_critic_obs/{a}"][t].copy() for a in self.agent_ids},
                {a: g[f"in_reward/{a}"][t].copy() for a in self.agent_ids}, {a: g[f"in_terminated/{a}"][t].copy() for a in self.agent_ids},
                {a: g[f"in_truncated/{a}"][t].copy() for a in self.agent_ids})

    def reset(self):
        self.t = 0
        o, c, _, _, _ = self.raw(0)
        return o, c

    def step(self, action):
        self.t += 1
        o, c, r, te, tr = self.raw(self.t)
        info = {a: [dict() for _ in range(self.E)] for a in self.agent_ids}
        return o, c, r, te, tr, info


@pytest.mark.parametrize("name", CASES)
def test_oracle_filter_stack_matches_reference(name):
    from oracle.filters import OracleFilterStack
    g = load_golden(name)
    env = _Replay(g)
    st = OracleFilterStack(env.agent_ids, int(g["in_obs_dim"]), int(g["in_critic_dim"]), env.E, gamma=float(g["hp_gamma"]),
                           obs_clip=tuple(g["hp_obs_clip"]), reward_clip=tuple(g["hp_reward_clip"]))
    for t in range(env.T + 1):
        o, c, r, te, tr = env.raw(t)
        fo, fc = st.filter_obs(o, c)
        for a in env.agent_ids:
            np.testing.assert_allclose(fo[a], g[f"t{t}/obs/{a}"], rtol=1e-6, atol=1e-6)
            np.testing.assert_allclose(fc[a], g[f"t{t}/critic_obs/{a}"], rtol=1e-6, atol=1e-6)
        if t > 0:
            fr = st.filter_reward(r, te, tr)
            for a in env.agent_ids:
                np.testing.assert_allclose(fr[a], g[f"t{t}/reward/{a}"], rtol=1e-6, atol=1e-7)
    for a in env.agent_ids:
        np.testing.assert_allclose(st.reward[a].variance, g[f"final/reward/{a}/variance"], rtol=1e-9)
        np.testing.assert_allclose(st.running_reward[a], g[f"final/running_reward/{a}"], rtol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("fused_clip", [False, True])
def test_device_filter_stack_matches_reference(name, fused_clip):
    """Device-resident stack (libppoaf_b200.so kernels) over the same replay: filtered observations, rewards and the
    natural-reward book-keeping per step, running statistics and the running-reward vector at the end.  `fused_clip`
    folds the two clippers into the normalisers' kernels (one pass) — same results."""
    from ppo_and_friends_b200.environments.filter_wrappers import (ObservationClipper, ObservationNormalizer, RewardClipper,
                                                                     RewardNormalizer)
    g = load_golden(name)
    env = _Replay(g)
    oc_rng, rc_rng = tuple(float(v) for v in g["hp_obs_clip"]), tuple(float(v) for v in g["hp_reward_clip"])
    if fused_clip:
        on = ObservationNormalizer(env, clip_range=oc_rng)
        rn = RewardNormalizer(on, gamma=float(g["hp_gamma"]), clip_range=rc_rng)
        top = rn
    else:
        on = ObservationNormalizer(env)
        oc = ObservationClipper(on, clip_range=oc_rng)
        rn = RewardNormalizer(oc, gamma=float(g["hp_gamma"]))
        top = RewardClipper(rn, clip_range=rc_rng)
    obs, cobs = top.reset()
    for t in range(env.T + 1):
        if t > 0:
            obs, cobs, rew, term, trunc, info = top.step(None)
        for a in env.agent_ids:
            assert obs[a].is_cuda and cobs[a].is_cuda
            np.testing.assert_allclose(obs[a].cpu().numpy(), g[f"t{t}/obs/{a}"], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(cobs[a].cpu().numpy(), g[f"t{t}/critic_obs/{a}"], rtol=1e-5, atol=1e-5)
            if t > 0:
                np.testing.assert_allclose(rew[a].cpu().numpy(), g[f"t{t}/reward/{a}"], rtol=1e-5, atol=1e-6)
                nat = np.array([i["natural reward"] for i in info[a]], dtype=np.float64)
                np.testing.assert_array_equal(nat.astype(np.float32), g[f"t{t}/natural/{a}"].astype(np.float32))
    for a in env.agent_ids:
        for tag, rs in (("actor", on.actor_running_stats[a]), ("critic", on.critic_running_stats[a]), ("reward", rn.running_stats[a])):
            np.testing.assert_allclose(rs.mean, g[f"final/{tag}/{a}/mean"], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(rs.variance, g[f"final/{tag}/{a}/variance"], rtol=1e-5)
            assert abs(rs.count - float(g[f"final/{tag}/{a}/count"])) < 1e-6
        np.testing.assert_allclose(rn.running_reward[a].cpu().numpy(), g[f"final/running_reward/{a}"], rtol=1e-12)


@pytest.mark.gpu
def test_two_rank_filter_stack_pools_ranks_like_the_reference():
    """R = 2: every statistic update pools the ranks (triples exchanged instead of the reference's pickled raw batches,
    utils/stats.py:47-53), including the reward normaliser's E sequential updates per step; against the numpy oracle
    replaying both ranks together (tests/multi_gpu_filter_worker.py)."""
    import os, subprocess, sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    port = 29400 + os.getpid() % 500
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          os.path.join(root, "tests", "multi_gpu_filter_worker.py")], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "MULTI_GPU_FILTER_CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
