"""
torchrun worker: R ranks drive the device-resident normaliser / clipper stack (environments/filter_wrappers.py) over
their own replayed environments; rank 0 replays all ranks together through the numpy oracle, in which every statistic
update pools the ranks' batches in rank order exactly like the reference's allgather + concatenate
(utils/stats.py:47-53) - including the RewardNormalizer's E sequential updates per step, where the pooled batch at
element e is the concatenation of every rank's partially updated running-reward vector - and compares.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class Replay:
    def __init__(self, seed, T, E, agents, obs_dim, critic_dim):
        rng = np.random.default_rng(seed)
        self.agent_ids, self.E, self.T = tuple(agents), E, T

        class _S:
            def __init__(self, shape):
                self.shape = shape
        self.observation_space = {a: _S((obs_dim,)) for a in agents}
        self.critic_observation_space = {a: _S((critic_dim,)) for a in agents}
        self.d = {}
        for i, a in enumerate(agents):
            self.d[f"obs/{a}"] = (rng.standard_normal((T + 1, E, obs_dim)) * (1 + i) + seed % 7).astype(np.float32)
            self.d[f"critic_obs/{a}"] = (rng.standard_normal((T + 1, E, critic_dim)) * 2.0).astype(np.float32)
            self.d[f"reward/{a}"] = (rng.standard_normal((T + 1, E)) * 3.0 + 0.5).astype(np.float32)
            self.d[f"terminated/{a}"] = rng.random((T + 1, E)) < 0.1
            self.d[f"truncated/{a}"] = rng.random((T + 1, E)) < 0.05
        self.t = 0

    def get_batch_size(self):
        return self.E

    def raw(self, t):
        return tuple({a: self.d[f"{k}/{a}"][t].copy() for a in self.agent_ids}
                     for k in ("obs", "critic_obs", "reward", "terminated", "truncated"))

    def reset(self):
        self.t = 0
        o, c, _, _, _ = self.raw(0)
        return o, c

    def step(self, action):
        self.t += 1
        o, c, r, te, tr = self.raw(self.t)
        return o, c, r, te, tr, {a: [dict() for _ in range(self.E)] for a in self.agent_ids}


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from oracle.stats import OracleRunningMeanStd
    from ppo_and_friends_b200.environments.filter_wrappers import ObservationNormalizer, RewardNormalizer
    T, E, agents, Do, Dc, gamma = 6, 5, ("a0", "a1"), 7, 9, 0.95
    obs_clip, rew_clip = (-2.5, 2.5), (-1.5, 1.5)
    env = Replay(900 + rank, T, E, agents, Do, Dc)
    on = ObservationNormalizer(env, clip_range=obs_clip, device=f"cuda:{local}")
    top = RewardNormalizer(on, gamma=gamma, clip_range=rew_clip, device=f"cuda:{local}")
    outs = []
    obs, cobs = top.reset()
    outs.append(dict(obs={a: obs[a].cpu().numpy() for a in agents}, cobs={a: cobs[a].cpu().numpy() for a in agents}))
    for t in range(1, T + 1):
        obs, cobs, rew, _, _, _ = top.step(None)
        outs.append(dict(obs={a: obs[a].cpu().numpy() for a in agents}, cobs={a: cobs[a].cpu().numpy() for a in agents},
                         rew={a: rew[a].cpu().numpy() for a in agents}))
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(data=env.d, outs=outs))
    ok = True
    if rank == 0:
        actor = {a: OracleRunningMeanStd(shape=(Do,)) for a in agents}
        critic = {a: OracleRunningMeanStd(shape=(Dc,)) for a in agents}
        reward = {a: OracleRunningMeanStd(shape=()) for a in agents}
        rr = [{a: np.zeros(E) for a in agents} for _ in range(world)]
        worst = 0.0
        for t in range(T + 1):
            for a in agents:
                for tab, key, okey in ((actor, "obs", "obs"), (critic, "critic_obs", "cobs")):
                    batches = [gathered[r]["data"][f"{key}/{a}"][t] for r in range(world)]
                    tab[a].update(batches[0], other_ranks=batches[1:])
                    for r in range(world):
                        ref = np.clip((batches[r] - tab[a].mean) / np.sqrt(tab[a].variance + 1e-8), *obs_clip)
                        worst = max(worst, float(np.max(np.abs(gathered[r]["outs"][t][okey][a] - ref))))
            if t == 0:
                continue
            for a in agents:
                rews = [gathered[r]["data"][f"reward/{a}"][t] for r in range(world)]
                for e in range(E):
                    for r in range(world):
                        rr[r][a][e] = rr[r][a][e] * gamma + rews[r][e]
                    reward[a].update(rr[0][a], other_ranks=[rr[r][a] for r in range(1, world)])
                for r in range(world):
                    done = np.logical_or(gathered[r]["data"][f"terminated/{a}"][t], gathered[r]["data"][f"truncated/{a}"][t])
                    rr[r][a][done] = 0.0
                    ref = np.clip(rews[r] / np.sqrt(reward[a].variance + 1e-8), *rew_clip)
                    worst = max(worst, float(np.max(np.abs(gathered[r]["outs"][t]["rew"][a] - ref))))
        print("worst |device - oracle| over all ranks, steps, agents:", worst)
        ok = worst < 2e-5
        for a in agents:
            v = float(top.running_stats[a].variance)
            print(a, "reward variance device", v, "oracle", float(reward[a].variance))
            ok &= abs(v - float(reward[a].variance)) <= 1e-5 * abs(float(reward[a].variance))
    flag = torch.tensor([int(ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_FILTER_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
