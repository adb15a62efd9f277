"""
torchrun worker: R ranks run one PPO epoch with the NCCL gradient all-reduce and the all-gathered
value-normaliser triples; rank 0 replays the same thing through the multi-rank CPU oracle
(gradients summed over ranks / R per tensor, value statistics from the concatenated minibatches,
reference utils/mpi_utils.py:89-111, utils/stats.py:47-59) and compares losses and parameters.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stage(msg):
    if os.environ.get("PPOAF_MG_VERBOSE"):
        print(f"[rank {os.environ.get('RANK')}] {msg}", flush=True)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stage("process group up")
    from helpers import make_policy, run_device_rollout
    from ppo_and_friends_b200.ppo import PPOUpdateState, _Loader, ppo_batch_train
    from ppo_and_friends_b200.synthetic import make_rollout

    discrete = os.environ.get("PPOAF_MG_DISCRETE", "0") == "1"
    ro = make_rollout(seed=500 + rank, T=24, E=8, obs_dim=12, act_dim=3, n_discrete=4 if discrete else 0,
                      max_ts_per_ep=8, obs_scale=False)
    torch.manual_seed(1000 + rank)                 # different init per rank: the broadcast must fix that
    pol = make_policy(ro, act="tanh", actor_hidden=32, critic_hidden=48, lr=1e-3, device=f"cuda:{local}")
    stage("policy built")
    p0 = pol.nets.flat_params.clone()
    dist.broadcast(p0, src=0)
    assert torch.equal(p0, pol.nets.flat_params), "parameters were not broadcast from rank 0"
    stage("broadcast verified")
    ds = run_device_rollout(pol, ro)
    stage("dataset built")
    host = {k: getattr(ds, k).cpu().numpy().copy() for k in ("critic_observations", "observations", "raw_actions",
                                                               "advantages", "log_probs", "rewards_to_go", "values")}
    init_a = {k: v.cpu().numpy().copy() for k, v in pol.actor.state_dict().items()}
    init_c = {k: v.cpu().numpy().copy() for k, v in pol.critic.state_dict().items()}
    state = PPOUpdateState({"pol": pol}, batch_size=64, epochs_per_iter=1, device=f"cuda:{local}")
    torch.manual_seed(77 + rank)                    # the reference seeds rank r with seed + r
    ppo_batch_train(state, _Loader(ds, 64), "pol")
    stage("epoch done")
    expect = os.environ.get("PPOAF_MG_EXPECT")     # the exchange the test asked for must be the one that ran
    if expect:
        got = {"NvlsGroup": "nvls", "PeerGroup": "push", "NoneType": "nccl"}[type(pol._engine.peer).__name__]
        assert got == expect, f"expected the {expect} exchange, the engine chose {got}"
    perm = pol._engine._perm_dev.cpu().numpy()
    gathered = [None] * world
    dist.all_gather_object(gathered, dict(host=host, perm=perm))
    sd = state.status_dict["pol"]
    ok = True
    if rank == 0:
        from oracle.update import OracleUpdater
        oracle = OracleUpdater(init_a, init_c, "tanh", discrete, lr=1e-3)
        st = oracle.batch_train([g["host"] for g in gathered], [g["perm"] for g in gathered], 64)
        for k in ("actor loss", "critic loss", "kl avg", "weighted entropy"):
            err = abs(sd[k] - st[k]) / max(abs(st[k]), 1e-3)
            print(f"{k}: device {sd[k]:.8e} oracle {st[k]:.8e} rel {err:.2e}")
            ok &= err < 1e-4
        ref = oracle.state()
        worst = 0.0
        for net, obj in (("actor", pol.actor), ("critic", pol.critic)):
            for k, v in obj.state_dict().items():
                r = ref[f"{net}/param/{k}"]
                e = np.max(np.abs(v.cpu().numpy() - r) / (1e-4 * np.abs(r) + 1e-6))
                worst = max(worst, float(e))
        print("worst parameter error (in units of the 1e-4 rel + 1e-6 abs tolerance):", worst)
        ok &= worst < 1.0
        vs = state.value_normalizers["pol"].running_stats
        print("value stats", float(vs.mean), float(vs.variance), vs.count, "oracle", float(oracle.value_stats.mean),
              float(oracle.value_stats.variance), oracle.value_stats.count)
        ok &= abs(float(vs.mean) - float(oracle.value_stats.mean)) < 1e-5 and abs(vs.count - oracle.value_stats.count) < 1e-6
    # every rank must hold identical parameters afterwards
    mine = pol.nets.flat_params.clone()
    ref0 = mine.clone()
    dist.broadcast(ref0, src=0)
    same = torch.equal(mine, ref0)
    flag = torch.tensor([int(ok and same)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("MULTI_GPU_CHECK", "PASS" if int(flag.item()) == 1 else "FAIL", "world", world)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
