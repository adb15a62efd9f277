"""
The UNMODIFIED reference timed on the host cores (bench.py `--impl reference` and `cpu_baseline`).

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by ppo_and_friends_b200).  R worker processes (one per
"rank", like `mpirun -n R ppoaf train ...`), rendezvous over gloo on 127.0.0.1, each running on its own synthetic
rollout shard of the bench workload:

    PPOPolicy.add_episode_info / end_episodes   (reference policies/ppo_policy.py:545-712)   untimed caller side ...
      ... except EpisodeInfo.end_episode        (utils/episode_info.py:419-465: GAE + reward-to-go)   TIMED
    PPOPolicy.finalize_dataset -> PPODataset.build  (utils/episode_info.py:745-914)                     TIMED
    PPO._ppo_batch_train through a real DataLoader  (ppo.py:2274-2485, 2181-2184), with mpi_avg_gradients
      (utils/mpi_utils.py:89-111) and the RunningMeanStd allgather (utils/stats.py:47-50) carried by gloo
      through the mpi4py stand-in of tests/golden/ref_harness.py                                       TIMED

`epochs_timed` of the workload's epochs are run per step and the update time is scaled to the full epoch count
(stated in `sample`); torch threads per rank follow the reference's set_torch_threads (utils/mpi_utils.py:37-48).
"""
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, w, steps, warmup, epochs_timed, total_threads, out_q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    # under torchrun the parent's environment carries the elastic agent's settings; TORCHELASTIC_USE_AGENT_STORE in particular
    # makes every rank (rank 0 included) CONNECT to MASTER_PORT instead of serving the store there - with our own port nobody
    # would serve it and the rendezvous would wait for its timeout
    for k in list(os.environ):
        if k.startswith("TORCHELASTIC_") or k in ("LOCAL_RANK", "GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE", "GROUP_WORLD_SIZE",
                                                  "ROLE_WORLD_SIZE", "ROLE_NAME", "OMP_NUM_THREADS"):
            os.environ.pop(k, None)
    import torch.distributed as dist
    torch.set_num_threads(max(int(total_threads), 1))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden as mg                                   # installs the import stubs, imports the reference
    from ppo_and_friends.utils import mpi_utils as ref_mpi
    from ppo_and_friends.utils.episode_info import EpisodeInfo
    from ppo_and_friends.utils.misc import RunningStatNormalizer
    from ppo_and_friends.ppo import PPO
    from torch.utils.data import DataLoader
    from ppo_and_friends_b200.synthetic import make_rollout

    ref_mpi.set_torch_threads()                                # the reference's own thread policy (threads / num_procs)
    agents = tuple(f"agent_{i}" for i in range(w["agents"]))
    ro = make_rollout(seed=1234 + rank, T=w["ts"], E=w["E"], agents=agents, obs_dim=w["Do"], critic_obs_dim=w["Dc"],
                      act_dim=w["Da"], n_discrete=w["n_disc"], max_ts_per_ep=w["max_ts_per_ep"], obs_scale=False,
                      shared_critic_obs=w["shared_critic"])
    torch.manual_seed(4321)
    import io, contextlib
    with contextlib.redirect_stdout(io.StringIO()):
        pol = mg.build_policy(ro, act=w["act"], actor_hidden=w["actor_hidden"], critic_hidden=w["critic_hidden"],
                              dist_range=w["dist_range"], lr=w["lr"], target_kl=float("inf"))
    mg.fill_policy_outputs(pol, ro, seed=99 + rank)
    for net in (pol.actor, pol.critic):
        ref_mpi.broadcast_model_parameters(net)

    clock = [0.0]
    orig_end = EpisodeInfo.end_episode

    def timed_end(self, *a, **k):                              # timing wrapper only: the reference function runs unmodified
        t0 = time.perf_counter()
        r = orig_end(self, *a, **k)
        clock[0] += time.perf_counter() - t0
        return r
    EpisodeInfo.end_episode = timed_end

    pid = "pol"
    ppo = object.__new__(PPO)
    ppo.policies = {pid: pol}
    ppo.normalize_values, ppo.normalize_adv = True, True
    ppo.value_normalizers = {pid: RunningStatNormalizer(pid + "-value_normalizer", torch.device("cpu"))}
    ppo.status_dict = {pid: {}, "global status": {"iteration": 0, "timesteps": 0}}
    ppo.user_huber_loss = pol.use_huber_loss
    pol.train()

    t_adv = t_build = t_upd = t_rec = 0.0
    for it in range(warmup + steps):
        clock[0] = 0.0
        t0 = time.perf_counter()
        pol.initialize_dataset()
        pol.initialize_episodes(ro.E, ppo.status_dict)
        lp_backup = dict(ro.log_probs)
        for a in ro.agents:
            ro.log_probs[a] = torch.tensor(lp_backup[a])
        from ppo_and_friends_b200.synthetic import replay_rollout
        replay_rollout(lambda a: pol, ro, to_bootstrap=lambda x: torch.tensor(x))
        ro.log_probs = lp_backup
        rec = time.perf_counter() - t0 - clock[0]
        t0 = time.perf_counter()
        pol.finalize_dataset()
        build = time.perf_counter() - t0
        loader = DataLoader(pol.dataset, batch_size=w["B"], shuffle=True)
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(epochs_timed):
            ppo._ppo_batch_train(loader, pid)
        upd = time.perf_counter() - t0
        if it >= warmup:
            t_adv += clock[0]; t_build += build; t_upd += upd; t_rec += rec
    per_step = (t_adv + t_build + t_upd * (w["epochs"] / epochs_timed)) / steps
    res = dict(rank=rank, per_step_s=per_step, t_end_episode=t_adv / steps, t_build=t_build / steps,
               t_update_timed=t_upd / steps, t_record_untimed=t_rec / steps, threads=torch.get_num_threads())
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, res)
        dist.barrier()
        dist.destroy_process_group()
    else:
        gathered = [res]
    if rank == 0:
        out_q.put(gathered)


def run(w, n_ranks=1, steps=1, warmup=0, epochs_timed=1, port=None):
    """Returns the JSON-able record of one reference measurement (value = whole-job env-steps/s over n_ranks ranks)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total_threads = os.cpu_count() or torch.get_num_threads()
    try:
        total_threads = len(os.sched_getaffinity(0))
    except Exception:
        pass
    port = port or (29650 + os.getpid() % 200)
    procs = [ctx.Process(target=_worker, args=(r, n_ranks, port, w, steps, warmup, epochs_timed, total_threads, q))
             for r in range(n_ranks)]
    for p in procs:
        p.start()
    import queue as _queue
    gathered = None
    deadline = time.time() + 3600
    while gathered is None:
        try:
            gathered = q.get(timeout=2)
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs) or time.time() > deadline:
                for p in procs:
                    if p.is_alive():
                        p.terminate()
                raise RuntimeError("a reference worker process died (exit codes %s)" % [p.exitcode for p in procs])
    for p in procs:
        p.join(timeout=60)
    slowest = max(g["per_step_s"] for g in gathered)
    env_steps = w["ts"] * w["E"] * n_ranks
    g0 = gathered[0]
    return dict(value=env_steps / slowest, unit="env-steps/s", cores=total_threads, kind="reference",
                ranks=n_ranks, threads_per_rank=g0["threads"], per_step_s=slowest,
                sample=(f"unmodified reference (oracle/_ref), {n_ranks} rank(s) x {g0['threads']} torch threads over gloo: "
                        f"EpisodeInfo.end_episode + PPODataset.build of the full shard ({w['ts']}x{w['E']} steps) and "
                        f"{epochs_timed} of {w['epochs']} epochs of PPO._ppo_batch_train through DataLoader per step "
                        f"(update time scaled x{w['epochs'] / epochs_timed:g}); rank 0: end_episode {g0['t_end_episode']:.2f}s "
                        f"build {g0['t_build']:.2f}s update(timed) {g0['t_update_timed']:.2f}s; add_episode_info caller side "
                        f"{g0['t_record_untimed']:.2f}s not counted"))
