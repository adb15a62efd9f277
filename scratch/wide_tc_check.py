# one-off: how far is the tcgen05 backend from the oracle on the wide-network / 640-row case?
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import numpy as np, torch
from helpers import make_policy, run_device_rollout
from oracle.update import OracleUpdater
from ppo_and_friends_b200 import ops
from ppo_and_friends_b200.ppo import PPOUpdateState, _Loader, ppo_batch_train
from ppo_and_friends_b200.synthetic import make_rollout
for backend in ("ffma", "tcgen05"):
    ops.set_gemm_backend(backend)
    ro = make_rollout(seed=78, T=20, E=64, obs_dim=24, act_dim=5, max_ts_per_ep=10, obs_scale=False)
    torch.manual_seed(6)
    pol = make_policy(ro, act="leaky_relu", actor_hidden=320, critic_hidden=576, depth=2, lr=3e-4)
    with torch.no_grad():
        for a in ro.agents:
            obs = ro.obs[a].reshape(-1, 24)
            mu = pol.actor(obs).cpu().numpy()
            sd = np.maximum(np.log1p(np.exp(-0.5)), 0.01)
            raw = (mu + sd * np.random.default_rng(2).standard_normal(mu.shape)).astype(np.float32)
            ro.raw_actions[a] = raw.reshape(ro.T, ro.E, 5)
            _, lp, _ = pol.evaluate(ro.critic_obs[a].reshape(-1, 24), obs, raw)
            ro.log_probs[a] = lp.cpu().numpy().reshape(ro.T, ro.E)
            ro.values[a] = pol.critic(ro.critic_obs[a].reshape(-1, 24)).cpu().numpy().reshape(ro.T, ro.E)
    ds = run_device_rollout(pol, ro)
    host = {k: getattr(ds, k).cpu().numpy().copy() for k in ("critic_observations", "observations", "raw_actions",
                                                               "advantages", "log_probs", "rewards_to_go", "values")}
    oracle = OracleUpdater({k: v.cpu().numpy() for k, v in pol.actor.state_dict().items()},
                           {k: v.cpu().numpy() for k, v in pol.critic.state_dict().items()}, "leaky_relu", False, lr=3e-4)
    init = {n: {k: v.cpu().numpy().copy() for k, v in o.state_dict().items()} for n, o in (("actor", pol.actor), ("critic", pol.critic))}
    state = PPOUpdateState({"pol": pol}, batch_size=640, epochs_per_iter=1)
    torch.manual_seed(12)
    ppo_batch_train(state, _Loader(ds, 640), "pol")
    perm = pol._engine._perm_dev.cpu().numpy()
    oracle.batch_train([host], [perm], 640)
    ref = oracle.state()
    print("backend", backend)
    for net, obj in (("actor", pol.actor), ("critic", pol.critic)):
        for k, v in obj.state_dict().items():
            r = ref[f"{net}/param/{k}"]; g = v.cpu().numpy()
            err = np.abs(g - r); tol = 1e-4 * np.abs(r) + 1e-6
            upd = np.abs(r - init[net][k])
            bad = err > tol
            print(f"  {net}/{k:28s} max err {err.max():.2e} (max update {upd.max():.2e})  over tol: {int(bad.sum())}/{bad.size}  worst err/tol {np.max(err / tol):.1f}")
