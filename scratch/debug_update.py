import sys, os
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
os.environ["PPOAF_NO_GRAPH"] = "1"
from conftest import load_golden
from test_gpu_parity import policy_from_update_golden
from test_oracle_golden import updater_from_golden, dataset_from_golden
from helpers import run_device_rollout
from ppo_and_friends_b200.ppo import PPOUpdateState, draw_minibatch_permutation
from ppo_and_friends_b200 import _lib
import ctypes as C
name = sys.argv[1] if len(sys.argv) > 1 else "upd_cat"
g = load_golden(name)
ro, pol = policy_from_update_golden(g)
ds = run_device_rollout(pol, ro)
for k in ("critic_observations", "observations", "raw_actions", "advantages", "log_probs", "rewards_to_go", "values"):
    print(k, np.abs(getattr(ds, k).cpu().numpy().astype(np.float64) - g["ds_" + k]).max())
B = int(g["hp_B"])
state = PPOUpdateState({"pol": pol}, batch_size=B, epochs_per_iter=1, normalize_adv=bool(g["hp_normalize_adv"]), normalize_values=bool(g["hp_normalize_values"]))
from ppo_and_friends_b200.ppo import _get_engine
eng = _get_engine(state, "pol", B)
upd = updater_from_golden(g)
ods = dataset_from_golden(g)
N = len(ds)
torch.manual_seed(int(g["hp_perm_seed"]))
perm = draw_minibatch_permutation(N)
# manual epoch on device, one minibatch at a time
n_mb = eng._ensure_epoch_buffers(N)
eng.refresh_hparams()
eng._perm_dev.copy_(perm)
lib = _lib.load()
from ppo_and_friends_b200._lib import ptr, stream_ptr, check
check(lib.ppoaf_epoch_prepare(ptr(eng._perm_dev), ptr(ds.advantages), ptr(ds.rewards_to_go), N, B, ptr(eng._mb_adv_stats), ptr(eng._mb_val_triples), stream_ptr()))
if eng.normalize_values:
    tr = eng._mb_val_triples.unsqueeze(0).contiguous()
    check(lib.ppoaf_value_stats_sequence(ptr(eng.value_normalizer.running_stats.state), ptr(tr), 1, n_mb, 1e-8, ptr(eng._mb_val_stats), stream_ptr()))
print("adv stats", eng._mb_adv_stats.cpu().numpy())
print("val stats", eng._mb_val_stats.cpu().numpy())
eng.epoch_stats.zero_(); eng.mb_cursor.zero_()
import copy
for k in range(n_mb):
    rows = min(B, N - k * B)
    idx = perm[k*B:k*B+rows]
    # oracle one minibatch
    st = upd.batch_train([ {kk: (vv[idx.numpy()] if kk != "values" else vv) for kk, vv in ods.items()} ] if False else [ods], [np.concatenate([perm.numpy()[k*B:k*B+rows]])] , B) if False else None
    bufs = eng._bufs(ds, rows)
    check(lib.ppoaf_ppo_minibatch_grads(C.byref(eng.cfg), C.byref(bufs), stream_ptr()))
    torch.cuda.synchronize()
    gd = {("actor/" + kk): vv.cpu().numpy().copy() for kk, vv in pol.actor.grad_dict().items()}
    gd.update({("critic/" + kk): vv.cpu().numpy().copy() for kk, vv in pol.critic.grad_dict().items()})
    # oracle grads for this minibatch with current oracle params
    i = idx
    rt = torch.as_tensor(ods["rewards_to_go"])[i]
    if upd.normalize_values:
        upd.value_stats.update(rt.numpy())
        mean = torch.tensor(upd.value_stats.mean, dtype=torch.float32); var = torch.tensor(upd.value_stats.variance, dtype=torch.float32)
        rt = (rt - mean) / torch.sqrt(var + torch.tensor([1e-8]))
    for p in upd.actor_parameters() + upd.critic.parameters(): p.grad = None
    al, cl, info = upd.minibatch_losses(torch.as_tensor(ods["critic_observations"])[i], torch.as_tensor(ods["observations"])[i], torch.as_tensor(ods["raw_actions"])[i], torch.as_tensor(ods["advantages"])[i], torch.as_tensor(ods["log_probs"])[i], rt)
    al.backward(); cl.backward()
    names_a = upd.actor.names + (["distribution.log_std"] if upd.log_std is not None else [])
    print(f"--- minibatch {k} rows {rows}: oracle actor {info['actor']:.6e} critic {info['critic']:.6e} kl {info['kl']:.3e}")
    es = eng.epoch_stats.cpu().numpy(); print("    device cumulative stats", es[:5])
    for nm, p in zip(names_a, upd.actor_parameters()):
        d = np.abs(gd["actor/" + nm] - p.grad.numpy()).max(); print(f"    grad actor/{nm}: maxabs {np.abs(p.grad.numpy()).max():.3e} err {d:.3e}")
    for nm, p in zip(upd.critic.names, upd.critic.parameters()):
        d = np.abs(gd["critic/" + nm] - p.grad.numpy()).max(); print(f"    grad critic/{nm}: maxabs {np.abs(p.grad.numpy()).max():.3e} err {d:.3e}")
    vals_dev = ds.values.cpu().numpy()[idx.numpy()]
    print("    values err", np.abs(vals_dev - info["values"].numpy()).max())
    upd._clip_and_step("actor", upd.actor_parameters()); upd._clip_and_step("critic", upd.critic.parameters())
    check(lib.ppoaf_ppo_minibatch_apply(C.byref(eng.cfg), C.byref(bufs), stream_ptr()))
    torch.cuda.synchronize()
    stt = upd.state()
    for net, obj in (("actor", pol.actor), ("critic", pol.critic)):
        for kk, vv in obj.state_dict().items():
            e = np.abs(vv.cpu().numpy() - stt[f"{net}/param/{kk}"]).max()
            print(f"    param {net}/{kk} err {e:.3e}")
