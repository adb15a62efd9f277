"""
Oracle (test infrastructure, see oracle/__init__.py): the PPO minibatch update in
torch-CPU fp32 — networks, action heads, losses, gradient clipping, Adam, value
normaliser, KL statistic and the minibatch permutation protocol.

Restates:
  FeedForwardNetwork.forward / create_sequential_network  networks/ppo_networks/feed_forward.py:66-86, networks/utils.py:114-191
  GaussianDistribution (softplus std, tanh-corrected log-prob, entropy = -log_prob(mean))
                                                          networks/distributions.py:491-558, 694
  CategoricalDistribution (softmax in the actor, Categorical(probs)) networks/distributions.py:221, 249, 1045
  PPOPolicy.evaluate / update_weights                     policies/ppo_policy.py:891-952, 1012-1055
  PPO._ppo_batch_train                                    ppo.py:2274-2485
  mpi_avg_gradients (sum over ranks / R, per tensor)      utils/mpi_utils.py:89-111
  RandomSampler / DataLoader draw protocol                torch/utils/data/sampler.py:160-185, dataloader.py:705-710

Third-party arithmetic (torch.optim.Adam `_single_tensor_adam`, clip_grad_norm_,
Normal.log_prob, Categorical's probs->logits) is restated from the torch 2.11
sources installed in this image — the only implementation available to the reference
here — and pinned against the reference's outputs in tests/golden/upd_*.npz.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .stats import OracleRunningMeanStd

ACTIVATIONS = {
    "leaky_relu": lambda x: F.leaky_relu(x, 0.01),
    "tanh": torch.tanh,
    "relu": torch.relu,
}

LOG_SQRT_2PI = math.log(math.sqrt(2.0 * math.pi))


def layer_keys(depth):
    """State-dict key stems of the reference Sequential (networks/utils.py:160-183)."""
    keys = ["sequential_net.0"]
    keys += [f"sequential_net.2.{2 * i}" for i in range(depth - 1)]
    keys += ["sequential_net.3"]
    return keys


class OracleMLP:
    def __init__(self, params, activation):
        """params: dict state-dict-key -> np.ndarray (reference naming); weights are [out, in]."""
        self.names = list(params.keys())
        self.p = {k: torch.tensor(np.asarray(v), dtype=torch.float32, requires_grad=True) for k, v in params.items()}
        stems = [k[:-len(".weight")] for k in self.names if k.endswith(".weight")]
        self.stems = stems
        self.act = ACTIVATIONS[activation]

    def forward(self, x):
        h = x.flatten(start_dim=1)
        for i, s in enumerate(self.stems):
            h = h @ self.p[s + ".weight"].t() + self.p[s + ".bias"]
            if i + 1 < len(self.stems):
                h = self.act(h)
        return h

    def parameters(self):
        return [self.p[k] for k in self.names]

    def numpy_params(self):
        return {k: v.detach().numpy().copy() for k, v in self.p.items()}


# ---- action heads ---------------------------------------------------------------------
def gaussian_std(log_std, min_std=0.01):
    return torch.clamp(F.softplus(log_std), min=min_std)                 # distributions.py:514-515


def gaussian_tanh_log_prob(mean, std, x, eps=1e-6):
    var = std * std
    lp = -((x - mean) ** 2) / (2 * var) - torch.log(std) - LOG_SQRT_2PI   # torch Normal.log_prob
    lp = torch.clamp(lp, -100, 100).sum(dim=-1)                           # distributions.py:551-553
    tp = torch.clamp(1.0 - torch.tanh(x) ** 2, min=eps)                   # :555-556
    return lp - torch.log(tp).sum(dim=-1)


def categorical_from_probs(probs):
    """torch Categorical(probs): renormalise, clamp to [eps, 1-eps], log (distributions/utils.py probs_to_logits)."""
    p = probs / probs.sum(dim=-1, keepdim=True)
    eps = torch.finfo(p.dtype).eps
    logits = torch.log(torch.clamp(p, eps, 1 - eps))
    return p, logits


class OracleUpdater:
    """One policy's actor + critic + Adam + value normaliser, trained the reference way."""

    def __init__(self, actor_params, critic_params, activation, discrete, lr=3e-4,
                 entropy_weight=0.01, surr_clip=0.2, vf_clip=None, gradient_clip=0.5,
                 kl_loss_weight=0.0, use_huber_loss=False, normalize_adv=True,
                 normalize_values=True, betas=(0.9, 0.999), adam_eps=1e-5):
        self.discrete = discrete
        log_std = actor_params.get("distribution.log_std")
        self.actor = OracleMLP({k: v for k, v in actor_params.items() if k != "distribution.log_std"}, activation)
        self.critic = OracleMLP(critic_params, activation)
        self.log_std = None
        if not discrete:
            self.log_std = torch.tensor(np.asarray(log_std), dtype=torch.float32, requires_grad=True)
        self.lr, self.entropy_weight, self.surr_clip = lr, entropy_weight, surr_clip
        self.vf_clip, self.gradient_clip, self.kl_loss_weight = vf_clip, gradient_clip, kl_loss_weight
        self.use_huber_loss = use_huber_loss
        self.normalize_adv, self.normalize_values = normalize_adv, normalize_values
        self.betas, self.adam_eps = betas, adam_eps
        self.value_stats = OracleRunningMeanStd()
        self.adam = {"actor": self._adam_state(self.actor_parameters()),
                     "critic": self._adam_state(self.critic.parameters())}

    def actor_parameters(self):
        ps = self.actor.parameters()
        return ps + ([self.log_std] if self.log_std is not None else [])

    @staticmethod
    def _adam_state(params):
        return dict(step=0, m=[torch.zeros_like(p) for p in params], v=[torch.zeros_like(p) for p in params])

    # -- PPOPolicy.evaluate (ppo_policy.py:891-952) -----------------------------------------
    def evaluate(self, critic_obs, obs, raw_actions):
        values = self.critic.forward(critic_obs).squeeze()
        pred = self.actor.forward(obs)
        if self.discrete:
            probs = torch.softmax(pred, dim=-1)
            p, logits = categorical_from_probs(probs)
            lp = logits.gather(-1, raw_actions.reshape(-1, 1).long()).reshape(-1)
            ent = -(p * logits).sum(dim=-1)
        else:
            std = gaussian_std(self.log_std)
            if raw_actions.dim() < 2:
                raw_actions = raw_actions.unsqueeze(1)
            lp = gaussian_tanh_log_prob(pred, std, raw_actions)
            ent = -gaussian_tanh_log_prob(pred, std, pred)               # distributions.py:694
        return values, lp, ent

    # -- clip_grad_norm_ + Adam (torch nn/utils/clip_grad.py, optim/adam.py) --------------------
    def _clip_and_step(self, which, params, grads_override=None):
        grads = [p.grad for p in params] if grads_override is None else grads_override
        if self.gradient_clip is not None:
            norms = torch.stack([torch.linalg.vector_norm(g, 2.0) for g in grads])
            total = torch.linalg.vector_norm(norms, 2.0)
            coef = torch.clamp(self.gradient_clip / (total + 1e-6), max=1.0)
            grads = [g * coef for g in grads]
        st = self.adam[which]
        st["step"] += 1
        t = st["step"]
        b1, b2 = self.betas
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        step_size = self.lr / bc1
        bc2_sqrt = bc2 ** 0.5
        with torch.no_grad():
            for p, g, m, v in zip(params, grads, st["m"], st["v"]):
                m.lerp_(g, 1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / bc2_sqrt).add_(self.adam_eps)
                p.addcdiv_(m, denom, value=-step_size)

    # -- one minibatch of PPO._ppo_batch_train for a list of simulated ranks -----------------
    def minibatch_losses(self, critic_obs, obs, raw_actions, adv, old_lp, rtg_norm):
        if self.normalize_adv:
            adv = (adv - adv.mean()) / (adv.std() + 1e-8)                # ppo.py:2325-2333
        values, lp, ent = self.evaluate(critic_obs, obs, raw_actions)
        values_det = values.detach().reshape(-1).clone()
        lp, old_lp, adv, ent = lp.flatten(), old_lp.flatten(), adv.flatten(), ent.flatten()
        values, rtg_norm = values.flatten(), rtg_norm.flatten()
        ratios = torch.exp(lp - old_lp)
        surr1 = ratios * adv
        surr2 = torch.clamp(ratios, 1 - self.surr_clip, 1 + self.surr_clip) * adv
        kl = (old_lp - lp).mean().item()                                  # ppo.py:2358
        actor_loss = (-torch.min(surr1, surr2)).mean()
        reported_actor = actor_loss.item()                                # before entropy/KL terms (:2392-2393)
        ent_mean = None
        if self.entropy_weight != 0.0:
            ent_mean = ent.mean().item()
            actor_loss = actor_loss - self.entropy_weight * ent.mean()
        if self.kl_loss_weight > 0.0:
            actor_loss = actor_loss + self.kl_loss_weight * kl            # constant, no gradient (Q7)
        loss_fn = (lambda a, b: F.huber_loss(a, b, delta=10.0)) if self.use_huber_loss else F.mse_loss
        critic_loss = loss_fn(values, rtg_norm)
        if self.vf_clip is not None:                                      # intended semantics (Q4)
            clipped = torch.clamp(values, -self.vf_clip, self.vf_clip)
            critic_loss = torch.max(critic_loss, loss_fn(clipped, rtg_norm))
        return actor_loss, critic_loss, dict(actor=reported_actor, critic=critic_loss.item(),
                                             kl=kl, entropy=ent_mean, values=values_det)

    def batch_train(self, datasets, perms, batch_size):
        """
        datasets: list (one per simulated rank) of dicts with flat tensors/arrays
            critic_observations, observations, raw_actions, advantages, log_probs,
            rewards_to_go, values (values is updated in place, ppo.py:2340).
        perms: list of int64 permutations, one per rank.
        Returns the status scalars the reference writes (ppo.py:2471-2485).
        """
        R = len(datasets)
        N = len(perms[0])
        tot = dict(actor=0.0, critic=0.0, entropy=0.0, kl=0.0, counter=0)
        for start in range(0, N, batch_size):
            idxs = [torch.as_tensor(p[start:start + batch_size], dtype=torch.long) for p in perms]
            rtgs = [torch.as_tensor(d["rewards_to_go"])[i] for d, i in zip(datasets, idxs)]
            if self.normalize_values:
                # every rank allgathers the raw minibatches, then integrates them (stats.py:47-59)
                self.value_stats.update(rtgs[0].numpy(), [r.numpy() for r in rtgs[1:]])
                mean = torch.tensor(self.value_stats.mean, dtype=torch.float32)
                var = torch.tensor(self.value_stats.variance, dtype=torch.float32)
                rtgs = [(r - mean) / torch.sqrt(var + torch.tensor([1e-8])) for r in rtgs]
            if idxs[0].numel() == 1:                                      # ppo.py:2305
                continue
            a_grads, c_grads, infos = [], [], []
            for d, i, rtg in zip(datasets, idxs, rtgs):
                for p in self.actor_parameters() + self.critic.parameters():
                    p.grad = None
                al, cl, info = self.minibatch_losses(
                    torch.as_tensor(d["critic_observations"])[i], torch.as_tensor(d["observations"])[i],
                    torch.as_tensor(d["raw_actions"])[i], torch.as_tensor(d["advantages"])[i],
                    torch.as_tensor(d["log_probs"])[i], rtg)
                al.backward()
                cl.backward()
                a_grads.append([p.grad.clone() for p in self.actor_parameters()])
                c_grads.append([p.grad.clone() for p in self.critic.parameters()])
                infos.append(info)
                vals = d["values"]
                if isinstance(vals, np.ndarray):
                    vals[i.numpy()] = info["values"].numpy()
                else:
                    vals[i] = info["values"]
            avg = lambda gs: [sum(g[k] for g in gs) / R for k in range(len(gs[0]))] if R > 1 else gs[0]
            self._clip_and_step("actor", self.actor_parameters(), avg(a_grads))
            self._clip_and_step("critic", self.critic.parameters(), avg(c_grads))
            for info in infos:
                tot["actor"] += info["actor"]
                tot["critic"] += info["critic"]
                tot["kl"] += info["kl"]
                if info["entropy"] is not None:
                    tot["entropy"] += info["entropy"]
                tot["counter"] += 1
        c = max(tot["counter"], 1)
        return {"actor loss": tot["actor"] / c, "critic loss": tot["critic"] / c, "kl avg": tot["kl"] / c,
                "weighted entropy": tot["entropy"] * self.entropy_weight / c, "counter": tot["counter"]}

    def state(self):
        out = {}
        for k, v in self.actor.numpy_params().items():
            out["actor/param/" + k] = v
        if self.log_std is not None:
            out["actor/param/distribution.log_std"] = self.log_std.detach().numpy().copy()
        for k, v in self.critic.numpy_params().items():
            out["critic/param/" + k] = v
        return out


def draw_minibatch_permutation(n):
    """
    The draws `for batch in DataLoader(ds, shuffle=True)` makes on torch's global CPU
    generator (dataloader.py:705-710 then sampler.py:160-185): #1 the (unused) base seed,
    #2 the seed of the private generator that feeds randperm.
    """
    _base_seed = torch.empty((), dtype=torch.int64).random_().item()
    seed = int(torch.empty((), dtype=torch.int64).random_().item())
    g = torch.Generator()
    g.manual_seed(seed)
    return torch.randperm(n, generator=g)
