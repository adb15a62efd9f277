"""
Actor / critic feed-forward networks as views into ONE flat device buffer per policy.

Mirrors the structure the reference builds with torch.nn (networks/ppo_networks/feed_forward.py:14-86,
networks/utils.py:53-80, 114-191, networks/actor_critic/wrappers.py:10-82): Linear(in,h) -> act ->
[Linear(h,h) -> act] x (depth-1) -> Linear(h,out); orthogonal init with gain sqrt(2), output gain
0.01 (actor) / 1.0 (critic), zero biases; the Gaussian actor owns `distribution.log_std`.
state_dict() keys are the reference's (`sequential_net.0.*`, `sequential_net.2.{0,2,..}.*`,
`sequential_net.3.*`, `distribution.log_std`), so reference checkpoints load and vice versa.

The flat layout [actor | critic] (params, grads, Adam m, Adam v share it) is what the CUDA step and
the NCCL all-reduce operate on: one gradient all-reduce per minibatch instead of one per tensor.
"""
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from .. import _lib, ops

_ACT_NAMES = {nn.ReLU: "relu", nn.LeakyReLU: "leaky_relu", nn.Tanh: "tanh", nn.Identity: "identity"}


def activation_name(activation):
    """Map a torch activation instance (the reference passes module instances) or a string."""
    if isinstance(activation, str):
        if activation not in _lib.ACT_IDS:
            raise ValueError(f"unsupported activation {activation}")
        return activation
    for cls, name in _ACT_NAMES.items():
        if isinstance(activation, cls):
            if cls is nn.LeakyReLU and abs(activation.negative_slope - 0.01) > 1e-12:
                raise ValueError("LeakyReLU slope other than 0.01 is not supported by the CUDA path")
            return name
    raise ValueError(f"unsupported activation {activation!r} (supported: ReLU, LeakyReLU(0.01), Tanh)")


def hidden_sizes(hidden_size, hidden_depth):
    """networks/utils.py:140-158."""
    if not isinstance(hidden_size, list):
        if (hidden_size == 0) != (hidden_depth == 0):
            raise ValueError("if either hidden_size or hidden_depth is 0, both must be 0")
        return [hidden_size] * hidden_depth
    return list(hidden_size)


def layer_key_stems(n_hidden):
    if n_hidden == 0:
        return ["sequential_net.0"]
    return (["sequential_net.0"] + [f"sequential_net.2.{2 * i}" for i in range(n_hidden - 1)]
            + ["sequential_net.3"])


def reference_init(dims, out_gain):
    """Host-side initialisation in the reference's construction order (so a given torch seed
    yields the same weights): nn.Linear's own reset, then orthogonal_/constant_ (networks/utils.py:53-80)."""
    layers = []
    n = len(dims) - 1
    for l in range(n):
        lin = nn.Linear(dims[l], dims[l + 1])
        gain = np.sqrt(2) if (l + 1 < n or out_gain is None) else out_gain
        nn.init.orthogonal_(lin.weight, gain)
        nn.init.constant_(lin.bias, 0.0)
        layers.append((lin.weight.detach().clone(), lin.bias.detach().clone()))
    return layers


class FlatNetwork:
    """One MLP living at [base, base + total) of the policy's flat buffers."""

    def __init__(self, name, dims, activation, log_std_dim=0):
        self.name = name
        self.dims = [int(d) for d in dims]
        self.activation = activation_name(activation)
        self.desc = _lib.MlpDesc.make(self.dims, self.activation)
        self.log_std_dim = int(log_std_dim)
        self.offsets, self.total = _lib.param_layout(self.desc, self.log_std_dim)
        self.stems = layer_key_stems(len(self.dims) - 2)
        self.base = 0
        self.owner = None

    def bind(self, owner, base):
        self.owner, self.base = owner, base

    def _views(self, flat):
        out = OrderedDict()
        n = len(self.dims) - 1
        for l in range(n):
            w0 = self.base + self.offsets[2 * l]
            b0 = self.base + self.offsets[2 * l + 1]
            out[self.stems[l] + ".weight"] = flat[w0:w0 + self.dims[l] * self.dims[l + 1]].view(self.dims[l + 1], self.dims[l])
            out[self.stems[l] + ".bias"] = flat[b0:b0 + self.dims[l + 1]]
        if self.log_std_dim:
            s0 = self.base + self.offsets[2 * n]
            out["distribution.log_std"] = flat[s0:s0 + self.log_std_dim]
        return out

    def state_dict(self):
        return self._views(self.owner.flat_params)

    def grad_dict(self):
        return self._views(self.owner.flat_grads)

    def adam_dicts(self):
        return self._views(self.owner.adam_m), self._views(self.owner.adam_v)

    def parameters(self):
        return list(self.state_dict().values())

    def load_state_dict(self, sd):
        mine = self.state_dict()
        missing = [k for k in mine if k not in sd]
        if missing:
            raise KeyError(f"{self.name}: missing keys {missing}")
        with torch.no_grad():
            for k, v in mine.items():
                v.copy_(torch.as_tensor(np.asarray(sd[k].detach().cpu() if torch.is_tensor(sd[k]) else sd[k]),
                                        dtype=torch.float32).reshape(v.shape))

    @property
    def flat(self):
        return self.owner.flat_params[self.base:self.base + self.total]

    def __call__(self, obs, softmax_out=False):
        """Forward through the CUDA MLP kernels; obs: [n, in] fp32 (numpy or tensor)."""
        x = obs if torch.is_tensor(obs) else torch.as_tensor(np.ascontiguousarray(obs))
        x = x.to(device=self.owner.device, dtype=torch.float32).reshape(x.shape[0], -1).contiguous()
        return ops.mlp_forward(self.desc, self.flat, x, softmax_out=softmax_out)

    def save(self, path, rank=0):
        import os
        torch.save(OrderedDict((k, v.detach().cpu().clone()) for k, v in self.state_dict().items()),
                   os.path.join(path, "{}_{}.model".format(self.name, rank)))

    def load(self, path, rank=0):
        import os
        f = os.path.join(path, "{}_{}.model".format(self.name, rank))
        if not os.path.exists(f):
            f = os.path.join(path, "{}_0.model".format(self.name))
        self.load_state_dict(torch.load(f))


class PolicyNetworks:
    """Flat [actor | critic] parameter / gradient / Adam buffers of one policy."""

    def __init__(self, device, actor_dims, critic_dims, actor_activation, critic_activation, gaussian, act_dim,
                 std_offset=0.5, init=True):
        self.device = torch.device(device)
        self.actor = FlatNetwork("actor", actor_dims, actor_activation, act_dim if gaussian else 0)
        self.critic = FlatNetwork("critic", critic_dims, critic_activation, 0)
        self.n_actor, self.n_critic = self.actor.total, self.critic.total
        n = self.n_actor + self.n_critic
        self.flat_params = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.flat_grads = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.adam_m = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.adam_v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.adam_step = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.actor.bind(self, 0)
        self.critic.bind(self, self.n_actor)
        if init:
            # actor first, then critic: the reference's construction order (ppo_policy.py:433-446)
            a_layers = reference_init(actor_dims, 0.01)
            c_layers = reference_init(critic_dims, 1.0)
            sd = OrderedDict()
            for stem, (w, b) in zip(self.actor.stems, a_layers):
                sd[stem + ".weight"], sd[stem + ".bias"] = w, b
            if gaussian:
                sd["distribution.log_std"] = torch.full((act_dim,), -float(std_offset))
            self.actor.load_state_dict(sd)
            sd = OrderedDict()
            for stem, (w, b) in zip(self.critic.stems, c_layers):
                sd[stem + ".weight"], sd[stem + ".bias"] = w, b
            self.critic.load_state_dict(sd)


class _AdamView:
    """Just enough of torch.optim.Adam's surface for the trainer (`param_groups[i]['lr']`)."""

    def __init__(self, lr):
        self.param_groups = [dict(lr=lr, betas=(0.9, 0.999), eps=1e-5)]
