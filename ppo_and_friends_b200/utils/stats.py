"""
RunningMeanStd with device-resident state (reference utils/stats.py:9-94).

Same constructor, same `update(data, gather_stats=True)`, same public attributes
`mean` / `variance` / `count` (host numpy views, so checkpoints can pickle them), but the
moments are computed by the CUDA Welford/Chan kernels and the cross-rank step exchanges the
(mean, M2, n) triples instead of all-gathering the raw batch (utils/stats.py:47-50).
"""
import numpy as np
import torch

from .. import ops
from . import mpi_utils


class RunningMeanStd(object):

    def __init__(self, shape=(), epsilon=1e-4, device=None):
        self.shape = tuple(shape) if not isinstance(shape, int) else (shape,)
        self.dim = int(np.prod(self.shape)) if len(self.shape) else 1
        self.device = torch.device("cuda") if device is None else torch.device(device)
        st = np.concatenate([np.zeros(self.dim), np.ones(self.dim), [epsilon]])
        # state = mean[dim] | var[dim] | count, fp64 on device
        self.state = torch.tensor(st, dtype=torch.float64, device=self.device)

    # -- reference-visible attributes ----------------------------------------------------------
    @property
    def mean(self):
        return self.state[:self.dim].cpu().numpy().reshape(self.shape)

    @property
    def variance(self):
        return self.state[self.dim:2 * self.dim].cpu().numpy().reshape(self.shape)

    @property
    def count(self):
        return float(self.state[2 * self.dim].item())

    def update(self, data, gather_stats=True):
        """data: [n, *shape] numpy array or CUDA tensor (any float dtype; computed in fp32/fp64)."""
        x = data if torch.is_tensor(data) else torch.as_tensor(np.ascontiguousarray(data))
        x = x.to(device=self.device, dtype=torch.float32).contiguous()
        if x.numel() == 0:
            return
        triple = ops.batch_moments(x, self.dim)
        if gather_stats and mpi_utils.get_num_procs() > 1:
            triples = mpi_utils.all_gather_cat(triple).contiguous()
        else:
            triples = triple
        ops.stats_merge(self.state, triples, self.dim)

    def load_reference(self, mean, variance, count):
        st = np.concatenate([np.asarray(mean, dtype=np.float64).reshape(-1),
                             np.asarray(variance, dtype=np.float64).reshape(-1), [float(count)]])
        self.state.copy_(torch.tensor(st, dtype=torch.float64))

    def __getstate__(self):
        return dict(shape=self.shape, dim=self.dim, mean=self.mean, variance=self.variance, count=self.count,
                    device=str(self.device))

    def __setstate__(self, d):
        self.shape, self.dim = d["shape"], d["dim"]
        self.device = torch.device(d["device"])
        self.state = torch.empty(2 * self.dim + 1, dtype=torch.float64, device=self.device)
        self.load_reference(d["mean"], d["variance"], d["count"])
