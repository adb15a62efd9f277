"""
RunningStatNormalizer and update_optimizer_lr (reference utils/misc.py:49-58, 61-172) on the
device-resident RunningMeanStd.
"""
import os
import pickle

import numpy as np
import torch

from .. import ops
from . import mpi_utils
from .stats import RunningMeanStd


def update_optimizer_lr(optim, lr):
    """Write `lr` into an optimizer-like object (utils/misc.py:49-58)."""
    for group in optim.param_groups:
        group["lr"] = lr


class RunningStatNormalizer(object):
    """(x - mean) / sqrt(var + eps) with optional statistics update on every call
    (utils/misc.py:84-111); `denormalize` is the inverse (:113-128)."""

    def __init__(self, name, device, test_mode=False, epsilon=1e-8):
        self.device = torch.device(device)
        self.name = name
        self.test_mode = test_mode
        self.running_stats = RunningMeanStd(device=self.device)
        self.epsilon = float(epsilon)

    def _as_device(self, data):
        x = data if torch.is_tensor(data) else torch.as_tensor(np.ascontiguousarray(data))
        return x.to(device=self.device, dtype=torch.float32).contiguous()

    def normalize(self, data, update_stats=True, gather_stats=True):
        x = self._as_device(data)
        if update_stats:
            self.running_stats.update(x.reshape(-1), gather_stats)
        return ops.normalize_clip(x.reshape(-1), self.running_stats.state, 1, self.epsilon).reshape(x.shape)

    def denormalize(self, data):
        x = self._as_device(data)
        return ops.denormalize(x.reshape(-1), self.running_stats.state, 1, self.epsilon).reshape(x.shape)

    def save_info(self, path):
        if self.test_mode:
            return
        f_name = "{}_stats_{}.pickle".format(self.name, mpi_utils.get_rank())
        with open(os.path.join(path, f_name), "wb") as fh:
            pickle.dump(self.running_stats, fh)

    def load_info(self, path):
        if self.test_mode:
            f_name = "{}_stats_0.pickle".format(self.name)
        else:
            f_name = "{}_stats_{}.pickle".format(self.name, mpi_utils.get_rank())
        in_file = os.path.join(path, f_name)
        if not os.path.exists(in_file):
            in_file = os.path.join(path, "{}_stats_0.pickle".format(self.name))
        with open(in_file, "rb") as fh:
            self.running_stats = pickle.load(fh)
