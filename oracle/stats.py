"""
Oracle (test infrastructure, see oracle/__init__.py): running mean / variance and the
normalisers built on it.

Restates:
  RunningMeanStd.__init__/update/_integrate_batch_data   utils/stats.py:16-94
  RunningStatNormalizer.normalize/denormalize             utils/misc.py:84-128
  observation normalise + clip (microbench semantics)     environments/filter_wrappers.py:155-258, 617-660
"""
import numpy as np


class OracleRunningMeanStd:
    """mean/variance/count with the Chan parallel-variance merge (utils/stats.py:61-94)."""

    def __init__(self, shape=(), epsilon=1e-4):
        self.mean = np.zeros(shape, dtype=np.float32)
        self.variance = np.ones(shape, dtype=np.float32)
        self.count = epsilon

    def update(self, data, other_ranks=()):
        """`other_ranks`: the batches the other ranks hold; the reference allgathers the
        raw batches and concatenates in rank order before taking moments (stats.py:47-53)."""
        data = np.asarray(data)
        if len(other_ranks):
            data = np.concatenate([data] + [np.asarray(o) for o in other_ranks])
        self.integrate(np.mean(data, axis=0), np.var(data, axis=0), data.shape[0])

    def integrate(self, batch_mean, batch_var, batch_count):
        delta = batch_mean - self.mean
        tot = self.count + batch_count
        self.mean = self.mean + delta * (batch_count / tot)
        m2 = (self.variance * self.count + batch_var * batch_count
              + np.square(delta) * self.count * batch_count / (self.count + batch_count))
        self.variance = m2 / (self.count + batch_count)
        self.count += batch_count


def normalize(x, mean, variance, eps=1e-8):
    """(x - mu) / sqrt(var + eps) in fp32 (utils/misc.py:106-111)."""
    x = np.asarray(x, dtype=np.float32)
    mean = np.asarray(mean).astype(np.float32)
    variance = np.asarray(variance).astype(np.float32)
    return ((x - mean) / np.sqrt(variance + np.float32(eps))).astype(np.float32)


def denormalize(x, mean, variance, eps=1e-8):
    """mu + x * sqrt(var + eps) in fp32 (utils/misc.py:124-128)."""
    x = np.asarray(x, dtype=np.float32)
    mean = np.asarray(mean).astype(np.float32)
    variance = np.asarray(variance).astype(np.float32)
    return (mean + x * np.sqrt(variance + np.float32(eps))).astype(np.float32)


def normalize_clip_obs(obs, mean, variance, clip=10.0, eps=1e-8):
    """Observation normalise then clip to +-clip (filter_wrappers.py:220-221, 655-657)."""
    return np.clip(normalize(obs, mean, variance, eps), -clip, clip)
