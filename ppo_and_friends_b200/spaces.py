"""
Minimal Box / Discrete stand-ins with the attributes the update path reads from gymnasium spaces
(`shape`, `dtype`, `low`, `high`, `n`).  gymnasium is not part of this image; real gymnasium
spaces work unchanged because the policy only duck-types these attributes.
"""
import numpy as np


class Box:
    def __init__(self, low, high, shape=None, dtype=np.float32):
        shape = np.shape(low) if shape is None else tuple(shape)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()


class Discrete:
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = start
        self.shape = ()
        self.dtype = np.dtype(np.int64)
