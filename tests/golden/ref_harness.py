"""
Import harness for the UNMODIFIED reference (LLNL/ppo_and_friends) at /root/reference.

Only `tests/golden/make_golden.py` uses this, and only in the build container
(/root/reference does not exist on the GPU box).  It installs stand-in modules for
the packages the reference imports at module scope but that are absent from this
image (mpi4py, gym, gymnasium, plotly, moviepy) and exposes the reference tree as
the package `ppo_and_friends` (reference setup.py:6-12 maps the repo root to that
package name) through a symlink in a temp dir.  No reference source is edited.
"""
import os
import sys
import tempfile
import types

import numpy as np

_REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _find_reference():
    """/root/reference in the build container; on the GPU box the git-ignored copy made by oracle/build_ref.py."""
    for cand in (os.environ.get("PPOAF_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO, "oracle", "_ref", "ppo_and_friends")):
        if cand and os.path.isdir(cand) and os.path.exists(os.path.join(cand, "ppo.py")):
            return cand
    return os.environ.get("PPOAF_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_reference()


def available():
    return os.path.isdir(REFERENCE_ROOT) and os.path.exists(os.path.join(REFERENCE_ROOT, "ppo.py"))


def _install_mpi_stub():
    mpi4py = types.ModuleType("mpi4py")
    MPI = types.ModuleType("mpi4py.MPI")

    class _Comm:
        """COMM_WORLD stand-in.  Single process: identity.  When torch.distributed is initialised (gloo, the R-rank CPU
        arm of bench.py) the calls the reference makes on this path (utils/mpi_utils.py:50-111, utils/stats.py:47-50,
        ppo.py:2471-2475) are carried by the matching gloo collectives."""

        @staticmethod
        def _dist():
            import torch.distributed as dist
            return dist if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 else None

        def Get_rank(self):
            d = self._dist()
            return d.get_rank() if d else 0

        def Get_size(self):
            d = self._dist()
            return d.get_world_size() if d else 1

        def allreduce(self, x, op=None):
            d = self._dist()
            if d is None:
                return x
            import torch
            red = {"SUM": d.ReduceOp.SUM, "MAX": d.ReduceOp.MAX, "MIN": d.ReduceOp.MIN}[op or "SUM"]
            if isinstance(x, np.ndarray):
                t = torch.from_numpy(np.ascontiguousarray(x).copy())
                d.all_reduce(t, op=red)
                return t.numpy()
            t = torch.tensor([x], dtype=torch.float64)
            d.all_reduce(t, op=red)
            return type(x)(t.item()) if isinstance(x, (int, float)) else t.item()

        def allgather(self, x):
            d = self._dist()
            if d is None:
                return [x]
            out = [None] * d.get_world_size()
            d.all_gather_object(out, x)
            return out

        def Bcast(self, buf, root=0):
            d = self._dist()
            if d is None:
                return None
            import torch
            t = torch.from_numpy(buf)
            d.broadcast(t, src=root)
            return None

        def bcast(self, x, root=0):
            d = self._dist()
            if d is None:
                return x
            box = [x]
            d.broadcast_object_list(box, src=root)
            return box[0]

        def barrier(self):
            d = self._dist()
            if d is not None:
                d.barrier()

        def Barrier(self):
            self.barrier()

        def Abort(self, code=1):
            raise RuntimeError("comm.Abort() called by the reference")

    MPI.COMM_WORLD = _Comm()
    MPI.SUM, MPI.MAX, MPI.MIN = "SUM", "MAX", "MIN"
    mpi4py.MPI = MPI
    sys.modules["mpi4py"] = mpi4py
    sys.modules["mpi4py.MPI"] = MPI


class _Space:
    dtype = None
    shape = None

    def seed(self, seed=None):
        return [seed]

    def sample(self):
        raise NotImplementedError


class Box(_Space):
    def __init__(self, low, high, shape=None, dtype=np.float32):
        if shape is None:
            shape = np.shape(low)
        self.shape = tuple(shape)
        self.dtype = np.dtype(dtype)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()


class Discrete(_Space):
    def __init__(self, n, start=0):
        self.n = int(n)
        self.start = start
        self.shape = ()
        self.dtype = np.dtype(np.int64)


class MultiDiscrete(_Space):
    def __init__(self, nvec, dtype=np.int64, start=None):
        self.nvec = np.asarray(nvec, dtype=dtype)
        self.start = np.zeros_like(self.nvec) if start is None else np.asarray(start)
        self.shape = self.nvec.shape
        self.dtype = np.dtype(dtype)


class MultiBinary(_Space):
    def __init__(self, n):
        self.n = n
        self.shape = (n,)
        self.dtype = np.dtype(np.int8)


class Tuple(_Space):
    def __init__(self, spaces):
        self.spaces = tuple(spaces)

    def __iter__(self):
        return iter(self.spaces)

    def __len__(self):
        return len(self.spaces)

    def __getitem__(self, i):
        return self.spaces[i]


class Dict(dict, _Space):
    shape = None


def _install_gym_stubs():
    for top in ("gymnasium", "gym"):
        mod = types.ModuleType(top)
        spaces = types.ModuleType(top + ".spaces")
        utils = types.ModuleType(top + ".spaces.utils")
        for cls in (Box, Discrete, MultiDiscrete, MultiBinary, Tuple, Dict):
            setattr(spaces, cls.__name__, cls)
        spaces.Space = _Space
        spaces.utils = utils
        utils.flatten_space = lambda s: s
        mod.spaces = spaces
        mod.Env = object
        mod.Wrapper = object
        sys.modules[top] = mod
        sys.modules[top + ".spaces"] = spaces
        sys.modules[top + ".spaces.utils"] = utils


def _install_misc_stubs():
    for name in ("plotly", "plotly.graph_objects", "plotly.express", "moviepy", "moviepy.editor"):
        sys.modules.setdefault(name, types.ModuleType(name))


_linked = None


def install():
    """Make `import ppo_and_friends` resolve to the unmodified reference tree."""
    global _linked
    if _linked is not None:
        return _linked
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_mpi_stub()
    _install_gym_stubs()
    _install_misc_stubs()
    d = tempfile.mkdtemp(prefix="ppoaf_ref_")
    os.symlink(REFERENCE_ROOT, os.path.join(d, "ppo_and_friends"))
    sys.path.insert(0, d)
    sys.dont_write_bytecode = True  # /root/reference is read-only
    _linked = d
    return d
