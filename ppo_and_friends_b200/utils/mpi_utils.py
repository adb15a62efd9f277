"""
Rank/communicator layer: the reference's utils/mpi_utils.py re-stated on torch.distributed
(NCCL over NVLink for device buffers; gloo in the CPU tests).  Same function names and meaning:

  rank_print                    utils/mpi_utils.py:11-35
  set_torch_threads             utils/mpi_utils.py:37-48
  broadcast_model_parameters    utils/mpi_utils.py:50-63   (one broadcast of the flat buffer, not one per tensor)
  mpi_avg                       utils/mpi_utils.py:65-86   (SUM all-reduce / num_procs)
  mpi_avg_gradients             utils/mpi_utils.py:89-111  (ONE all-reduce of the flat [actor|critic] gradient)

One process per GPU; when torch.distributed is not initialised everything is the size-1
communicator, exactly like the reference run without mpirun.  The reference aborts the MPI job on
errors (`rank_print(msg); comm.Abort()`); here `abort` raises after printing the same message.
"""
import sys

import numpy as np
import torch
import torch.distributed as dist


def is_distributed():
    return dist.is_available() and dist.is_initialized()


def get_rank():
    return dist.get_rank() if is_distributed() else 0


def get_num_procs():
    return dist.get_world_size() if is_distributed() else 1


def rank_print(msg, root=0, debug=False):
    if get_rank() == root:
        print("{}: {}".format(root, msg))
    sys.stdout.flush()


def abort(msg):
    """`rank_print(msg); comm.Abort()` of the reference, as an exception."""
    rank_print(msg, root=get_rank())
    raise RuntimeError(msg)


def set_torch_threads():
    if torch.get_num_threads() == 1:
        return
    torch.set_num_threads(max(int(torch.get_num_threads() / get_num_procs()), 1))


def barrier():
    if is_distributed():
        dist.barrier()


def broadcast_model_parameters(flat_params, root=0):
    """Rank `root`'s flat parameter buffer to everyone (a model object with `.flat_params` also works)."""
    if get_num_procs() == 1:
        return
    buf = getattr(flat_params, "flat_params", flat_params)
    dist.broadcast(buf, src=root)


def mpi_avg(data):
    """Average a python number / numpy array over ranks (SUM / num_procs)."""
    if not isinstance(data, (int, float, np.ndarray, np.floating, np.integer)):
        abort("ERROR: mpi_avg requires input to be of type float, int, or numpy ndarray.")
    n = get_num_procs()
    if n == 1:
        return data / n
    t = torch.as_tensor(np.asarray(data, dtype=np.float64))
    t = _to_comm_device(t)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    out = t.cpu().numpy() / n
    return out if isinstance(data, np.ndarray) else float(out)


def allreduce_sum_(tensor):
    """In-place SUM all-reduce of a tensor living where the backend wants it (CUDA for NCCL)."""
    if get_num_procs() > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def mpi_avg_gradients(flat_grads):
    """SUM the flat gradient buffer over ranks; the 1/R factor is applied inside the Adam kernel
    (PPOAF_HP_INV_WORLD), so sum-then-scale happens in the same order as the reference."""
    if get_num_procs() == 1:
        return
    buf = getattr(flat_grads, "flat_grads", flat_grads)
    dist.all_reduce(buf, op=dist.ReduceOp.SUM)


def all_gather_cat(tensor):
    """All-gather equal-shaped tensors and stack them in rank order: [R, ...]."""
    n = get_num_procs()
    if n == 1:
        return tensor.unsqueeze(0)
    out = torch.empty((n,) + tuple(tensor.shape), dtype=tensor.dtype, device=tensor.device)
    dist.all_gather_into_tensor(out, tensor.contiguous()) if tensor.is_cuda else \
        dist.all_gather(list(out.unbind(0)), tensor.contiguous())
    return out


def _to_comm_device(t):
    if is_distributed() and dist.get_backend() == "nccl":
        return t.cuda()
    return t
