import os, torch, torch.distributed as dist
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
try:
    import torch.distributed._symmetric_memory as symm_mem
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=f"cuda:{local}")
    t.fill_(rank + 1.0)
    hdl = symm_mem.rendezvous(t, group=dist.group.WORLD)
    print(f"[{rank}] rendezvous ok: multicast_ptr={getattr(hdl, 'multicast_ptr', None)} n_bufs={len(hdl.buffer_ptrs)} signal_pads={len(hdl.signal_pad_ptrs)} world={hdl.world_size}", flush=True)
    hdl.barrier()
    if rank == 0:
        peer = hdl.get_buffer(1, (4,), torch.float32)
        print("[0] peer buffer of rank 1 reads", peer.tolist(), flush=True)
    try:
        out = torch.ops.symm_mem.multimem_all_reduce_(t, "sum", dist.group.WORLD.group_name)
        torch.cuda.synchronize()
        print(f"[{rank}] multimem_all_reduce_ ok: {t[:2].tolist()}", flush=True)
    except Exception as e:
        print(f"[{rank}] multimem_all_reduce_ failed: {type(e).__name__}: {str(e)[:300]}", flush=True)
except Exception as e:
    print(f"[{rank}] symm_mem failed: {type(e).__name__}: {str(e)[:500]}", flush=True)
dist.barrier(); dist.destroy_process_group()
