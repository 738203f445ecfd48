"""Row sharding across the GPUs of one box: one process per GPU (torchrun), A split into
contiguous row blocks, x and every n-vector replicated, one sum-allreduce of the n-vector
A_i'(A_i x) per normal-operator apply.  torch.distributed is used only to
ship the NCCL unique id and for barriers / max-over-ranks timing."""
from __future__ import annotations

import os


def row_range(m, rank, nranks, align=4):
    """Contiguous row block [lo, hi) of rank `rank`; blocks are multiples of `align` rows
    (128-bit loads) except possibly the last."""
    per = -(-m // nranks)
    per = -(-per // align) * align
    lo = min(m, rank * per)
    hi = min(m, lo + per)
    return lo, hi


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_comm(ctx, rank=None, nranks=None, peer_floats=1 << 18):
    """Join this context to the job's NCCL communicator.  Requires torch.distributed to be
    initialised (any backend) for the id exchange."""
    import torch
    import torch.distributed as dist
    if rank is None:
        rank, nranks = dist.get_rank(), dist.get_world_size()
    if nranks == 1:
        return ctx
    payload = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(payload, src=0)
    ctx.comm_init(rank, nranks, payload[0])
    if peer_floats and os.environ.get("RLS_P2P", "0") == "1":
        # exchange buffers of the one-shot NVLink all-reduce (csrc/rls_p2p.cu): CUDA IPC handles through the host
        # transport; a box without peer access keeps the NCCL path
        try:
            mine = ctx.peer_export(peer_floats)
            handles = [None] * nranks
            dist.all_gather_object(handles, mine)
            ctx.peer_import(handles)
        except Exception as e:  # noqa: BLE001
            ok = 0
            import warnings
            warnings.warn(f"peer-memory all-reduce unavailable, using NCCL: {e}")
        else:
            ok = 1
        t = torch.tensor([ok], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if int(t[0]) == 0:
            os.environ["RLS_P2P"] = "0"   # all ranks must take the same path
    return ctx


def shard_rows(A_global, rank, nranks, align=4):
    """Host-side helper: the row block of a NumPy matrix owned by `rank`."""
    lo, hi = row_range(A_global.shape[0], rank, nranks, align)
    return A_global[lo:hi], (lo, hi)
