"""torch.distributed plumbing of the multi-GPU path (one process per GPU).  Backend-agnostic on purpose: bench.py uses
it over NCCL, the CPU tests drive the same functions over gloo with world_size 2."""
import os

import numpy as np


def env_rank():
    """(rank, local_rank, world) as torchrun exports them."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend, local_rank=0):
    """Join the process group described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT."""
    import torch
    import torch.distributed as dist
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend)
    return dist


def share_bytes(make_on_root, nbytes, device="cpu"):
    """Rank 0 produces `nbytes` bytes (e.g. the NCCL unique id of dem_mgpu_unique_id); every rank returns them."""
    import torch
    import torch.distributed as dist
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        src = np.ascontiguousarray(make_on_root(), "u1")
        assert src.size == nbytes
        buf.copy_(torch.from_numpy(src).to(device))
    dist.broadcast(buf, 0)
    return buf.cpu().numpy()


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def gather_counts(values, device="cpu"):
    """All ranks' integer tuples, as an array [world, len(values)]."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([o.cpu().numpy() for o in out])
