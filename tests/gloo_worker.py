"""Worker of tests/test_mgpu_host_gloo.py: one process per (pretend) GPU, world_size 2, gloo on CPU.  Exercises the
host side of the multi-GPU path: rendezvous, the unique-id broadcast bench.py does, the slab decomposition and the
ownership / halo rule (dem_host_partition_owners), and the max-over-ranks reduction of timings.  No device compute."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dem-engine_b200"))

from pyapi import demb200, dist_util, scenes  # noqa: E402


def main():
    out_path = sys.argv[1]
    rank, local_rank, world = dist_util.env_rank()
    dist = dist_util.init("gloo")
    assert dist.get_world_size() == world == 2

    # 1. rank 0's 128 opaque bytes reach every rank unchanged (the NCCL unique id travels this way in bench.py)
    token = dist_util.share_bytes(lambda: (np.arange(128) * 7 + 3).astype("u1"), 128)
    assert np.array_equal(token, (np.arange(128) * 7 + 3).astype("u1"))

    # 2. every rank builds the same scene; slabs tile the x axis
    sc = scenes.config2_clumps(24, 6, 4, spacing=2.7)
    f = scenes.flatten(sc)
    p = demb200.params_from_flat(f)
    lo, hi = demb200.host_slab_bounds(p, world, rank)
    halo = 0.012
    role, send = demb200.host_partition_owners(p, world, rank, halo, f.voxelID[: f.nClumps], f.locX[: f.nClumps])
    n_own, n_ghost = int((role == 1).sum()), int((role == 2).sum())
    n_send_l, n_send_r = int((send & 1).astype(bool).sum()), int((send & 2).astype(bool).sum())
    counts = dist_util.gather_counts([n_own, n_ghost, n_send_l, n_send_r])

    # 3. ownership is a partition: exchange the role vectors and check owner by owner
    import torch
    mine = torch.from_numpy(role.astype("i8"))
    both = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    roles = np.stack([b.numpy() for b in both])
    owned_by = (roles == 1).sum(0)
    assert (owned_by == 1).all(), "every owner is owned by exactly one rank"
    # what my neighbour sends me is exactly what I hold as ghosts
    snd = torch.from_numpy(send.astype("i8"))
    sends = [torch.zeros_like(snd) for _ in range(world)]
    dist.all_gather(sends, snd)
    sends = np.stack([s.numpy() for s in sends])
    other = 1 - rank
    bit = 2 if other < rank else 1  # the neighbour on my left sends right (bit 1 -> value 2), and vice versa
    assert np.array_equal((sends[other] & bit) != 0, roles[rank] == 2)

    # 4. timings are reported as the slowest rank's
    slow = dist_util.max_over_ranks([10.0 + rank, 5.0 - rank])
    assert slow == [11.0, 5.0]

    dist.barrier()
    with open(out_path + ".%d" % rank, "w") as fh:
        json.dump({"rank": rank, "lo": lo, "hi": hi, "counts": counts.tolist(), "n": int(f.nClumps)}, fh)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
