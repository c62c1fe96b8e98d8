#!/usr/bin/env python3
"""Multi-GPU profiling helper (torchrun; not part of the product): stage times of a contact-list rebuild and of the
per-step kernels on every rank of the slab decomposition of the C2 bed."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import bench  # noqa: E402
from pyapi import demb200, dist_util, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clumps", type=int, default=1000000)
ap.add_argument("--settle-steps", type=int, default=40000)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist_util.init("nccl", local)
sc, dims = bench.build_scene(args.clumps, 20, 2.7)
f = scenes.flatten(sc)
eng = demb200.Engine(local)
eng.load_flat(f, contact_capacity=0 if world == 1 else int(f.nSpheres) * 6 // world + 200000)
if world > 1:
    eng.mgpu_init(rank, world, dist_util.share_bytes(demb200.Engine.mgpu_unique_id, 128, device="cuda"))
eng.step(args.settle_steps)
for rep in range(3):
    if world > 1:
        dist.barrier()
    r = eng.profile_rebuild()
    print("rank %d rebuild %s" % (rank, json.dumps({k: round(v, 1) for k, v in r.items()})), flush=True)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
eng.step(2000)
eng.sync()
if world > 1:
    dist.barrier()
dt = time.perf_counter() - t0
print("rank %d: 2000 steps %.1f steps/s; info %s" % (rank, 2000 / dt, eng.mgpu_info() if world > 1 else {}), flush=True)
if world > 1:
    dist.destroy_process_group()
