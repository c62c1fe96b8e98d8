#!/usr/bin/env python3
"""Multi-GPU parity check, one process per GPU (run under torchrun on a box with >= 2 GPUs; launched by
tests/test_gpu_mgpu.py::test_torchrun_ranks_match_single_gpu when the box has the GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/mgpu_check.py

Every rank builds the same scene and joins the slab decomposition; rank 0 additionally runs the same scene on one GPU,
plain and with velocities perturbed by 1e-6 (the round-off sensitivity yardstick of tests/test_gpu_parity.py).  At each
checkpoint the state merged from the ranks (each owner taken from the rank whose slab holds it) must agree with the
single-GPU run within 10x that sensitivity, owners must have crossed the cuts, and the halo must be non-empty.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from pyapi import demb200, scenes  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dims = [int(v) for v in os.environ.get("MGPU_CHECK_DIMS", "28,8,6").split(",")]
    sc = scenes.config2_clumps(dims[0], dims[1], dims[2], cd_update_freq=10, spacing=2.7)
    n = len(sc.clump_type)
    rng = np.random.RandomState(7)
    # a shearing, colliding bed: owners cross the cuts in both directions
    sc.clump_vel = np.stack([np.where(sc.clump_xyz[:, 2] > sc.clump_xyz[:, 2].mean(), 1.5, -1.5) + rng.normal(size=n) * 0.1,
                             rng.normal(size=n) * 0.1, np.full(n, -1.0)], 1).astype("f4")
    f = scenes.flatten(sc)
    eng = demb200.Engine(local)
    eng.load_flat(f)
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.from_numpy(demb200.Engine.mgpu_unique_id()).cuda())
    dist.broadcast(uid, 0)
    eng.mgpu_init(rank, world, uid.cpu().numpy())
    if rank == 0:
        e1 = demb200.Engine(local)
        e1.load_flat(f)
        fp = scenes.flatten(sc)
        for name in ("vX", "vY", "vZ"):
            a = getattr(fp, name)
            a[:] = (a.astype("f8") * (1.0 + 1e-6)).astype("f4")
        e1p = demb200.Engine(local)
        e1p.load_flat(fp)
    import ctypes as C
    p = eng.params
    bounds = []
    for r in range(world):
        lo, hi = C.c_float(), C.c_float()
        demb200.load_library().dem_host_slab_bounds(C.byref(p), world, r, C.byref(lo), C.byref(hi))
        bounds.append((lo.value, hi.value))
    x_init = eng.positions()[:n, 0].copy()
    ok = True
    done = 0
    for cp in (200, 600, 1200, 2000):
        eng.step(cp - done)
        if rank == 0:
            e1.step(cp - done)
            e1p.step(cp - done)
        done = cp
        pos = torch.from_numpy(eng.positions()[:n]).cuda()
        vel = torch.from_numpy(eng.owner_state()["vel"][:n].astype("f8")).cuda()
        allp = [torch.zeros_like(pos) for _ in range(world)]
        allv = [torch.zeros_like(vel) for _ in range(world)]
        dist.all_gather(allp, pos)
        dist.all_gather(allv, vel)
        info = eng.mgpu_info()
        infos = [None] * world
        dist.all_gather_object(infos, info)
        if rank == 0:
            ref_p, ref_v = e1.positions()[:n], e1.owner_state()["vel"][:n].astype("f8")
            xrel = ref_p[:, 0] - float(f.LBF[0])
            merged_p, merged_v = np.zeros_like(ref_p), np.zeros_like(ref_v)
            owner_rank = np.zeros(n, "i4")
            for r, (lo, hi) in enumerate(bounds):
                m = (xrel >= lo) & (xrel < hi)
                merged_p[m] = allp[r].cpu().numpy()[m]
                merged_v[m] = allv[r].cpu().numpy()[m]
                owner_rank[m] = r
            sens_x = np.abs(e1p.positions()[:n] - ref_p).max()
            sens_v = np.abs(e1p.owner_state()["vel"][:n].astype("f8") - ref_v).max()
            err_x, err_v = np.abs(merged_p - ref_p).max(), np.abs(merged_v - ref_v).max()
            x0rel = x_init - float(f.LBF[0])
            init_rank = np.zeros(n, "i4")
            for r, (lo, hi) in enumerate(bounds):
                init_rank[(x0rel >= lo) & (x0rel < hi)] = r
            crossed = int((init_rank != owner_rank).sum())
            good = err_x <= 10 * sens_x + 1e-7 and err_v <= 10 * sens_v + 1e-5
            ok = ok and good
            print("step %5d: |dx| %.2e (sens %.2e) |dv| %.2e (sens %.2e) crossed cuts %d  %s  %s" % (
                cp, err_x, sens_x, err_v, sens_v, crossed, "OK" if good else "FAIL",
                " ".join("r%d own %d act %d halo %dB" % (r, i["n_own"], i["n_active"], i["halo_bytes_per_step"]) for r, i in enumerate(infos))), flush=True)
            if cp == 2000:
                ok = ok and crossed > 0 and all(i["halo_bytes_per_step"] > 0 for i in infos)
                ok = ok and sum(i["n_own"] for i in infos) == n  # every clump has exactly one owner rank
                for r, i in enumerate(infos):  # interior ranks push both ways
                    ok = ok and (i["n_send_left"] > 0) == (r > 0) and (i["n_send_right"] > 0) == (r < world - 1)
    res = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(res, 0)
    if rank == 0:
        print("exchange: NVLink peer stores + flags, device-driven (no NCCL, no host synchronisation on the path)", flush=True)
        print("MGPU_CHECK", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(res.item()) == 1 else 1)


if __name__ == "__main__":
    main()
