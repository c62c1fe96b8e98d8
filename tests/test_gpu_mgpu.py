"""Multi-GPU parity (config 3 of BASELINE.json): the slab decomposition over 2 / 4 / 8 GPUs against the SAME scene on one
GPU, through the C ABI.  Skipped when the box has fewer GPUs than the case needs (the driver's box has one; run with
`gpurun --gpus N -- python -m pytest tests/test_gpu_mgpu.py -m gpu`).

Two forms of the decomposition are exercised:
  * in-process (dem_mgpu_init_local + dem_group_step_async / dem_group_sync): what deme::DEMSolver(nGPUs) uses;
  * one process per GPU under torchrun (dem_mgpu_init; tests/mgpu_check.py), launched from here as a subprocess.
The bed shears (upper half moves +x, lower half -x) so owners cross the cuts in both directions, and with >= 3 ranks the
interior ranks push halo records both ways.  Tolerance: the merged state may differ from the single-GPU state by at most
10x what the single-GPU run differs from itself when its initial velocities are perturbed by 1e-6 relative (the yardstick
of tests/test_gpu_parity.py), plus 1e-7 m / 1e-5 m/s.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import _cuda_device_count
from pyapi import demb200, scenes

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def shear_bed(nx, ny, nz, seed=7, cd=10):
    sc = scenes.config2_clumps(nx, ny, nz, cd_update_freq=cd, spacing=2.7)
    n = len(sc.clump_type)
    rng = np.random.RandomState(seed)
    up = sc.clump_xyz[:, 2] > sc.clump_xyz[:, 2].mean()
    sc.clump_vel = np.stack([np.where(up, 1.5, -1.5) + rng.normal(size=n) * 0.1, rng.normal(size=n) * 0.1,
                             np.full(n, -1.0)], 1).astype("f4")
    return sc


def _need(world):
    have = _cuda_device_count()
    if have < world:
        pytest.skip("needs %d GPUs, this box has %d" % (world, have))


@pytest.mark.parametrize("world,dims", [(2, (28, 8, 6)), (2, (64, 28, 28)), (4, (64, 28, 28)), (8, (64, 28, 28))])
def test_local_group_matches_single_gpu(built, world, dims):
    _need(world)
    sc = shear_bed(*dims)
    f = scenes.flatten(sc)
    n = f.nClumps
    fp = scenes.flatten(sc)
    for name in ("vX", "vY", "vZ"):
        a = getattr(fp, name)
        a[:] = (a.astype("f8") * (1.0 + 1e-6)).astype("f4")
    e1, e1p = demb200.Engine(0), demb200.Engine(0)
    e1.load_flat(f)
    e1p.load_flat(fp)
    grp = demb200.EngineGroup(f, list(range(world)))
    x_init = e1.positions()[:n, 0].copy()
    done = 0
    for cp in (200, 600, 1200, 2000):
        grp.step(cp - done)
        e1.step(cp - done)
        e1p.step(cp - done)
        done = cp
        e0 = grp.gather()
        ref_p, ref_v = e1.positions()[:n], e1.owner_state()["vel"][:n].astype("f8")
        got_p, got_v = e0.positions()[:n], e0.owner_state()["vel"][:n].astype("f8")
        sens_x = np.abs(e1p.positions()[:n] - ref_p).max()
        sens_v = np.abs(e1p.owner_state()["vel"][:n].astype("f8") - ref_v).max()
        err_x, err_v = np.abs(got_p - ref_p).max(), np.abs(got_v - ref_v).max()
        infos = [e.mgpu_info() for e in grp.engines]
        print("N=%d step %5d: |dx| %.2e (sens %.2e) |dv| %.2e (sens %.2e)  %s" % (
            world, cp, err_x, sens_x, err_v, sens_v,
            " ".join("r%d own %d act %d halo %dB" % (r, i["n_own"], i["n_active"], i["halo_bytes_per_step"]) for r, i in enumerate(infos))))
        assert err_x <= 10 * sens_x + 1e-7, (cp, err_x, sens_x)
        assert err_v <= 10 * sens_v + 1e-5, (cp, err_v, sens_v)
        # every clump is owned by exactly one rank, every rank has a halo, interior ranks push both ways
        assert sum(i["n_own"] for i in infos) == n
        assert all(i["halo_bytes_per_step"] > 0 for i in infos)
        for r, i in enumerate(infos):
            assert (i["n_send_left"] > 0) == (r > 0) and (i["n_send_right"] > 0) == (r < world - 1), (r, i)
    # owners crossed the cuts
    import ctypes as C
    xrel0 = x_init - float(f.LBF[0])
    xrel1 = e1.positions()[:n, 0] - float(f.LBF[0])
    crossed = 0
    for r in range(world):
        lo, hi = C.c_float(), C.c_float()
        demb200.load_library().dem_host_slab_bounds(C.byref(e1.params), world, r, C.byref(lo), C.byref(hi))
        crossed += int((((xrel0 >= lo.value) & (xrel0 < hi.value)) & ~((xrel1 >= lo.value) & (xrel1 < hi.value))).sum())
    print("owners that changed rank:", crossed)
    assert crossed > 0
    # the stepping path ran as replayed cycle graphs with no host synchronisation inside
    st = grp.engines[0].stats()
    assert st.n_rebuilds >= 2000 // f.cd_update_freq
    grp.close()
    e1.close()
    e1p.close()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_torchrun_ranks_match_single_gpu(built, world):
    """one process per GPU (dem_mgpu_init, cudaIpc-mapped peer blocks): tests/mgpu_check.py under torchrun"""
    _need(world)
    env = dict(os.environ)
    env.setdefault("MGPU_CHECK_DIMS", "28,8,6" if world == 2 else "64,28,28")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(29520 + world), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0 and "MGPU_CHECK PASS" in r.stdout


@pytest.mark.parametrize("world", [2, 4])
def test_group_context_is_a_drop_in_for_one_gpu(built, world):
    """dem_ctx_create_group: ONE context driving several GPUs (what deme::DEMSolver(nGPUs) creates).  Every C-ABI call a
    single-GPU caller makes must give the same answers: stepping, state read-back (merged from the ranks), reductions,
    the contact list (union of the ranks' lists, pairs across a cut listed once), host-side state changes in between
    (broadcast to the ranks, ownership re-derived), and the same again after more steps."""
    _need(world)
    sc = shear_bed(40, 12, 10)
    f = scenes.flatten(sc)
    n = f.nClumps
    fp = scenes.flatten(sc)
    for name in ("vX", "vY", "vZ"):
        a = getattr(fp, name)
        a[:] = (a.astype("f8") * (1.0 + 1e-6)).astype("f4")
    e1, e1p = demb200.Engine(0), demb200.Engine(0)
    e1.load_flat(f)
    e1p.load_flat(fp)
    eg = demb200.Engine(devices=list(range(world)))
    eg.set_option("group_min_owners", 100)   # (the default only shards scenes with >= 50 000 clumps per GPU)
    eg.load_flat(f)
    assert eg.mgpu_info()["world"] == world

    def compare(label):
        ref_p, ref_v = e1.positions()[:n], e1.owner_state()["vel"][:n].astype("f8")
        sens_x = np.abs(e1p.positions()[:n] - ref_p).max()
        sens_v = np.abs(e1p.owner_state()["vel"][:n].astype("f8") - ref_v).max()
        err_x = np.abs(eg.positions()[:n] - ref_p).max()
        err_v = np.abs(eg.owner_state()["vel"][:n].astype("f8") - ref_v).max()
        print("%s: |dx| %.2e (sens %.2e) |dv| %.2e (sens %.2e)" % (label, err_x, sens_x, err_v, sens_v))
        assert err_x <= 10 * sens_x + 1e-7 and err_v <= 10 * sens_v + 1e-5, (label, err_x, sens_x, err_v, sens_v)
        # inspectors: same reductions over ALL clumps
        a = eg.reduce_many((demb200.REDUCE_MAX_ABSV, demb200.REDUCE_KINETIC_ENERGY, demb200.REDUCE_TOTAL_MASS, demb200.REDUCE_MAX_Z))
        b = e1.reduce_many((demb200.REDUCE_MAX_ABSV, demb200.REDUCE_KINETIC_ENERGY, demb200.REDUCE_TOTAL_MASS, demb200.REDUCE_MAX_Z))
        assert abs(a[demb200.REDUCE_TOTAL_MASS] - b[demb200.REDUCE_TOTAL_MASS]) <= 1e-9 * b[demb200.REDUCE_TOTAL_MASS]
        assert abs(a[demb200.REDUCE_KINETIC_ENERGY] - b[demb200.REDUCE_KINETIC_ENERGY]) <= 1e-3 * b[demb200.REDUCE_KINETIC_ENERGY] + 20 * sens_v
        assert abs(a[demb200.REDUCE_MAX_Z] - b[demb200.REDUCE_MAX_Z]) <= 10 * sens_x + 1e-6
        # contact list: the touching pairs of the merged list are those of the single-GPU list
        ga, gb, gt, gw = eg.contacts()
        sa, sb, st_, sw = e1.contacts()
        key = lambda a_, b_, t_: (t_.astype("u8") << np.uint64(60)) | (a_.astype("u8") << np.uint64(30)) | b_.astype("u8")
        kg, ks = key(ga, gb, gt), key(sa, sb, st_)
        assert len(np.unique(kg)) == len(kg), "a pair across a cut is listed twice"
        tg, ts = set(kg[np.abs(gw).max(1) > 0].tolist()), set(ks[np.abs(sw).max(1) > 0].tolist())
        diff = len(tg ^ ts)
        print("%s: %d / %d touching pairs, symmetric difference %d" % (label, len(tg), len(ts), diff))
        assert diff <= max(2, len(ts) // 200)

    eg.step(200); e1.step(200); e1p.step(200)
    compare("N=%d after 200 steps" % world)
    # a host-side change of state in the middle (what trackers' SetVel / SetPos do): kick a block of clumps
    kick = np.tile(np.array([[0.0, 0.3, 0.5]], "f4"), (50, 1))
    for e in (eg, e1, e1p):
        e.upload_owner_state(100, vel=kick)
    eg.step(400); e1.step(400); e1p.step(400)
    compare("N=%d after the kick + 400 steps" % world)
    for e in (eg, e1, e1p):
        e.close()


def test_group_context_with_a_prescribed_mesh(built):
    """A triangle mesh under the slab decomposition (every rank registers the facets that can meet its slab): clumps thrown
    into a tilted, spinning box of facets, 2 GPUs against 1.  The mesh owner is replicated on both ranks, which is exact
    because its motion is fully prescribed; a free-moving mesh makes the group fall back to one GPU instead."""
    _need(2)
    sc = scenes.config2_clumps(20, 6, 4, cd_update_freq=5, spacing=2.7, init_vel=(0.3, 0.1, -2.0))
    sc.bounding = "none"
    lo, hi = sc.clump_xyz.min(0), sc.clump_xyz.max(0)
    v, fc = scenes.box_mesh((hi[0] - lo[0]) * 1.3, (hi[1] - lo[1]) * 1.6, (hi[2] - lo[2]) * 2.2, n=6, inward=True)
    sc.add_mesh(v, fc, mat=0, mass=1.0, moi=(1, 1, 1), pos=tuple((lo + hi) / 2), quat=(0.9950042, 0.0998334, 0, 0), family=10)
    sc.prescribed[10] = dict(linvel=(0.0, 0.0, 0.0), angvel=(0.0, 0.0, 2.0))
    f = scenes.flatten(sc)
    n = f.nClumps
    fp = scenes.flatten(sc)
    for name in ("vX", "vY", "vZ"):
        a = getattr(fp, name)
        a[:n] = (a[:n].astype("f8") * (1.0 + 1e-6)).astype("f4")
    e1, e1p = demb200.Engine(0), demb200.Engine(0)
    e1.load_flat(f)
    e1p.load_flat(fp)
    eg = demb200.Engine(devices=[0, 1])
    eg.set_option("group_min_owners", 10)
    eg.load_flat(f)
    assert eg.mgpu_info()["world"] == 2
    for cp in (300, 900):
        for e in (eg, e1, e1p):
            e.step(300 if cp == 300 else 600)
        ref = e1.positions()[:n]
        sens = np.abs(e1p.positions()[:n] - ref).max()
        err = np.abs(eg.positions()[:n] - ref).max()
        print("mesh, 2 GPUs, step %d: |dx| %.2e (sens %.2e), facet contacts %d vs %d" % (
            cp, err, sens, eg.stats().n_contacts_st, e1.stats().n_contacts_st))
        assert err <= 10 * sens + 1e-7, (cp, err, sens)
    assert e1.stats().n_contacts_st > 0
    # nothing left the box of facets
    pos = eg.positions()[:n]
    assert np.isfinite(pos).all()
    # a mesh that moves freely cannot be replicated: the same scene without the prescription runs on one GPU
    sc.prescribed.pop(10)
    f2 = scenes.flatten(sc)
    eg2 = demb200.Engine(devices=[0, 1])
    eg2.set_option("group_min_owners", 10)
    eg2.load_flat(f2)
    assert eg2.mgpu_info()["world"] == 1
    eg2.step(50)
    for e in (eg, eg2, e1, e1p):
        e.close()
