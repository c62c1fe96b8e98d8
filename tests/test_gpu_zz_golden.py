"""The CUDA path against the COMMITTED golden vectors (tests/golden/golden_*.npz), without any oracle call: the fixtures were
dumped from the reference's own kernel text compiled for the host (tests/golden/make_golden.py); the CPU suite checks that the
oracle reproduces them bit for bit.  (Named to run last in the GPU suite.)"""
import numpy as np
import pytest

from pyapi import demb200, scenes
from test_gpu_parity import EARLY_BOUND_V, EARLY_BOUND_X, _mk

pytestmark = pytest.mark.gpu


def _golden_cases():
    from test_oracle_vs_ref import GOLDEN_CASES
    return GOLDEN_CASES


@pytest.mark.parametrize("kind", _golden_cases())
def test_early_state_matches_golden_fixture(built, kind):
    """The CUDA path against the COMMITTED golden vectors (tests/golden/golden_*.npz, dumped from the reference's own kernel
    text by tests/golden/make_golden.py; the oracle reproduces them bit for bit on the CPU side): state after 200 steps,
    before collisions have amplified round-off, within the fixed bounds 1e-7 m / 1e-3 m/s.  No oracle call on this path."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_%s.npz" % kind))
    f = scenes.flatten(_mk(kind))
    eng = demb200.Engine(0)
    eng.load_flat(f)
    eng.step(int(g["early_nsteps"]))
    nC = f.nClumps
    err_x = np.abs(eng.positions()[:nC] - g["early_pos"]).max()
    st = eng.owner_state()
    err_v = np.abs(st["vel"][:nC] - g["early_vel"]).max()
    # orientation: compare as rotations (q and -q are the same orientation)
    q, qg = st["oriQ"][:nC].astype("f8"), g["early_quat"].astype("f8")
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    qg /= np.linalg.norm(qg, axis=1, keepdims=True)
    err_q = (1.0 - np.abs((q * qg).sum(1))).max()
    print("%s: after %d steps |dx| %.2e m, |dv| %.2e m/s, 1 - |q.q_golden| %.2e" % (kind, int(g["early_nsteps"]), err_x, err_v, err_q))
    assert err_x <= EARLY_BOUND_X and err_v <= EARLY_BOUND_V and err_q <= 1e-6, (err_x, err_v, err_q)
    eng.close()


def test_sphere_level_reductions_match_numpy(built):
    """dem_reduce with the sphere-level kinds (what the facade's clump_max_z / clump_min_z / clump_max_absv inspectors call;
    AuxClasses.cpp:19-50 of the reference): top / bottom of every sphere and the speed of every sphere centre, against the
    same quantities computed with numpy from the downloaded owner state and the templates."""
    f = scenes.flatten(_mk("clumps_roll"))
    eng = demb200.Engine(0)
    eng.load_flat(f)
    eng.step(400)                                       # falling, colliding, spinning clumps
    st = eng.owner_state()
    pos = eng.positions()
    q = st["oriQ"].astype("f8")                        # w, x, y, z
    own = f.ownerClumpBody[: f.nSpheres].astype("i8")
    comp = f.clumpComponentOffset[: f.nSpheres].astype("i8")
    rel = np.stack([f.CDRelPosX, f.CDRelPosY, f.CDRelPosZ], 1).astype("f8")[comp]
    rad = f.Radii.astype("f8")[comp]

    def rotate(v, qq):                                  # applyOriQToVector3, DEMHelperKernels.cuh:161-173
        w, x, y, z = qq[:, 0], qq[:, 1], qq[:, 2], qq[:, 3]
        return np.stack([(2 * (w * w + x * x) - 1) * v[:, 0] + 2 * (x * y - w * z) * v[:, 1] + 2 * (x * z + w * y) * v[:, 2],
                         2 * (x * y + w * z) * v[:, 0] + (2 * (w * w + y * y) - 1) * v[:, 1] + 2 * (y * z - w * x) * v[:, 2],
                         2 * (x * z - w * y) * v[:, 0] + 2 * (y * z + w * x) * v[:, 1] + (2 * (w * w + z * z) - 1) * v[:, 2]], 1)

    r_world = rotate(rel, q[own])
    zc = pos[own, 2] + r_world[:, 2]
    v_centre = st["vel"][own].astype("f8") + rotate(np.cross(st["omg"][own].astype("f8"), rel), q[own])
    scale = np.abs(pos).max()
    assert abs(eng.reduce(5) - (zc + rad).max()) < 1e-5 * scale        # DEM_REDUCE_SPHERE_MAX_Z (float arithmetic on the device)
    assert abs(eng.reduce(6) - (zc - rad).min()) < 1e-5 * scale        # DEM_REDUCE_SPHERE_MIN_Z
    vmax = np.linalg.norm(v_centre, axis=1).max()
    assert abs(eng.reduce(7) - vmax) < 1e-5 * max(vmax, 1.0)           # DEM_REDUCE_SPHERE_MAX_ABSV
    # and they differ from the owner-level forms where they should: a clump's spheres reach beyond its centre
    assert eng.reduce(5) > eng.reduce(demb200.REDUCE_MAX_Z) and eng.reduce(6) < eng.reduce(demb200.REDUCE_MIN_Z)
    eng.close()
