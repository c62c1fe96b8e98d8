"""Pins the CPU oracle (oracle/dem_oracle.c) against the reference's own kernel text compiled for the host
(oracle/_ref/libdemref.so) -- bit for bit -- and against the committed golden vectors generated from it.

There are no golden vectors or known-answer tests in the reference tree; what pins the oracle is
(i) the reference's kernels executed here through the host shim and (ii) tests/golden/*.npz dumped from that.
"""
import math
import os

import numpy as np
import pytest

from oracle import pyoracle
from pyapi import scenes

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
needs_ref = pytest.mark.skipif(pyoracle.ref() is None, reason="oracle/_ref/libdemref.so not built (no reference tree)")

STATE = ["voxelID", "locX", "locY", "locZ", "oriQw", "oriQx", "oriQy", "oriQz", "vX", "vY", "vZ", "omgBarX", "omgBarY",
         "omgBarZ", "aX", "aY", "aZ", "alphaX", "alphaY", "alphaZ"]


def _scene(kind):
    if kind == "clumps_full":
        sc = scenes.config2_clumps(5, 5, 4, cd_update_freq=5, spacing=2.7, init_vel=(0.2, 0.1, -2.0))
    elif kind == "clumps_roll":
        sc = scenes.config2_clumps(4, 4, 3, cd_update_freq=5, spacing=2.7, Crr=0.05, init_vel=(0.3, 0.0, -2.0))
    elif kind == "spheres_frictionless":
        sc = scenes.config1_spheres(n_side=7, cd_update_freq=5, jitter=0.02)
        sc.clump_vel[:] = (0.0, 0.0, -2.0)
    elif kind == "cylinder":
        sc = scenes.config2_clumps(4, 4, 3, cd_update_freq=5, spacing=2.7, init_vel=(1.5, 0.6, -2.0))
        sc.bounding = "only_bottom"
        sc.add_cylinder((0, 0, 0), (0, 0, 1), 0.04, 0, normal=0.0)
    elif kind == "mesh_tray":
        # clumps thrown into a tilted, spinning box made of triangles (sphere--triangle contacts)
        sc = scenes.config2_clumps(4, 4, 3, cd_update_freq=5, spacing=2.7, init_vel=(0.3, 0.1, -2.0))
        sc.bounding = "none"
        lo, hi = sc.clump_xyz.min(0), sc.clump_xyz.max(0)
        v, fc = scenes.box_mesh((hi[0] - lo[0]) * 1.6, (hi[1] - lo[1]) * 1.6, (hi[2] - lo[2]) * 2.2, n=3, inward=True)
        sc.add_mesh(v, fc, mat=0, mass=1.0, moi=(1, 1, 1), pos=tuple((lo + hi) / 2), quat=(0.9950042, 0.0998334, 0, 0), family=10)
        sc.prescribed[10] = dict(linvel=(0.0, 0.0, 0.0), angvel=(0.0, 0.0, 2.0))
    elif kind == "drum":
        # BASELINE configs[3] at oracle size: polydisperse clump templates in a rotating drum of triangles
        sc = scenes.config4_drum(1000, 1500, omega=6.0, init_vel=(0.2, 0.0, -1.0), cd_update_freq=10, spacing=2.7)
    elif kind == "families":
        # fixed and prescribed families, a disabled family pair, an extra contact margin
        sc = scenes.config2_clumps(5, 5, 4, cd_update_freq=5, spacing=2.7, init_vel=(0.2, 0.1, -2.0))
        fam = np.zeros(len(sc.clump_type), "u1")
        fam[::7] = 3
        fam[1::7] = 5
        sc.clump_family = fam
        sc.fixed_families.append(3)
        sc.prescribed[5] = dict(linvel=(0.1, None, -0.5), dictate=True)
        sc.disabled_pairs.append((3, 5))
    elif kind in ("forward_euler", "centered_difference"):
        sc = scenes.config2_clumps(4, 4, 3, cd_update_freq=5, spacing=2.7, init_vel=(0.3, 0.1, -2.0))
        sc.integrator = 0 if kind == "forward_euler" else 1  # DEM_FORWARD_EULER / DEM_CENTERED_DIFFERENCE
    else:
        raise KeyError(kind)
    return sc


def _assert_same_state(a, b):
    for name in STATE:
        x, y = getattr(a, name), getattr(b, name)
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), name
    assert a.nContacts == b.nContacts
    n = a.nContacts
    for k in range(4):
        assert np.array_equal(a.contactWildcards[k][:n].view("u4"), b.contactWildcards[k][:n].view("u4")), "wildcard %d" % k


@needs_ref
@pytest.mark.parametrize("kind", ["clumps_full", "clumps_roll", "spheres_frictionless", "cylinder", "mesh_tray", "drum",
                                  "families", "forward_euler", "centered_difference"])
def test_step_bit_exact_vs_reference_kernels(built, kind):
    f = scenes.flatten(_scene(kind))
    a = pyoracle.world_from_flat(f)
    b = a.copy()
    nsteps = 4000
    a.step(nsteps, cd_every=f.cd_update_freq)
    b.step(nsteps, cd_every=f.cd_update_freq, use_ref=True)
    assert a.nContacts > 0
    # something physical must have happened: contacts with non-zero history or forces
    assert np.abs(a.contactForces[: 3 * a.nContacts]).max() > 0
    _assert_same_state(a, b)
    n = a.nContacts
    for name in ("contactForces", "contactTorque_convToForce", "contactPointGeometryA", "contactPointGeometryB"):
        assert np.array_equal(getattr(a, name)[: 3 * n].view("u4"), getattr(b, name)[: 3 * n].view("u4")), name


@needs_ref
def test_codec_bit_exact(built):
    import ctypes as C
    f = scenes.flatten(_scene("clumps_full"))
    w = pyoracle.world_from_flat(f)
    rng = np.random.RandomState(1)
    s = w.struct()
    for _ in range(2000):
        xyz = (rng.uniform(0.0, 0.1, 3)).astype("f8")
        v1, v2 = C.c_uint64(), C.c_uint64()
        l1, l2 = (C.c_uint16 * 3)(), (C.c_uint16 * 3)()
        pyoracle.lib().orc_voxel_encode(C.byref(s), xyz.ctypes.data_as(C.c_void_p), C.byref(v1), l1)
        pyoracle.ref().ref_voxel_encode(C.byref(s), xyz.ctypes.data_as(C.c_void_p), C.byref(v2), l2)
        assert v1.value == v2.value and list(l1) == list(l2)
    d1, d2 = np.zeros(3), np.zeros(3)
    for o in range(0, w.nOwners, 7):
        pyoracle.lib().orc_voxel_decode(C.byref(s), C.c_uint32(o), d1.ctypes.data_as(C.c_void_p))
        pyoracle.ref().ref_voxel_decode(C.byref(s), C.c_uint32(o), d2.ctypes.data_as(C.c_void_p))
        assert np.array_equal(d1, d2)


@needs_ref
def test_margin_bit_exact(built):
    f = scenes.flatten(_scene("clumps_full"))
    a = pyoracle.world_from_flat(f)
    rng = np.random.RandomState(2)
    a.vX[:] = rng.normal(size=len(a.vX)).astype("f4")
    a.vZ[:] = rng.normal(size=len(a.vZ)).astype("f4") * 3
    b = a.copy()
    a.compute_margins(20)
    b.compute_margins(20, use_ref=True)
    assert np.array_equal(a.marginSize.view("u4"), b.marginSize.view("u4"))
    a.beta = b.beta = np.float32(1e-4)
    a.compute_margins(20)
    b.compute_margins(20, use_ref=True)
    assert np.array_equal(a.marginSize.view("u4"), b.marginSize.view("u4"))


@needs_ref
def test_sphere_analytical_contacts_match_reference_bin_kernels(built):
    """The oracle's sphere--analytical detection equals what getNumberOfBinsEachSphereTouches /
    populateBinSphereTouchingPairs (DEMBinSphereKernels.cu) emit."""
    import ctypes as C
    f = scenes.flatten(_scene("cylinder"))
    w = pyoracle.world_from_flat(f)
    w.step(3000, cd_every=5)
    w.compute_margins(5)
    w.detect_contacts()
    idA, idB, ct, _ = w.contacts()
    sa = ct > 10
    mine = sorted(zip(idA[sa].tolist(), idB[sa].tolist(), ct[sa].tolist()))
    cap = 10 * w.nSpheres + 16
    oS, oO, oT = np.zeros(cap, "u4"), np.zeros(cap, "u4"), np.zeros(cap, "u1")
    s = w.struct()
    n = pyoracle.ref().ref_sphere_anal_contacts(C.byref(s), C.c_double(0.02), C.c_uint32(64), C.c_uint32(64),
                                                C.c_uint32(64), oS.ctypes.data_as(C.c_void_p),
                                                oO.ctypes.data_as(C.c_void_p), oT.ctypes.data_as(C.c_void_p),
                                                C.c_long(cap), None)
    assert n >= 0
    theirs = sorted(zip(oS[:n].tolist(), oO[:n].tolist(), oT[:n].tolist()))
    assert len(mine) > 0 and mine == theirs


@needs_ref
def test_pair_acceptance_matches_calcContactPoint(built):
    """Sphere--sphere acceptance of the oracle's broad phase == calcContactPoint (DEMContactKernels_SphereSphere.cu:57)."""
    import ctypes as C
    f = scenes.flatten(_scene("clumps_full"))
    w = pyoracle.world_from_flat(f)
    w.step(4000, cd_every=5)
    w.compute_margins(5)
    w.detect_contacts()
    idA, idB, ct, _ = w.contacts()
    ss = set(zip(idA[ct == 1].tolist(), idB[ct == 1].tolist()))
    pos, rad = w.sphere_positions()
    rinf = (rad + w.marginSize[w.ownerClumpBody[: w.nSpheres]]).astype("f4")
    n = w.nSpheres
    hits = set()
    binid = C.c_uint32()
    for a in range(n):
        d = np.linalg.norm(pos[a + 1:] - pos[a], axis=1)
        for b in (np.nonzero(d < 2.2 * rinf.max())[0] + a + 1):
            if w.ownerClumpBody[a] == w.ownerClumpBody[b]:
                continue
            A, B = np.ascontiguousarray(pos[a]), np.ascontiguousarray(pos[b])
            if pyoracle.ref().ref_calc_contact_point(C.c_double(0.05), C.c_uint32(100), C.c_uint32(100),
                                                     A.ctypes.data_as(C.c_void_p), C.c_float(rinf[a]),
                                                     B.ctypes.data_as(C.c_void_p), C.c_float(rinf[b]), C.c_float(0),
                                                     C.c_float(0), C.byref(binid)):
                hits.add((a, int(b)))
    assert len(ss) > 0 and ss == hits


GOLDEN_CASES = ["clumps_full", "spheres_frictionless", "mesh_tray", "clumps_roll", "cylinder", "drum", "families",
                "forward_euler"]


@pytest.mark.parametrize("kind", GOLDEN_CASES)
def test_oracle_reproduces_golden_vectors(built, kind):
    """Golden trajectories dumped from the host-compiled reference kernels (tests/golden/make_golden.py)."""
    path = os.path.join(GOLDEN, "golden_%s.npz" % kind)
    g = np.load(path)
    f = scenes.flatten(_scene(kind))
    w = pyoracle.world_from_flat(f)
    w.step(int(g["nsteps"]), cd_every=f.cd_update_freq)
    for name in STATE:
        assert np.array_equal(getattr(w, name)[: w.nOwners].view(np.uint8), g[name].view(np.uint8)), name
    assert w.nContacts == int(g["nContacts"])
    assert np.array_equal(w.idGeometryA[: w.nContacts], g["idGeometryA"])
    assert np.array_equal(w.idGeometryB[: w.nContacts], g["idGeometryB"])
    # the early snapshot (what tests/test_gpu_parity.py holds the CUDA path against)
    we = pyoracle.world_from_flat(f)
    we.step(int(g["early_nsteps"]), cd_every=f.cd_update_freq)
    assert np.array_equal(we.positions_f64()[: f.nClumps], g["early_pos"])
    assert np.array_equal(np.stack([we.vX, we.vY, we.vZ], 1)[: f.nClumps].view("u4"), g["early_vel"].view("u4"))


def test_figure_out_nv_matches_product(built):
    from pyapi import demb200
    for box in [(1.0, 1.0, 1.0), (0.2, 0.2, 0.2), (1.5, 1.3, 1.1), (10.0, 1.0, 0.5), (0.3, 2.0, 0.7), (20, 20, 20)]:
        umin, umax, tmin, tmax = pyoracle.box_domain(*box)
        pmin, pmax, ptmin, ptmax = demb200.host_box_domain(*box)
        assert np.array_equal(tmin, ptmin) and np.array_equal(tmax, ptmax) and np.array_equal(umin, pmin)
        a = pyoracle.figure_out_nv(tmin, tmax)
        b = demb200.host_figure_out_nv(tmin, tmax)
        assert a[:3] == b[:3] and a[3] == b[3] and a[4] == b[4], (box, a, b)
        assert sum(a[:3]) == 64


@needs_ref
@pytest.mark.parametrize("scene", ["mesh_tray", "drum"])
@pytest.mark.parametrize("bin_mult", [1.05, 2.0, 5.0])
def test_facet_candidates_against_reference_triangle_broad_phase(built, bin_mult, scene):
    """Sphere--triangle broad phase pinned to reference code.  The reference sandwiches every facet between two offset,
    enlarged copies (makeTriangleSandwich), registers facets and spheres in bins (getNumberOfBinsEachTriangleTouches /
    populateBinTriangleTouchingPairs with DEMTriangleBoxIntersect.cu; getNumberOfBinsEachSphereTouches) and tests the pairs
    of every shared bin one-sidedly against both copies (triangle_sphere_CD_directional,
    DEMContactKernels_SphereTriangle.cu:196-262), attributing a pair to the bin of its contact point.  All of that runs
    here through the host shim on the oracle's state (ref_sphere_tri_contacts), for three bin sizes, next to the oracle's
    geometric rule "distance to the facet < radius + sphere margin + mesh margin - smaller family extra margin":
      * the oracle's list is exactly the pairs within that reach, measured with the reference's own snap_to_face;
      * what the reference lists beyond it (its one-sided test admits every sphere behind a sandwich copy) is out of
        reach, hence never in touch before the next rebuild;
      * every pair in touch deeper than the mesh owner's margin is in both lists; shallower (grazing) contacts can be
        missing from the reference's list -- the sandwich moves the rim of a facet by up to its margin and the contact
        point may then fall into a bin the two do not share -- and are present in the oracle's."""
    import ctypes as C
    if scene == "drum":  # BASELINE config 4 at oracle scale: polydisperse clumps in a rotating drum of 1500 facets
        f = scenes.flatten(scenes.config4_drum(600, 1500, omega=6.0, init_vel=(0.2, 0.0, -1.0), cd_update_freq=10, spacing=2.7))
        checkpoints = (1200, 600)
    else:
        f = scenes.flatten(_scene("mesh_tray"))
        checkpoints = (1500, 1500, 1000)
    w = pyoracle.world_from_flat(f)
    gapfn = pyoracle.ref().ref_tri_sphere_gap

    def narrow(s, pair):
        out = (C.c_double * 3)()
        hit = gapfn(C.byref(s), C.c_uint32(pair[0]), C.c_uint32(pair[1]), out)
        return bool(hit), out[0] - out[1], -out[2]  # in touch, gap (distance - radius), penetration

    n_ref = n_grazing = 0
    for nsteps in checkpoints:
        w.step(nsteps, cd_every=5)
        w.compute_margins(5)
        w.detect_contacts()
        w.prepare_acc()
        w.calc_forces()
        idA, idB, ct, _ = w.contacts()
        n = w.nContacts
        st = ct == 2  # ORC_SPHERE_MESH
        mine = set(zip(idA[st].tolist(), idB[st].tolist()))
        on = st & (np.abs(w.contactForces[: 3 * n].reshape(-1, 3)).max(1) > 0)
        touching = set(zip(idA[on].tolist(), idB[on].tolist()))
        rmax = float(np.max(w.Radii)) + float(np.max(w.marginSize))
        bin_size = bin_mult * 2.0 * rmax
        ext = [float(2 ** p) * float(w.voxelSize) for p in (w.nvXp2, w.nvYp2, w.nvZp2)]
        nb = [int(np.ceil(e / bin_size)) for e in ext]
        cap = 64 * w.nSpheres + 1024
        oS, oT = np.zeros(cap, "u4"), np.zeros(cap, "u4")
        tri_bins = np.zeros(max(w.nTri, 1), "u4")
        s = w.struct()
        fn = pyoracle.ref().ref_sphere_tri_contacts
        fn.restype = C.c_long
        got = fn(C.byref(s), C.c_double(bin_size), C.c_uint32(nb[0]), C.c_uint32(nb[1]), C.c_uint32(nb[2]),
                 oS.ctypes.data_as(C.c_void_p), oT.ctypes.data_as(C.c_void_p), C.c_long(cap),
                 tri_bins.ctypes.data_as(C.c_void_p))
        assert got >= 0
        theirs = set(zip(oS[:got].tolist(), oT[:got].tolist()))
        assert got == len(theirs), "the reference attributes a pair to exactly one bin"
        # the same through the reference's own block-cooperative per-bin kernels (getNumberOfSphTriContactsEachBin /
        # populateTriSphContactsEachBin, unchanged, every CUDA thread of a block on its own fiber): the very same list
        cS, cT = np.zeros(cap, "u4"), np.zeros(cap, "u4")
        fnc = pyoracle.ref().ref_sphere_tri_contacts_coop
        fnc.restype = C.c_long
        gotc = fnc(C.byref(s), C.c_double(bin_size), C.c_uint32(nb[0]), C.c_uint32(nb[1]), C.c_uint32(nb[2]),
                   cS.ctypes.data_as(C.c_void_p), cT.ctypes.data_as(C.c_void_p), C.c_long(cap))
        assert gotc == got and set(zip(cS[:gotc].tolist(), cT[:gotc].tolist())) == theirs
        assert (tri_bins[: w.nTri] > 0).all(), "every facet lies in at least one bin"

        def reach(pair):  # what the margins of this rebuild promise to cover
            oSph, oTri = w.ownerClumpBody[pair[0]], w.ownerMesh[pair[1]]
            extra = min(float(w.familyExtraMarginSize[w.familyID[oSph]]), float(w.familyExtraMarginSize[w.familyID[oTri]]))
            return float(w.marginSize[oSph]) + float(w.marginSize[oTri]) - extra, float(w.marginSize[oTri])

        for pair in mine:
            hit, gap, pen = narrow(s, pair)
            assert gap < reach(pair)[0] + 1e-9, (pair, gap)
            assert hit == (pair in touching)
        for pair in theirs - mine:
            hit, gap, pen = narrow(s, pair)
            assert not hit and gap >= reach(pair)[0] - 1e-9, (pair, gap, reach(pair))
        for pair in mine - theirs:
            hit, gap, pen = narrow(s, pair)
            if hit:
                assert pen < reach(pair)[1], (pair, pen, reach(pair))
                n_grazing += 1
        deep = {p for p in touching if narrow(s, p)[2] >= reach(p)[1]}
        assert deep <= theirs and touching <= mine
        print("bin %.4f: reference %d pairs, oracle %d (%d in common), in touch %d (%d deeper than the mesh margin); "
              "facets register in %.1f bins on average" % (bin_size, len(theirs), len(mine), len(mine & theirs),
                                                            len(touching), len(deep), tri_bins[: w.nTri].mean()))
        n_ref += len(theirs)
    assert n_ref > 10
    print("grazing contacts the reference's broad phase misses at this bin size:", n_grazing)


def _random_scene(seed):
    """A small scene drawn from a seed: mixed clump templates (1 to 5 spheres of random size and offset), two materials of random
    stiffness / restitution / friction / rolling resistance, random orientations, velocities and spins, a random integrator,
    a random bounding mode plus an inclined plane, random step size and list lifetime."""
    rng = np.random.RandomState(seed)
    s = scenes.Scene()
    mats = [s.load_material(E=float(10 ** rng.uniform(7, 9)), nu=float(rng.uniform(0.2, 0.4)), CoR=float(rng.uniform(0.2, 0.9)),
                            mu=float(rng.uniform(0.0, 0.8)), Crr=float(rng.choice([0.0, 0.0, rng.uniform(0.01, 0.1)])))
            for _ in range(2)]
    scale = 0.004
    types = []
    for _ in range(3):
        ncomp = int(rng.randint(1, 6))
        radii = rng.uniform(0.5, 1.0, ncomp) * scale
        rel = rng.uniform(-0.6, 0.6, (ncomp, 3)) * scale if ncomp > 1 else np.zeros((1, 3))
        mass = float(2.6e3 * 4.19 * scale ** 3 * ncomp * rng.uniform(0.5, 1.0))
        moi = mass * scale ** 2 * rng.uniform(0.3, 0.6, 3)
        types.append(s.load_clump_type(mass, moi, radii, rel, [mats[int(rng.randint(2))] for _ in range(ncomp)]))
    n_side = int(rng.randint(3, 5))
    g = np.stack(np.meshgrid(*[np.arange(n_side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    sep = 3.4 * scale
    pts = (g - (n_side - 1) / 2.0) * sep + rng.uniform(-0.1, 0.1, g.shape) * scale
    half = n_side * sep / 2 + 2.5 * scale
    s.box = (2 * half, 2 * half, 2 * half * 1.3)
    s.bounding, s.bounding_mat = str(rng.choice(["top_open", "all", "only_bottom"])), mats[0]
    n = len(pts)
    s.add_clumps(rng.randint(0, 3, n).astype("i4") + types[0], pts, quat=scenes.random_unit_quats(n, seed + 17),
                 vel=np.asarray(rng.uniform(-1, 1, (n, 3)) + np.array([0, 0, -1.5]), "f4"),
                 omg=np.asarray(rng.uniform(-30, 30, (n, 3)), "f4"))
    nrm = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), 1.0])
    s.add_plane((0.0, 0.0, -half * 0.9), nrm, mats[1])
    if rng.rand() < 0.4:
        # every third scene or so: the grains sit in a tilted, spinning box of triangles instead of the analytical walls
        s.bounding = "none"
        v, fc = scenes.box_mesh(2 * half * 0.95, 2 * half * 0.95, 2 * half * 1.2, n=int(rng.randint(2, 5)), inward=True)
        tilt = float(rng.uniform(-0.3, 0.3))
        s.add_mesh(v, fc, mat=mats[0], mass=1.0, moi=(1, 1, 1), pos=(0.0, 0.0, 0.0),
                   quat=(math.cos(tilt / 2), math.sin(tilt / 2), 0.0, 0.0), family=10)
        s.prescribed[10] = dict(linvel=(0.0, 0.0, 0.0), angvel=(0.0, 0.0, float(rng.uniform(-4, 4))))
    s.h = float(rng.choice([5e-6, 1e-5]))
    s.G = (0.0, 0.0, -9.81)
    s.integrator = int(rng.randint(0, 3))
    s.force_model = 0 if rng.rand() < 0.75 else 1  # DEM_HERTZIAN / DEM_HERTZIAN_FRICTIONLESS
    s.cd_update_freq = int(rng.choice([1, 3, 5, 10]))
    return s


@needs_ref
@pytest.mark.parametrize("seed", list(range(100, 124)))
def test_random_scenes_bit_exact_vs_reference_kernels(built, seed):
    """The oracle against the reference's kernel text on scenes nobody tuned: 24 seeds, 6000 steps each (compared every 1500), every owner
    state word, contact count, history word, force and contact point identical."""
    f = scenes.flatten(_random_scene(seed))
    a = pyoracle.world_from_flat(f, contact_capacity=64 * f.nSpheres + 1024)  # (a pile-up in a corner lists many pairs)
    b = a.copy()
    touched = False
    for _ in range(4):
        a.step(1500, cd_every=f.cd_update_freq)
        b.step(1500, cd_every=f.cd_update_freq, use_ref=True)
        touched = touched or (a.nContacts > 0 and np.abs(a.contactForces[: 3 * a.nContacts]).max() > 0)
        _assert_same_state(a, b)
    assert touched, "nothing collided: the scene tests nothing"
    n = a.nContacts
    for name in ("contactForces", "contactTorque_convToForce", "contactPointGeometryA", "contactPointGeometryB"):
        assert np.array_equal(getattr(a, name)[: 3 * n].view("u4"), getattr(b, name)[: 3 * n].view("u4")), name


@needs_ref
@pytest.mark.parametrize("threads", [2, 5, 16, 64])
@pytest.mark.parametrize("kind", ["clumps_full", "mesh_tray", "cylinder", "families", "drum"])
def test_threaded_broad_phase_lists_the_same_contacts(built, kind, threads):
    """The reference arm of bench.py spreads the oracle's broad phase over the host cores (orc_set_threads; only the OpenMP build
    inside oracle/_ref does): the sphere range is cut into chunks -- also in the middle of a clump --, every chunk is searched and
    sorted on its own, the sorted runs are merged.  The list (pairs, types, carried history words) must be the serial one, entry
    by entry, whatever the number of threads."""
    f = scenes.flatten(_scene(kind))
    a = pyoracle.world_from_flat(f)
    a.step(1500, cd_every=f.cd_update_freq)              # a bed with live contact history
    assert a.nContacts > 0 and max(np.abs(a.contactWildcards[k][: a.nContacts]).max() for k in range(4)) > 0
    a.compute_margins(f.cd_update_freq)
    b = a.copy()
    assert a._call(pyoracle.lib().orc_detect_contacts) == 0
    try:
        pyoracle.ref_set_threads(threads)
        assert b._call(pyoracle.ref().orc_detect_contacts) == 0
    finally:
        pyoracle.ref_set_threads(1)
    n = a.nContacts
    assert n == b.nContacts and n > 0
    for name in ("idGeometryA", "idGeometryB", "contactType"):
        assert np.array_equal(getattr(a, name)[:n], getattr(b, name)[:n]), name
    for k in range(4):
        assert np.array_equal(a.contactWildcards[k][:n].view("u4"), b.contactWildcards[k][:n].view("u4")), "wildcard %d" % k


def _reference_sphere_sphere_sweep(w, bin_size):
    """The reference's own per-bin sweep kernels on the world's current state (ref_sphere_sphere_contacts: the two
    block-cooperative kernels of DEMContactKernels_SphereSphere.cu run on fibers).  Returns (pairs as a list, stats)."""
    import ctypes as C
    ext = [float(2 ** p) * float(w.voxelSize) for p in (w.nvXp2, w.nvYp2, w.nvZp2)]
    nb = [int(np.ceil(e / bin_size)) for e in ext]
    cap = 64 * w.nSpheres + 1024
    oA, oB = np.zeros(cap, "u4"), np.zeros(cap, "u4")
    stats = np.zeros(3, "u8")
    s = w.struct()
    fn = pyoracle.ref().ref_sphere_sphere_contacts
    fn.restype = C.c_long
    got = fn(C.byref(s), C.c_double(bin_size), C.c_uint32(nb[0]), C.c_uint32(nb[1]), C.c_uint32(nb[2]),
             oA.ctypes.data_as(C.c_void_p), oB.ctypes.data_as(C.c_void_p), C.c_long(cap), stats.ctypes.data_as(C.c_void_p))
    assert got >= 0
    return [(min(a, b), max(a, b)) for a, b in zip(oA[:got].tolist(), oB[:got].tolist())], stats


@needs_ref
@pytest.mark.parametrize("kind,bin_mult", [("clumps_full", 1.05), ("clumps_full", 3.0), ("clumps_roll", 1.5), ("mesh_tray", 1.05),
                                           ("families", 2.0), ("drum", 2.0), ("drum", 30.0), ("seed101", 1.2), ("seed107", 2.5),
                                           ("seed113", 1.05), ("seed119", 6.0)])
def test_sphere_sphere_candidates_are_the_reference_sweeps(built, kind, bin_mult):
    """Sphere--sphere broad phase pinned to reference code, end to end.  The reference registers spheres in bins, groups them by
    bin and sweeps every bin with one thread block: shared-memory staging of up to 512 spheres, all pairs of the batch
    spread over the 512 threads, the rest of a fuller bin against the batch, a pair attributed to the bin of its contact point
    (getNumberOfSphereContactsEachBin / populateSphSphContactPairsEachBin, DEMContactKernels_SphereSphere.cu:91-440).  Those two
    kernels run here UNCHANGED -- every CUDA thread of a block on its own fiber, __syncthreads as a real barrier -- behind the
    reference's own sphere -> bin kernels, on states of settling beds, for several bin sizes (one of them puts all 3000 spheres
    of the drum scene into ONE bin, which takes the batch loop and its left-over path).  The list must be the oracle's, pair for
    pair: same-owner and family-masked pairs dropped, every pair once."""
    if kind.startswith("seed"):
        f = scenes.flatten(_random_scene(int(kind[4:])))
        w = pyoracle.world_from_flat(f, contact_capacity=64 * f.nSpheres + 1024)
    else:
        f = scenes.flatten(_scene(kind))
        w = pyoracle.world_from_flat(f)
    total = 0
    for nsteps in (600, 900, 1500):
        w.step(nsteps, cd_every=f.cd_update_freq)
        w.compute_margins(f.cd_update_freq)
        w.detect_contacts()
        idA, idB, ct, _ = w.contacts()
        ss = ct == 1                                        # ORC_SPHERE_SPHERE (oracle/dem_oracle.h)
        mine = sorted(zip(idA[ss].tolist(), idB[ss].tolist()))
        rmax = float(np.max(w.Radii)) + float(np.max(w.marginSize))
        theirs, stats = _reference_sphere_sphere_sweep(w, bin_mult * 2.0 * rmax)
        assert len(theirs) == len(set(theirs)), "the reference attributes a pair to exactly one bin"
        assert sorted(theirs) == mine, (len(theirs), len(mine), sorted(set(mine) ^ set(theirs))[:6])
        assert int(stats[2]) == len(theirs), "the count pass and the fill pass of the reference agree"
        total += len(mine)
        print("%s step +%d: %d pairs in %d active bins (most spheres in a bin: %d)" % (kind, nsteps, len(mine), stats[0], stats[1]))
    assert total > 0


@needs_ref
@pytest.mark.parametrize("kind", ["clumps_full", "clumps_roll", "mesh_tray", "cylinder", "families", "seed104", "seed117"])
def test_history_carry_over_is_the_reference_persistent_map(built, kind):
    """Row a9 against reference code: when a contact list is rebuilt the reference looks every new contact up among the old
    contacts of the same sphere by (idB, type) -- buildPersistentMap, DEMHistoryMappingKernels.cu:17-61 -- and carries the
    history words of the partner over (zeros for a new contact).  That kernel runs here through the shim on the oracle's old
    and new lists (ref_history_map); the words the oracle carried must be exactly the partner's, for every contact of every
    type, on beds where contacts appear and disappear between rebuilds."""
    import ctypes as C
    if kind.startswith("seed"):
        f = scenes.flatten(_random_scene(int(kind[4:])))
        w = pyoracle.world_from_flat(f, contact_capacity=64 * f.nSpheres + 1024)
    else:
        f = scenes.flatten(_scene(kind))
        w = pyoracle.world_from_flat(f)
    fn = pyoracle.ref().ref_history_map
    fn.restype = C.c_int
    carried = fresh = gone = 0
    for nsteps in (700, 400, 400, 800):
        w.step(nsteps, cd_every=f.cd_update_freq)          # (the list is one update period old now: the bed has moved)
        w.step(2 * f.cd_update_freq, cd_every=10 ** 9)     # ... and two more periods without a rebuild
        oA, oB, oT, oW = w.contacts()
        w.compute_margins(f.cd_update_freq)
        w.detect_contacts()
        nA, nB, nT, nW = w.contacts()
        po, pn = np.argsort(oA, kind="stable"), np.argsort(nA, kind="stable")
        mapping = np.zeros(max(len(nA), 1), "u4")
        args = [np.ascontiguousarray(a[p].astype(dt)) for a, p, dt in ((nA, pn, "u4"), (nB, pn, "u4"), (nT, pn, "u1"),
                                                                         (oA, po, "u4"), (oB, po, "u4"), (oT, po, "u1"))]
        rc = fn(C.c_uint32(w.nSpheres), C.c_uint32(len(nA)), *[x.ctypes.data_as(C.c_void_p) for x in args[:3]],
                C.c_uint32(len(oA)), *[x.ctypes.data_as(C.c_void_p) for x in args[3:]], mapping.ctypes.data_as(C.c_void_p))
        assert rc == 0
        m = mapping[: len(nA)]
        found = m != 0xFFFFFFFF
        live = np.zeros(len(nA), bool)
        for k in range(4):
            expect = np.zeros(len(nA), "f4")
            expect[found] = oW[po, k][m[found]]            # (contacts() returns copies: n x 4 history words)
            assert np.array_equal(expect.view("u4"), np.ascontiguousarray(nW[pn, k]).view("u4")), (kind, "wildcard %d" % k)
            live |= nW[pn, k] != 0
        carried += int((found & live).sum())
        fresh += int((~found).sum())
        gone += len(oA) - int(found.sum())
    assert carried > 0 and fresh + gone > 0, (carried, fresh, gone)


@needs_ref
@pytest.mark.parametrize("kind,bin_mult,nsteps", [("clumps_full", 8.0, 1000), ("families", 10.0, 600), ("cylinder", 12.0, 800),
                                                  ("seed109", 8.0, 800)])
def test_trajectory_with_the_reference_rebuild_is_the_oracle_trajectory(built, kind, bin_mult, nsteps):
    """Everything at once: a run in which EVERY kernel is the reference's -- margins, sphere -> bin registration, the per-bin
    sweeps (on fibers), buildPersistentMap for the history words, force, accumulation, integration (ref_step with
    ref_use_reference_rebuild) -- against the oracle stepping on its own, bit for bit after every stretch: owner state, contact
    list, history words.  (Sphere scenes; a rebuild through fibers costs 512 context switches per bin and barrier, hence the short
    runs and the large bins -- the list does not depend on the bin size, see the sweep test above.)"""
    if kind.startswith("seed"):
        f = scenes.flatten(_random_scene(int(kind[4:])))
        a = pyoracle.world_from_flat(f, contact_capacity=64 * f.nSpheres + 1024)
    else:
        f = scenes.flatten(_scene(kind))
        a = pyoracle.world_from_flat(f)
    if a.nTri:
        pytest.skip("the reference's facet list legitimately differs from the oracle's (see the facet test)")
    b = a.copy()
    touched = False
    try:
        pyoracle.ref().ref_use_reference_rebuild(C_double(bin_mult))
        for _ in range(2):
            a.step(nsteps // 2, cd_every=f.cd_update_freq)
            b.step(nsteps // 2, cd_every=f.cd_update_freq, use_ref=True)
            _assert_same_state(a, b)
            n = a.nContacts
            for name in ("idGeometryA", "idGeometryB", "contactType"):
                assert np.array_equal(getattr(a, name)[:n], getattr(b, name)[:n]), name
            touched = touched or (n > 0 and np.abs(a.contactForces[: 3 * n]).max() > 0)
    finally:
        pyoracle.ref().ref_use_reference_rebuild(C_double(0.0))
    assert touched


def C_double(x):
    import ctypes
    return ctypes.c_double(x)
