"""Synthetic DEM scenes and the host-side flattening that DEMSolver::Initialize performs.

`Scene` is the user-level description (materials, clump templates, clumps, analytical boundaries, solver settings)
mirroring the calls a reference demo script makes (LoadMaterial / LoadClumpType / AddClumps / InstructBoxDomain* ...).
`flatten(scene)` produces the reference's flattened SoA arrays (owners ordered clumps, analytical objects, meshes;
src/DEM/dT.cpp:638-1024 populateEntityArrays of the reference) as a `FlatWorld`, which both the CUDA engine
(pyapi.demb200.Engine.load_flat) and the CPU checker used by the tests consume.

World sizing uses the product's own host routines (dem_host_box_domain / dem_host_figure_out_nv /
dem_host_encode_positions of libdemcore.so).
"""
import ctypes as C
import math
from types import SimpleNamespace

import numpy as np

from . import demb200 as D

RESERVED_FAMILY = 255
ANAL_PLANE, ANAL_CYL_INF = 0, 2

# data/clumps/3_clump.csv of the reference (x,y,z,r) -- three overlapping spheres
CLUMP3 = np.array([[0.5, 0.341729, 0.0, 0.8], [0.0, -0.658271, 0.0, 0.8], [-0.5, 0.341729, 0.0, 0.8]], "f4")
CLUMP3_VOLUME = 5.5886717
CLUMP3_MOI = (2.928, 2.6029, 3.9908)  # as DEMdemo_Mixer.cpp:68-72 uses them


class Scene:
    def __init__(self):
        self.materials = []          # dicts E, nu, CoR, mu, Crr
        self.material_pairs = {}     # (prop, i, j) -> value
        self.templates = []          # dicts mass, moi(3), radii, relpos(n,3), mats(n)
        self.clump_type = np.zeros(0, "i4")
        self.clump_xyz = np.zeros((0, 3), "f4")
        self.clump_vel = np.zeros((0, 3), "f4")
        self.clump_omg = np.zeros((0, 3), "f4")
        self.clump_quat = np.zeros((0, 4), "f4")   # w,x,y,z
        self.clump_family = np.zeros(0, "u1")
        self.meshes = []             # dicts family, mass, moi, pos, quat(wxyz), vel, omg, verts(n,3), faces(m,3), mat
        self.ext_objs = []           # dicts family, mass, moi, pos, quat(wxyz), comps=[dict(type,pos,dir,size1,normal,mat)]
        self.box = (1.0, 1.0, 1.0)
        self.bounding = "none"       # none | all | top_open | only_bottom | only_sides
        self.bounding_mat = 0
        self.h = 1e-5
        self.G = (0.0, 0.0, -9.81)
        self.force_model = D.HERTZIAN
        self.integrator = D.EXTENDED_TAYLOR
        self.cd_update_freq = 20
        self.beta = -1.0             # <0: velocity based margin
        self.approxMaxVel = 1e15
        self.expSafetyMulti = 1.0
        self.expSafetyAdder = 3.0    # m_expand_base_vel default, API.h:1484
        self.errOutVel = 1e3
        self.record_contact_forces = 0
        self.fixed_families = [RESERVED_FAMILY]
        self.disabled_pairs = []
        self.prescribed = {}         # family -> dict(linvel=(..), angvel=(..), dictate=True)

    # -- LoadMaterial / LoadClumpType / LoadSphereType -------------------------------------------------------------
    def load_material(self, **props):
        self.materials.append(dict(props))
        return len(self.materials) - 1

    def load_clump_type(self, mass, moi, radii, relpos, mat):
        radii = np.asarray(radii, "f4").reshape(-1)
        relpos = np.asarray(relpos, "f4").reshape(-1, 3)
        mats = np.full(len(radii), mat, "u2") if np.isscalar(mat) else np.asarray(mat, "u2")
        self.templates.append(dict(mass=np.float32(mass), moi=np.asarray(moi, "f4"), radii=radii, relpos=relpos, mats=mats))
        return len(self.templates) - 1

    def load_sphere_type(self, mass, radius, mat):
        moi = 2.0 / 5.0 * mass * radius * radius  # LoadSphereType, API.h:386
        return self.load_clump_type(mass, (moi, moi, moi), [radius], [[0, 0, 0]], mat)

    def add_clumps(self, types, xyz, vel=None, omg=None, quat=None, family=0):
        xyz = np.asarray(xyz, "f4").reshape(-1, 3)
        n = len(xyz)
        types = np.full(n, types, "i4") if np.isscalar(types) else np.asarray(types, "i4")
        self.clump_type = np.concatenate([self.clump_type, types])
        self.clump_xyz = np.concatenate([self.clump_xyz, xyz])
        self.clump_vel = np.concatenate([self.clump_vel, np.zeros((n, 3), "f4") if vel is None else np.broadcast_to(np.asarray(vel, "f4"), (n, 3))])
        self.clump_omg = np.concatenate([self.clump_omg, np.zeros((n, 3), "f4") if omg is None else np.broadcast_to(np.asarray(omg, "f4"), (n, 3))])
        q = np.tile(np.array([1, 0, 0, 0], "f4"), (n, 1)) if quat is None else np.asarray(quat, "f4").reshape(n, 4)
        self.clump_quat = np.concatenate([self.clump_quat, q])
        fam = np.full(n, family, "u1") if np.isscalar(family) else np.asarray(family, "u1")
        self.clump_family = np.concatenate([self.clump_family, fam])

    def add_plane(self, pos, normal, mat, family=RESERVED_FAMILY):
        nrm = np.asarray(normal, "f4")
        nrm = nrm / np.float32(math.sqrt(float(np.dot(nrm, nrm))))
        self.ext_objs.append(dict(family=family, mass=1e6, moi=(1e6, 1e6, 1e6), pos=(0, 0, 0), quat=(1, 0, 0, 0),
                                  comps=[dict(type=ANAL_PLANE, pos=pos, dir=nrm, size1=0.0, normal=0.0, mat=mat)]))
        return len(self.ext_objs) - 1

    def add_cylinder(self, pos, axis, rad, mat, normal=0.0, family=RESERVED_FAMILY):
        ax = np.asarray(axis, "f4")
        ax = ax / np.float32(math.sqrt(float(np.dot(ax, ax))))
        self.ext_objs.append(dict(family=family, mass=1e6, moi=(1e6, 1e6, 1e6), pos=(0, 0, 0), quat=(1, 0, 0, 0),
                                  comps=[dict(type=ANAL_CYL_INF, pos=pos, dir=ax, size1=rad, normal=normal, mat=mat)]))
        return len(self.ext_objs) - 1


    def add_mesh(self, verts, faces, mat, mass=1.0, moi=(1.0, 1.0, 1.0), pos=(0, 0, 0), quat=(1, 0, 0, 0),
                 family=RESERVED_FAMILY, vel=(0, 0, 0), omg=(0, 0, 0)):
        """AddWavefrontMeshObject / AddMesh + SetInitPos/SetInitQuat/SetFamily/SetMass/SetMOI (API.h:629-655): vertices
        in the mesh frame, counter-clockwise faces (right-hand-rule normal points at the contact side)."""
        self.meshes.append(dict(verts=np.asarray(verts, "f4").reshape(-1, 3), faces=np.asarray(faces, "i8").reshape(-1, 3),
                                mat=mat, mass=mass, moi=moi, pos=pos, quat=quat, family=family, vel=vel, omg=omg))
        return len(self.meshes) - 1


def _bounding_box_planes(scene, umin, umax):
    """addWorldBoundingBox, src/DEM/APIPrivate.cpp:955-1014 of the reference: ONE external object with up to 6 planes."""
    mode = scene.bounding
    if mode == "none":
        return None
    bottom = mode in ("only_bottom", "top_open", "all")
    sides = mode in ("only_sides", "top_open", "all")
    top = mode == "all"
    c = ((umin + umax) / np.float32(2.0)).astype("f4")
    comps = []
    mk = lambda pos, n: dict(type=ANAL_PLANE, pos=np.asarray(pos, "f4"), dir=np.asarray(n, "f4"), size1=0.0, normal=0.0,
                             mat=scene.bounding_mat)
    if bottom:
        comps.append(mk((c[0], c[1], umin[2]), (0, 0, 1)))
    if sides:
        comps.append(mk((umin[0], c[1], c[2]), (1, 0, 0)))
        comps.append(mk((umax[0], c[1], c[2]), (-1, 0, 0)))
        comps.append(mk((c[0], umin[1], c[2]), (0, 1, 0)))
        comps.append(mk((c[0], umax[1], c[2]), (0, -1, 0)))
    if top:
        comps.append(mk((c[0], c[1], umax[2]), (0, 0, -1)))
    return dict(family=RESERVED_FAMILY, mass=1e6, moi=(1e6, 1e6, 1e6), pos=(0, 0, 0), quat=(1, 0, 0, 0), comps=comps)


def flatten(scene):
    f = SimpleNamespace()
    umin, umax, tmin, tmax = D.host_box_domain(*[float(v) for v in scene.box])
    f.userBoxMin, f.userBoxMax = umin, umax
    f.nvXp2, f.nvYp2, f.nvZp2, f.l, f.voxelSize = D.host_figure_out_nv(tmin, tmax)
    f.LBF = tmin.copy()
    f.G = np.asarray(scene.G, "f4")
    f.h = np.float32(scene.h)
    f.integrator, f.force_model, f.cd_update_freq = scene.integrator, scene.force_model, int(scene.cd_update_freq)
    f.beta, f.approxMaxVel = np.float32(scene.beta), np.float32(scene.approxMaxVel)
    f.expSafetyMulti, f.expSafetyAdder = np.float32(scene.expSafetyMulti), np.float32(scene.expSafetyAdder)
    f.errOutVel, f.record_contact_forces = np.float32(scene.errOutVel), int(scene.record_contact_forces)

    ext = list(scene.ext_objs)
    bb = _bounding_box_planes(scene, umin, umax)
    if bb is not None:
        ext.append(bb)  # added at Initialize, i.e. after the user's own external objects

    # ---- templates ----
    comp_start, radii, rel = [], [], []
    for t in scene.templates:
        comp_start.append(len(radii))
        radii.extend(t["radii"].tolist())
        rel.extend(t["relpos"].tolist())
    f.nComp = len(radii)
    f.Radii = np.asarray(radii, "f4")
    relarr = np.asarray(rel, "f4").reshape(-1, 3)
    f.CDRelPosX, f.CDRelPosY, f.CDRelPosZ = (np.ascontiguousarray(relarr[:, k]) for k in range(3))
    # mass properties: clump templates, then external objects (then meshes)
    meshes = list(scene.meshes)
    mass = [t["mass"] for t in scene.templates] + [e["mass"] for e in ext] + [m["mass"] for m in meshes]
    moi = [t["moi"] for t in scene.templates] + [e["moi"] for e in ext] + [m["moi"] for m in meshes]
    f.nMassProps = len(mass)
    f.MassProperties = np.asarray(mass, "f4")
    moi = np.asarray(moi, "f4").reshape(-1, 3)
    f.moiX, f.moiY, f.moiZ = (np.ascontiguousarray(moi[:, k]) for k in range(3))

    # ---- materials (equipMaterials, APIPrivate.cpp:1877-2026) ----
    n = len(scene.materials)
    f.nMat = n
    f.E = np.array([m.get("E", 0.0) for m in scene.materials], "f4")
    f.nu = np.array([m.get("nu", 0.0) for m in scene.materials], "f4")
    for prop in ("CoR", "mu", "Crr"):
        t = np.zeros((n, n), "f4")
        for i, m in enumerate(scene.materials):
            t[i, i] = np.float32(m.get(prop, 0.0))
        for i in range(n):
            for j in range(n):
                if i != j:
                    t[i, j] = np.float32((float(t[i, i]) + float(t[j, j])) / 2.0)
        for (p, i, j), v in scene.material_pairs.items():
            if p == prop:
                t[i, j] = t[j, i] = np.float32(v)
        setattr(f, prop, np.ascontiguousarray(t.reshape(-1)))

    # ---- owners: clumps, then analytical objects, then meshes ----
    nC, nE, nM = len(scene.clump_type), len(ext), len(meshes)
    f.nOwners = nC + nE + nM
    f.nClumps = nC
    f.nMeshes = nM
    xyz = np.concatenate([scene.clump_xyz, np.asarray([e["pos"] for e in ext], "f4").reshape(-1, 3),
                          np.asarray([m["pos"] for m in meshes], "f4").reshape(-1, 3)]).astype("f4")
    p = D.DemSimParams()
    p.nvXp2, p.nvYp2, p.nvZp2, p.l, p.voxelSize = f.nvXp2, f.nvYp2, f.nvZp2, f.l, f.voxelSize
    for k in range(3):
        p.LBF[k] = float(f.LBF[k])
    f.voxelID, f.locX, f.locY, f.locZ = (np.zeros(max(f.nOwners, 1), dt) for dt in ("u8", "u2", "u2", "u2"))
    xyzc = np.ascontiguousarray(xyz)
    D.load_library().dem_host_encode_positions(C.byref(p), D._p(xyzc), C.c_uint64(f.nOwners), D._p(f.voxelID),
                                               D._p(f.locX), D._p(f.locY), D._p(f.locZ))
    quat = np.concatenate([scene.clump_quat, np.asarray([e["quat"] for e in ext], "f4").reshape(-1, 4),
                           np.asarray([m["quat"] for m in meshes], "f4").reshape(-1, 4)]).astype("f4")
    f.oriQw, f.oriQx, f.oriQy, f.oriQz = (np.ascontiguousarray(quat[:, k]) if len(quat) else np.zeros(1, "f4") for k in range(4))
    vel = np.concatenate([scene.clump_vel, np.zeros((nE, 3), "f4"), np.asarray([m["vel"] for m in meshes], "f4").reshape(-1, 3)]).astype("f4")
    omg = np.concatenate([scene.clump_omg, np.zeros((nE, 3), "f4"), np.asarray([m["omg"] for m in meshes], "f4").reshape(-1, 3)]).astype("f4")
    f.vX, f.vY, f.vZ = (np.ascontiguousarray(vel[:, k]) if len(vel) else np.zeros(1, "f4") for k in range(3))
    f.omgBarX, f.omgBarY, f.omgBarZ = (np.ascontiguousarray(omg[:, k]) if len(omg) else np.zeros(1, "f4") for k in range(3))
    f.familyID = np.concatenate([scene.clump_family, np.asarray([e["family"] for e in ext], "u1"),
                                 np.asarray([m["family"] for m in meshes], "u1")]).astype("u1")
    if len(f.familyID) == 0:
        f.familyID = np.zeros(1, "u1")
    nT = len(scene.templates)
    f.inertiaPropOffsets = np.concatenate([scene.clump_type.astype("u2"), (nT + np.arange(nE + nM)).astype("u2")]).astype("u2")
    if len(f.inertiaPropOffsets) == 0:
        f.inertiaPropOffsets = np.zeros(1, "u2")

    # ---- spheres ----
    ncomp_of = np.array([len(t["radii"]) for t in scene.templates], "i8")
    counts = ncomp_of[scene.clump_type] if nC else np.zeros(0, "i8")
    f.nSpheres = int(counts.sum())
    f.ownerClumpBody = np.repeat(np.arange(nC, dtype="u4"), counts).astype("u4")
    starts = np.asarray(comp_start, "i8")
    first = np.cumsum(counts) - counts
    within = np.arange(f.nSpheres, dtype="i8") - np.repeat(first, counts)
    f.clumpComponentOffset = (np.repeat(starts[scene.clump_type], counts) + within).astype("u2") if nC else np.zeros(1, "u2")
    allmats = np.concatenate([t["mats"] for t in scene.templates]).astype("u2") if nT else np.zeros(1, "u2")
    f.sphereMaterialOffset = allmats[f.clumpComponentOffset].astype("u2") if nC else np.zeros(1, "u2")
    if f.nSpheres == 0:
        f.ownerClumpBody = np.zeros(1, "u4")

    # ---- analytical components ----
    comps = [(nC + i, c) for i, e in enumerate(ext) for c in e["comps"]]
    f.nAnal = len(comps)
    g = lambda fn, dt: np.asarray([fn(o, c) for o, c in comps], dt) if comps else np.zeros(1, dt)
    f.objOwner = g(lambda o, c: o, "u4")
    f.objType = g(lambda o, c: c["type"], "u1")
    f.objMaterial = g(lambda o, c: c["mat"], "u2")
    f.objNormal = g(lambda o, c: c["normal"], "f4")
    f.objRelPosX, f.objRelPosY, f.objRelPosZ = (g(lambda o, c, k=k: c["pos"][k], "f4") for k in range(3))
    f.objRotX, f.objRotY, f.objRotZ = (g(lambda o, c, k=k: c["dir"][k], "f4") for k in range(3))
    f.objSize1 = g(lambda o, c: c["size1"], "f4")
    f.objSize2 = np.zeros(max(f.nAnal, 1), "f4")
    f.objSize3 = np.zeros(max(f.nAnal, 1), "f4")
    f.objMass = g(lambda o, c: ext[o - nC]["mass"], "f4")

    # ---- triangles (dT.cpp:960-1010: facets of all meshes back to back, nodes in the mesh frame) ----
    f.nTri = int(sum(len(m["faces"]) for m in meshes))
    if f.nTri:
        f.ownerMesh = np.concatenate([np.full(len(m["faces"]), nC + nE + i, "u4") for i, m in enumerate(meshes)]).astype("u4")
        f.triMaterialOffset = np.concatenate([np.full(len(m["faces"]), m["mat"], "u2") for m in meshes]).astype("u2")
        for k, name in enumerate(("relPosNode1", "relPosNode2", "relPosNode3")):
            setattr(f, name, np.ascontiguousarray(np.concatenate([m["verts"][m["faces"][:, k]] for m in meshes]).astype("f4")))

    # ---- families ----
    f.familyMasks = np.zeros(D.PRESC_DTYPE.itemsize and 32896, "u1")
    for a, b in scene.disabled_pairs:
        i, j = (a, b) if a <= b else (b, a)
        f.familyMasks[(1 + j) * j // 2 + i] = 1
    f.familyExtraMarginSize = np.zeros(256, "f4")
    f.prescriptions = np.zeros(256, D.PRESC_DTYPE)
    for fam in scene.fixed_families:
        pr = f.prescriptions[fam]
        pr["used"] = 1
        for k in ("linVelPrescribed", "rotVelPrescribed", "linPosPrescribed", "hasLinVel", "hasRotVel"):
            pr[k] = 1
        pr["rotPosPrescribed"] = 1
    for fam, d in scene.prescribed.items():
        pr = f.prescriptions[fam]
        pr["used"] = 1
        dictate = 1 if d.get("dictate", True) else 0
        for key, has, val, flag in (("linvel", "hasLinVel", "linVel", "linVelPrescribed"),
                                    ("angvel", "hasRotVel", "rotVel", "rotVelPrescribed")):
            if key in d:
                for k in range(3):
                    if d[key][k] is not None:
                        pr[has][k] = 1
                        pr[val][k] = d[key][k]
                        pr[flag][k] = dictate
    return f


# -----------------------------------------------------------------------------------------------------------------
# tessellated test geometry
def plate_mesh(sx, sy, nx, ny, z=0.0):
    """A flat plate [-sx/2,sx/2]x[-sy/2,sy/2] at height z, nx*ny*2 facets, normals +z."""
    xs = np.linspace(-sx / 2, sx / 2, nx + 1, dtype="f8")
    ys = np.linspace(-sy / 2, sy / 2, ny + 1, dtype="f8")
    X, Y = np.meshgrid(xs, ys, indexing="ij")
    verts = np.stack([X.ravel(), Y.ravel(), np.full(X.size, z)], 1).astype("f4")
    idx = lambda i, j: i * (ny + 1) + j
    faces = []
    for i in range(nx):
        for j in range(ny):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            faces += [(a, b, c), (a, c, d)]
    return verts, np.asarray(faces, "i8")


def box_mesh(sx, sy, sz, n=1, inward=True):
    """Closed box surface centred at the origin, each face n*n*2 facets; normals point inward (a container) or outward."""
    V, F = [], []
    half = np.array([sx, sy, sz], "f8") / 2
    for axis in range(3):
        for sign in (-1, 1):
            u, v = [(1, 2), (2, 0), (0, 1)][axis]
            pv, pf = plate_mesh(2 * half[u], 2 * half[v], n, n)
            P = np.zeros((len(pv), 3))
            P[:, u], P[:, v], P[:, axis] = pv[:, 0], pv[:, 1], sign * half[axis]
            # plate normal is +axis (u x v = axis for the cyclic choice above)
            want = -sign if inward else sign
            ff = pf if want > 0 else pf[:, ::-1]
            F.append(ff + sum(len(x) for x in V))
            V.append(P)
    return np.concatenate(V).astype("f4"), np.concatenate(F).astype("i8")


# -----------------------------------------------------------------------------------------------------------------
# samplers (HCPSampler / DEMBoxGridSampler of src/DEM/utils/Samplers.hpp, float arithmetic)
def hcp_box(center, halfsize, sep):
    f = np.float32
    c, s = np.asarray(center, "f4"), np.asarray(halfsize, "f4")
    bl = c - s
    dx = f(sep)
    dy = f(sep) * f(math.sqrt(3.0) / 2)
    dz = f(sep) * f(math.sqrt(2.0 / 3.0))
    nx, ny, nz = int(2 * s[0] / dx) + 1, int(2 * s[1] / dy) + 1, int(2 * s[2] / dz) + 1
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    offy = np.where(k % 2 == 0, f(0), dy / f(3)).astype("f4")
    offx = np.where((j + k) % 2 == 0, f(0), dx / f(2)).astype("f4")
    pts = np.stack([bl[0] + (offx + i.astype("f4") * dx), bl[1] + (offy + j.astype("f4") * dy),
                    bl[2] + k.astype("f4") * dz], -1).astype("f4").reshape(-1, 3)
    eps = 1e-6
    ok = np.all(np.abs(pts - c) <= s + eps, axis=1)
    return pts[ok]


def grid_box(center, halfsize, sep):
    c, s = np.asarray(center, "f4"), np.asarray(halfsize, "f4")
    bl = c - s
    sep = np.broadcast_to(np.asarray(sep, "f4"), (3,))
    n = [int(2 * s[k] / sep[k]) + 1 for k in range(3)]
    i, j, k = np.meshgrid(np.arange(n[0]), np.arange(n[1]), np.arange(n[2]), indexing="ij")
    pts = np.stack([bl[0] + i.astype("f4") * sep[0], bl[1] + j.astype("f4") * sep[1], bl[2] + k.astype("f4") * sep[2]], -1)
    return pts.astype("f4").reshape(-1, 3)


def random_unit_quats(n, seed=4150):
    rng = np.random.RandomState(seed)
    q = rng.normal(size=(n, 4)).astype("f4")
    q /= np.linalg.norm(q, axis=1, keepdims=True).astype("f4")
    return q.astype("f4")


# -----------------------------------------------------------------------------------------------------------------
# BASELINE.json configs
def config1_spheres(n_side=22, h=2e-6, cd_update_freq=10, seed=0, jitter=0.0, box=0.2, drop_height=None):
    """C1: ~10k monodisperse single-sphere grains (r = 1.25 mm, BallDrop materials) on a cubic lattice of spacing 2.01 r
    dropped into a 0.2 x 0.2 m top-open box, frictionless Hertz."""
    s = Scene()
    r = 0.00125
    mat = s.load_material(E=7e7, nu=0.24, CoR=0.9, mu=0.0, Crr=0.0)
    mass = 2500.0 * 4.0 / 3.0 * math.pi * r ** 3
    t = s.load_sphere_type(mass, r, mat)
    s.box = (box, box, box)
    s.bounding, s.bounding_mat = "top_open", mat
    sep = 2.01 * r
    half = (n_side - 1) * sep / 2.0
    zc = -box / 2.0 + r * 1.05 + half if drop_height is None else drop_height
    pts = grid_box((0, 0, zc), (half + 1e-7, half + 1e-7, half + 1e-7), sep)
    if jitter > 0:
        rng = np.random.RandomState(seed)
        pts = pts + (rng.uniform(-1, 1, pts.shape) * jitter * r).astype("f4")
    s.add_clumps(t, pts)
    s.h, s.G = h, (0, 0, -9.81)
    s.force_model = D.HERTZIAN_FRICTIONLESS
    s.cd_update_freq = cd_update_freq
    return s


def config2_clumps(nx, ny, nz, scale=0.005, h=5e-6, cd_update_freq=20, seed=4150, mu=0.2, Crr=0.0,
                   force_model=D.HERTZIAN, spacing=3.0, init_vel=None):
    """C2: clumps of 3_clump.csv scaled to `scale` (mass / MOI as DEMdemo_Mixer.cpp:68-72), HCP-like lattice of spacing
    3*scale with nx*ny*nz sites, random orientations, in a top-open box, Hertz-Mindlin with history."""
    s = Scene()
    mat = s.load_material(E=1e8, nu=0.3, CoR=0.6, mu=mu, Crr=Crr)
    mass = 2.6e3 * CLUMP3_VOLUME * scale ** 3
    moi = np.array(CLUMP3_MOI) * 2.6e3 * scale ** 5
    t = s.load_clump_type(mass, moi, CLUMP3[:, 3] * scale, CLUMP3[:, :3] * scale, mat)
    sep = spacing * scale
    dx, dy, dz = sep, sep * math.sqrt(3.0) / 2, sep * math.sqrt(2.0 / 3.0)
    wall_gap = 2.0 * scale
    bx, by = nx * dx + 2 * wall_gap, ny * dy + 2 * wall_gap
    bz = (nz * dz + 2 * wall_gap) * 1.25
    s.box = (bx, by, bz)
    s.bounding, s.bounding_mat = "top_open", mat
    k, j, i = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    offy = np.where(k % 2 == 0, 0.0, dy / 3)
    offx = np.where((j + k) % 2 == 0, 0.0, dx / 2)
    x = -bx / 2 + wall_gap + 0.25 * dx + offx + i * dx
    y = -by / 2 + wall_gap + 0.25 * dy + offy + j * dy
    z = -bz / 2 + wall_gap + k * dz
    pts = np.stack([x, y, z], -1).reshape(-1, 3).astype("f4")
    s.add_clumps(t, pts, quat=random_unit_quats(len(pts), seed), vel=init_vel)
    s.h, s.G = h, (0, 0, -9.81)
    s.force_model = force_model
    s.cd_update_freq = cd_update_freq
    return s


def drum_mesh(radius, length, n_circ, n_axial, n_radial=None):
    """Closed drum about the y axis, inward-facing facets: mantle n_circ x n_axial quads (2 facets each) and two caps
    tessellated as n_radial rings.  Returns (verts, faces)."""
    n_radial = n_radial or max(1, n_axial // 2)
    th = 2.0 * np.pi * np.arange(n_circ) / n_circ
    ring = np.stack([np.cos(th), np.zeros(n_circ), np.sin(th)], 1)
    V, F = [], []
    ys = np.linspace(-length / 2, length / 2, n_axial + 1)
    for y in ys:
        V.append(ring * radius + np.array([0.0, y, 0.0]))
    idm = lambda a, c: a * n_circ + (c % n_circ)
    for a in range(n_axial):
        for c in range(n_circ):
            p, q, r, s = idm(a, c), idm(a, c + 1), idm(a + 1, c + 1), idm(a + 1, c)
            F += [(p, q, r), (p, r, s)]
    base = (n_axial + 1) * n_circ
    for side, y in ((0, -length / 2), (1, length / 2)):
        start = base
        for k in range(1, n_radial + 1):
            V.append(ring * (radius * k / n_radial) + np.array([0.0, y, 0.0]))
        V.append(np.array([[0.0, y, 0.0]]))
        centre = start + n_radial * n_circ
        idc = lambda k, c: start + (k - 1) * n_circ + (c % n_circ)
        tri = []
        for c in range(n_circ):
            tri.append((centre, idc(1, c + 1), idc(1, c)))
            for k in range(1, n_radial):
                p, q, r, s = idc(k, c), idc(k, c + 1), idc(k + 1, c + 1), idc(k + 1, c)
                tri += [(p, q, r), (p, r, s)]
        tri = np.asarray(tri, "i8")
        if side == 1:
            tri = tri[:, ::-1]
        F += tri.tolist()
        base = centre + 1
    return np.concatenate(V).astype("f4"), np.asarray(F, "i8")


def config4_drum(n_clumps=500000, n_tri=50000, scale=0.004, h=5e-6, cd_update_freq=20, seed=7, omega=3.0, fill=0.45,
                 spacing=2.9, init_vel=None):
    """C4: polydisperse 3-sphere clumps (three sizes, 1 : 0.8 : 0.65) in a rotating drum made of ~n_tri facets; all wall
    contacts go through the sphere--triangle path (no analytical boundary).  Hertz-Mindlin with history."""
    s = Scene()
    mat = s.load_material(E=1e8, nu=0.3, CoR=0.5, mu=0.4, Crr=0.0)
    matw = s.load_material(E=2e8, nu=0.3, CoR=0.5, mu=0.6, Crr=0.0)
    types = []
    for k in (1.0, 0.8, 0.65):
        sc = scale * k
        types.append(s.load_clump_type(2.6e3 * CLUMP3_VOLUME * sc ** 3, np.array(CLUMP3_MOI) * 2.6e3 * sc ** 5,
                                       CLUMP3[:, 3] * sc, CLUMP3[:, :3] * sc, mat))
    sep = spacing * scale
    # drum sized so that the requested number of lattice sites fills `fill` of its height
    vol_site = sep ** 3 / math.sqrt(2.0)
    vol = n_clumps * vol_site / fill
    radius = (vol / (math.pi * 1.2)) ** (1.0 / 3.0)       # length = 1.2 R
    length = 1.2 * radius
    pts = hcp_box((0, 0, 0), (radius, length / 2 - 2.2 * scale, radius), sep)
    rr = np.sqrt(pts[:, 0] ** 2 + pts[:, 2] ** 2)
    pts = pts[rr < radius - 2.2 * scale]
    pts = pts[np.argsort(pts[:, 2], kind="stable")][:n_clumps]     # the lowest sites
    rng = np.random.RandomState(seed)
    ty = np.asarray(types, "i4")[rng.randint(0, 3, len(pts))]
    s.add_clumps(ty, pts, quat=random_unit_quats(len(pts), seed), vel=init_vel)
    # facet count: mantle 2*nc*na + caps 2*nc*(2*nr-1), with na ~ nc*L/(2 pi R), nr ~ nc/(2 pi)
    nc = max(12, int(round(math.sqrt(n_tri / (2 * 1.2 / (2 * math.pi) + 4 / (2 * math.pi))))))
    na = max(1, int(round(nc * 1.2 / (2 * math.pi))))
    nr = max(1, int(round(nc / (2 * math.pi))))
    v, f = drum_mesh(radius, length, nc, na, nr)
    s.add_mesh(v, f, mat=matw, mass=1.0, moi=(1, 1, 1), family=10)
    s.prescribed[10] = dict(linvel=(0.0, 0.0, 0.0), angvel=(0.0, omega, 0.0))
    d = 2.4 * radius
    s.box = (d, 1.4 * length, d)
    s.bounding = "none"
    s.h, s.G = h, (0, 0, -9.81)
    s.cd_update_freq = cd_update_freq
    s.drum_radius, s.drum_length = radius, length
    return s


def config5_spheres(n=5000000, r=0.001, packing=0.5, seed=11, x_range=None):
    """C5: n monodisperse spheres (r = 1 mm) uniformly random in a cubic box at 50 % packing -- the neighbour-search
    stress case (binning + sort only: the spheres overlap at random, no forces are evaluated).  `x_range=(a, b)` keeps
    only the spheres whose x/box fraction lies in [a, b): the pre-partitioned share of one GPU."""
    s = Scene()
    mat = s.load_material(E=1e8, nu=0.3, CoR=0.6, mu=0.0, Crr=0.0)
    t = s.load_sphere_type(2500.0 * 4.0 / 3.0 * math.pi * r ** 3, r, mat)
    side = (n * 4.0 / 3.0 * math.pi * r ** 3 / packing) ** (1.0 / 3.0)
    rng = np.random.RandomState(seed)
    pts = (rng.uniform(-0.5, 0.5, (n, 3)) * side).astype("f4")
    if x_range is not None:
        frac = pts[:, 0] / side + 0.5
        pts = pts[(frac >= x_range[0]) & (frac < x_range[1])]
    s.add_clumps(t, pts)
    s.box = (side * 1.02, side * 1.02, side * 1.02)
    s.bounding = "none"
    s.h, s.G = 2e-6, (0, 0, -9.81)
    s.force_model = D.HERTZIAN_FRICTIONLESS
    s.cd_update_freq = 20
    return s


def write_scene_file(path, scene, xyz=None, quat=None, vel=None, omg=None, scale=0.005):
    """The scene file baseline/run_ref.cpp reads (it rebuilds the same C2 bed through the UNMODIFIED reference's own API):
    "DEMS", u32 version 1, u32 nClumps, f32 box[3], scale, h, E, nu, CoR, mu, Crr, beta, u32 has_vel, then xyz[n][3],
    quat_wxyz[n][4] and, when has_vel, vel[n][3], omgBar[n][3] (all f32, little endian).  xyz / quat / vel / omg default
    to the scene's initial state; pass a downloaded state to hand a settled bed to the reference."""
    import struct
    xyz = np.ascontiguousarray(scene.clump_xyz if xyz is None else xyz, "<f4").reshape(-1, 3)
    n = len(xyz)
    quat = np.ascontiguousarray(scene.clump_quat if quat is None else quat, "<f4").reshape(n, 4)
    has_vel = vel is not None or omg is not None or bool(np.any(scene.clump_vel)) or bool(np.any(scene.clump_omg))
    m = scene.materials[0]
    with open(path, "wb") as fh:
        fh.write(b"DEMS")
        fh.write(struct.pack("<II", 1, n))
        fh.write(struct.pack("<3f", *[float(v) for v in scene.box]))
        fh.write(struct.pack("<8f", float(scale), float(scene.h), float(m["E"]), float(m["nu"]), float(m["CoR"]),
                             float(m["mu"]), float(m.get("Crr", 0.0)), float(scene.beta)))
        fh.write(struct.pack("<I", 1 if has_vel else 0))
        fh.write(xyz.tobytes())
        fh.write(quat.tobytes())
        if has_vel:
            fh.write(np.ascontiguousarray(scene.clump_vel if vel is None else vel, "<f4").reshape(n, 3).tobytes())
            fh.write(np.ascontiguousarray(scene.clump_omg if omg is None else omg, "<f4").reshape(n, 3).tobytes())
    return path
