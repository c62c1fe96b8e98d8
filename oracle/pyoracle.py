"""CPU ORACLE python binding (test infrastructure, NOT product code).

ctypes binding of oracle/liboracle.so (dem_oracle.c, the plain-C restatement of the reference hot path) and,
when present, oracle/_ref/libdemref.so (the reference's own kernel text compiled for the host).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.

`World` owns numpy arrays laid out like the reference SoA contract (src/DEM/Defines.h:269-373 of the reference)
and exposes them to C through the OrcWorld struct of dem_oracle.h.  The host-side set-up arithmetic that the
reference performs in DEMSolver::Initialize (world sizing figureOutNV, src/DEM/APIPrivate.cpp:373-487; material
pair table, :1877-2026) is restated here in python for the same purpose.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

NOT_A_CONTACT, SPHERE_SPHERE, SPHERE_MESH, SPHERE_PLANE, SPHERE_CYL = 0, 1, 2, 11, 13
ANAL_PLANE, ANAL_PLATE, ANAL_CYL_INF = 0, 1, 2
FORWARD_EULER, CENTERED_DIFFERENCE, EXTENDED_TAYLOR = 0, 1, 2
HERTZIAN, HERTZIAN_FRICTIONLESS = 0, 1
NUM_MASKS = 32896
RESERVED_FAMILY = 255


class Prescription(C.Structure):
    _fields_ = [
        ("used", C.c_uint8),
        ("linVelPrescribed", C.c_uint8 * 3),
        ("rotVelPrescribed", C.c_uint8 * 3),
        ("linPosPrescribed", C.c_uint8 * 3),
        ("rotPosPrescribed", C.c_uint8),
        ("hasLinVel", C.c_uint8 * 3),
        ("hasRotVel", C.c_uint8 * 3),
        ("hasLinPos", C.c_uint8 * 3),
        ("hasAcc", C.c_uint8 * 3),
        ("hasAngAcc", C.c_uint8 * 3),
        ("pad_", C.c_uint8 * 2),
        ("linVel", C.c_float * 3),
        ("rotVel", C.c_float * 3),
        ("linPos", C.c_float * 3),
        ("acc", C.c_float * 3),
        ("angAcc", C.c_float * 3),
    ]


PRESC_DTYPE = np.dtype([
    ("used", "u1"), ("linVelPrescribed", "u1", 3), ("rotVelPrescribed", "u1", 3), ("linPosPrescribed", "u1", 3),
    ("rotPosPrescribed", "u1"), ("hasLinVel", "u1", 3), ("hasRotVel", "u1", 3), ("hasLinPos", "u1", 3),
    ("hasAcc", "u1", 3), ("hasAngAcc", "u1", 3), ("pad_", "u1", 2), ("linVel", "f4", 3), ("rotVel", "f4", 3),
    ("linPos", "f4", 3), ("acc", "f4", 3), ("angAcc", "f4", 3)])

_P = C.c_void_p

_WORLD_FIELDS = [
    ("nvXp2", C.c_uint32), ("nvYp2", C.c_uint32), ("nvZp2", C.c_uint32), ("integrator", C.c_uint32),
    ("force_model", C.c_uint32), ("pad0_", C.c_uint32),
    ("l", C.c_double), ("voxelSize", C.c_double), ("timeElapsed", C.c_double),
    ("LBF", C.c_float * 3), ("G", C.c_float * 3), ("h", C.c_float), ("beta", C.c_float),
    ("approxMaxVel", C.c_float), ("expSafetyMulti", C.c_float), ("expSafetyAdder", C.c_float), ("pad1_", C.c_float),
    ("nOwners", C.c_uint32), ("nSpheres", C.c_uint32), ("nTri", C.c_uint32), ("nAnal", C.c_uint32),
    ("nMat", C.c_uint32), ("nComp", C.c_uint32), ("nMassProps", C.c_uint32), ("pad2_", C.c_uint32),
]
# pointer members in declaration order: (name, numpy dtype)
_OWNER_ARRAYS = [("voxelID", "u8"), ("locX", "u2"), ("locY", "u2"), ("locZ", "u2"),
                 ("oriQw", "f4"), ("oriQx", "f4"), ("oriQy", "f4"), ("oriQz", "f4"),
                 ("vX", "f4"), ("vY", "f4"), ("vZ", "f4"),
                 ("omgBarX", "f4"), ("omgBarY", "f4"), ("omgBarZ", "f4"),
                 ("aX", "f4"), ("aY", "f4"), ("aZ", "f4"),
                 ("alphaX", "f4"), ("alphaY", "f4"), ("alphaZ", "f4"),
                 ("familyID", "u1"), ("inertiaPropOffsets", "u2"), ("accSpecified", "u1"), ("angAccSpecified", "u1")]
_SPHERE_ARRAYS = [("ownerClumpBody", "u4"), ("clumpComponentOffset", "u2"), ("sphereMaterialOffset", "u2")]
_COMP_ARRAYS = [("Radii", "f4"), ("CDRelPosX", "f4"), ("CDRelPosY", "f4"), ("CDRelPosZ", "f4")]
_MASS_ARRAYS = [("MassProperties", "f4"), ("moiX", "f4"), ("moiY", "f4"), ("moiZ", "f4")]
_MAT_ARRAYS = [("E", "f4"), ("nu", "f4"), ("CoR", "f4"), ("mu", "f4"), ("Crr", "f4")]
_ANAL_ARRAYS = [("objOwner", "u4"), ("objType", "u1"), ("objMaterial", "u2"), ("objNormal", "f4"),
                ("objRelPosX", "f4"), ("objRelPosY", "f4"), ("objRelPosZ", "f4"),
                ("objRotX", "f4"), ("objRotY", "f4"), ("objRotZ", "f4"),
                ("objSize1", "f4"), ("objSize2", "f4"), ("objSize3", "f4"), ("objMass", "f4")]
_TRI_ARRAYS = [("ownerMesh", "u4"), ("relPosNode1", "f4"), ("relPosNode2", "f4"), ("relPosNode3", "f4"),
               ("triMaterialOffset", "u2")]
_FAM_ARRAYS = [("familyMasks", "u1"), ("familyExtraMarginSize", "f4"), ("prescriptions", None)]
_PTR_GROUP_1 = _OWNER_ARRAYS + _SPHERE_ARRAYS + _COMP_ARRAYS + _MASS_ARRAYS + _MAT_ARRAYS + _ANAL_ARRAYS + \
    _TRI_ARRAYS + _FAM_ARRAYS
_CONTACT_ARRAYS = [("idGeometryA", "u4"), ("idGeometryB", "u4"), ("contactType", "u1")]
_CONTACT_VEC_ARRAYS = [("contactForces", "f4"), ("contactTorque_convToForce", "f4"),
                       ("contactPointGeometryA", "f4"), ("contactPointGeometryB", "f4")]


class OrcWorld(C.Structure):
    _fields_ = (_WORLD_FIELDS + [(n, _P) for n, _ in _PTR_GROUP_1] +
                [("nContacts", C.c_uint64), ("contactCapacity", C.c_uint64)] +
                [(n, _P) for n, _ in _CONTACT_ARRAYS] + [("contactWildcards", _P * 4)] +
                [(n, _P) for n, _ in _CONTACT_VEC_ARRAYS] + [("marginSize", _P)])


def build(force=False):
    """Compile liboracle.so (and _ref/libdemref.so when /root/reference exists)."""
    lib = os.path.join(_HERE, "liboracle.so")
    if force or not os.path.exists(lib) or os.path.isdir("/root/reference/src/kernel"):
        subprocess.run(["make", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
    return lib


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _lib = C.CDLL(path)
        _lib.orc_sizeof_world.restype = C.c_size_t
        _lib.orc_sizeof_prescription.restype = C.c_size_t
        assert _lib.orc_sizeof_world() == C.sizeof(OrcWorld), (_lib.orc_sizeof_world(), C.sizeof(OrcWorld))
        assert _lib.orc_sizeof_prescription() == C.sizeof(Prescription) == PRESC_DTYPE.itemsize
        _lib.orc_step.restype = C.c_int
        _lib.orc_detect_contacts.restype = C.c_int
    return _lib


def ref():
    """The host-compiled reference kernels, or None when oracle/_ref/libdemref.so was never built."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libdemref.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        _ref.ref_step.restype = C.c_int
        _ref.ref_sphere_anal_contacts.restype = C.c_long
        _ref.ref_calc_contact_point.restype = C.c_int
        _ref.ref_set_threads(1)
    return _ref


def ref_set_threads(n):
    """Host threads the reference kernels are spread over (1 = serial, deterministic summation order)."""
    if ref() is not None:
        ref().ref_set_threads(int(n))


def _ptr(a):
    return a.ctypes.data_as(_P) if a is not None else None


class World:
    """Numpy-backed DEM world in the reference's SoA layout."""

    def __init__(self, n_owners, n_spheres, n_comp, n_massprops, n_mat, n_anal=0, n_tri=0, contact_capacity=None):
        self.nOwners, self.nSpheres, self.nComp = int(n_owners), int(n_spheres), int(n_comp)
        self.nMassProps, self.nMat, self.nAnal, self.nTri = int(n_massprops), int(n_mat), int(n_anal), int(n_tri)
        self.nvXp2 = self.nvYp2 = self.nvZp2 = 0
        self.l = self.voxelSize = 0.0
        self.timeElapsed = 0.0
        self.LBF = np.zeros(3, "f4")
        self.G = np.zeros(3, "f4")
        self.h = np.float32(0)
        self.beta = np.float32(-1.0)
        self.approxMaxVel = np.float32(1e15)
        self.expSafetyMulti = np.float32(1.0)
        self.expSafetyAdder = np.float32(3.0)  # m_expand_base_vel, API.h:1484
        self.integrator = EXTENDED_TAYLOR
        self.force_model = HERTZIAN
        self.userBoxMin = np.zeros(3, "f4")
        self.userBoxMax = np.zeros(3, "f4")
        z = lambda n, dt: np.zeros(max(int(n), 1), dt)
        for name, dt in _OWNER_ARRAYS:
            setattr(self, name, z(n_owners, dt))
        self.oriQw[:] = 1.0
        for name, dt in _SPHERE_ARRAYS:
            setattr(self, name, z(n_spheres, dt))
        for name, dt in _COMP_ARRAYS:
            setattr(self, name, z(n_comp, dt))
        for name, dt in _MASS_ARRAYS:
            setattr(self, name, z(n_massprops, dt))
        self.E, self.nu = z(n_mat, "f4"), z(n_mat, "f4")
        self.CoR, self.mu, self.Crr = z(n_mat * n_mat, "f4"), z(n_mat * n_mat, "f4"), z(n_mat * n_mat, "f4")
        for name, dt in _ANAL_ARRAYS:
            setattr(self, name, z(n_anal, dt))
        self.ownerMesh = z(n_tri, "u4")
        self.relPosNode1, self.relPosNode2, self.relPosNode3 = z(3 * n_tri, "f4"), z(3 * n_tri, "f4"), z(3 * n_tri, "f4")
        self.triMaterialOffset = z(n_tri, "u2")
        self.familyMasks = np.zeros(NUM_MASKS, "u1")
        self.familyExtraMarginSize = np.zeros(256, "f4")
        self.prescriptions = np.zeros(256, PRESC_DTYPE)
        self.marginSize = z(n_owners, "f4")
        self.nContacts = 0
        self.step_counter = C.c_uint64(0)
        self._alloc_contacts(contact_capacity if contact_capacity else 16 + 12 * self.nSpheres)
        # the reserved family is always fixed (APIPublic.cpp:980-1011)
        self.set_family_fixed(RESERVED_FAMILY)

    # ---- helpers -------------------------------------------------------------------------------------------
    def _alloc_contacts(self, cap):
        self.contactCapacity = int(cap)
        self.idGeometryA = np.zeros(cap, "u4")
        self.idGeometryB = np.zeros(cap, "u4")
        self.contactType = np.zeros(cap, "u1")
        self.contactWildcards = [np.zeros(cap, "f4") for _ in range(4)]
        self.contactForces = np.zeros(3 * cap, "f4")
        self.contactTorque_convToForce = np.zeros(3 * cap, "f4")
        self.contactPointGeometryA = np.zeros(3 * cap, "f4")
        self.contactPointGeometryB = np.zeros(3 * cap, "f4")

    def set_family_fixed(self, fam):
        p = self.prescriptions[fam]
        p["used"] = 1
        for k in ("linVelPrescribed", "rotVelPrescribed", "linPosPrescribed", "hasLinVel", "hasRotVel"):
            p[k] = 1
        p["rotPosPrescribed"] = 1
        p["linVel"] = 0
        p["rotVel"] = 0

    def set_family_prescribed_lin_vel(self, fam, vel, dictate=True):
        """SetFamilyPrescribedLinVel with numeric-constant strings (API.h:704-775); None == "none"."""
        p = self.prescriptions[fam]
        p["used"] = 1
        for k in range(3):
            if vel[k] is not None:
                p["hasLinVel"][k] = 1
                p["linVel"][k] = vel[k]
                p["linVelPrescribed"][k] = 1 if dictate else 0
        # a family with prescribed lin vel has its rot vel prescribed too unless set otherwise? No: independent.

    def set_family_prescribed_ang_vel(self, fam, omg, dictate=True):
        p = self.prescriptions[fam]
        p["used"] = 1
        for k in range(3):
            if omg[k] is not None:
                p["hasRotVel"][k] = 1
                p["rotVel"][k] = omg[k]
                p["rotVelPrescribed"][k] = 1 if dictate else 0

    def disable_contact_between_families(self, a, b):
        i, j = (a, b) if a <= b else (b, a)
        self.familyMasks[(1 + j) * j // 2 + i] = 1

    def struct(self):
        s = OrcWorld()
        for name in ("nvXp2", "nvYp2", "nvZp2", "integrator", "force_model", "nOwners", "nSpheres", "nTri", "nAnal",
                     "nMat", "nComp", "nMassProps"):
            setattr(s, name, int(getattr(self, name)))
        s.l, s.voxelSize, s.timeElapsed = float(self.l), float(self.voxelSize), float(self.timeElapsed)
        for k in range(3):
            s.LBF[k] = float(self.LBF[k])
            s.G[k] = float(self.G[k])
        s.h, s.beta = float(self.h), float(self.beta)
        s.approxMaxVel, s.expSafetyMulti, s.expSafetyAdder = (float(self.approxMaxVel), float(self.expSafetyMulti),
                                                             float(self.expSafetyAdder))
        for name, _ in _PTR_GROUP_1:
            setattr(s, name, _ptr(getattr(self, name)))
        s.nContacts, s.contactCapacity = int(self.nContacts), int(self.contactCapacity)
        for name, _ in _CONTACT_ARRAYS + _CONTACT_VEC_ARRAYS:
            setattr(s, name, _ptr(getattr(self, name)))
        for k in range(4):
            s.contactWildcards[k] = _ptr(self.contactWildcards[k])
        s.marginSize = _ptr(self.marginSize)
        return s

    def _call(self, fn, *args):
        s = self.struct()
        rc = fn(C.byref(s), *args)
        self.nContacts = int(s.nContacts)
        self.timeElapsed = float(s.timeElapsed)
        return rc

    # ---- oracle entry points -------------------------------------------------------------------------------
    def encode_positions(self, xyz, first=0):
        xyz = np.ascontiguousarray(xyz, "f4").reshape(-1, 3)
        self._call(lib().orc_encode_positions, _ptr(xyz), C.c_uint32(first), C.c_uint32(len(xyz)))

    def decode_positions(self, first=0, n=None):
        n = self.nOwners - first if n is None else n
        out = np.zeros((n, 3), "f4")
        self._call(lib().orc_decode_positions, _ptr(out), C.c_uint32(first), C.c_uint32(n))
        return out

    def positions_f64(self):
        """Decoded world positions in double (for tight comparisons)."""
        vx = self.voxelID & np.uint64((1 << self.nvXp2) - 1)
        vy = (self.voxelID >> np.uint64(self.nvXp2)) & np.uint64((1 << self.nvYp2) - 1)
        vz = self.voxelID >> np.uint64(self.nvXp2 + self.nvYp2)
        out = np.empty((len(self.voxelID), 3), "f8")
        out[:, 0] = vx.astype("f8") * self.voxelSize + self.locX.astype("f8") * self.l + float(self.LBF[0])
        out[:, 1] = vy.astype("f8") * self.voxelSize + self.locY.astype("f8") * self.l + float(self.LBF[1])
        out[:, 2] = vz.astype("f8") * self.voxelSize + self.locZ.astype("f8") * self.l + float(self.LBF[2])
        return out[:self.nOwners]

    def compute_margins(self, max_drift, use_ref=False):
        self._call((ref().ref_compute_margins if use_ref else lib().orc_compute_margins), C.c_uint32(max_drift))

    def detect_contacts(self):
        rc = self._call(lib().orc_detect_contacts)
        if rc != 0:
            raise RuntimeError("oracle contact capacity exceeded")

    def prepare_acc(self, use_ref=False):
        self._call(ref().ref_prepare_acc if use_ref else lib().orc_prepare_acc)

    def calc_forces(self, use_ref=False):
        self._call(ref().ref_calc_forces if use_ref else lib().orc_calc_forces)

    def force_to_acc(self, use_ref=False):
        self._call(ref().ref_force_to_acc if use_ref else lib().orc_force_to_acc)

    def integrate(self, use_ref=False):
        self._call(ref().ref_integrate if use_ref else lib().orc_integrate)

    def step(self, nsteps, cd_every=1, use_ref=False):
        fn = ref().ref_step if use_ref else lib().orc_step
        rc = self._call(fn, C.c_uint32(nsteps), C.c_uint32(cd_every), C.byref(self.step_counter))
        if rc != 0:
            raise RuntimeError("oracle step failed rc=%d" % rc)

    def sphere_positions(self):
        xyz = np.zeros((max(self.nSpheres, 1), 3), "f8")
        rad = np.zeros(max(self.nSpheres, 1), "f4")
        self._call(lib().orc_sphere_positions, _ptr(xyz), _ptr(rad))
        return xyz[:self.nSpheres], rad[:self.nSpheres]

    def contacts(self):
        n = self.nContacts
        return (self.idGeometryA[:n].copy(), self.idGeometryB[:n].copy(), self.contactType[:n].copy(),
                np.stack([w[:n] for w in self.contactWildcards], 1).copy())

    def copy(self):
        import copy as _c
        w = _c.copy(self)
        for k, v in self.__dict__.items():
            if isinstance(v, np.ndarray):
                setattr(w, k, v.copy())
        w.contactWildcards = [a.copy() for a in self.contactWildcards]
        w.step_counter = C.c_uint64(self.step_counter.value)
        return w


# ---------------------------------------------------------------------------------------------------------------
# Host-side set-up arithmetic of the reference, restated.
# ---------------------------------------------------------------------------------------------------------------
def figure_out_nv(box_min, box_max, exact_dir=None):
    """DEMSolver::figureOutNV (src/DEM/APIPrivate.cpp:373-487), m_box_dir_length_is_exact == NONE branch.
    Returns (nvXp2, nvYp2, nvZp2, l, voxelSize)."""
    assert exact_dir is None
    size = [np.float32(box_max[k]) - np.float32(box_min[k]) for k in range(3)]
    XYZ = [np.float32(s) for s in size]
    rank = [0, 1, 2]
    for i in range(2):
        for j in range(i + 1, 3):
            if XYZ[i] > XYZ[j]:
                XYZ[i], XYZ[j] = XYZ[j], XYZ[i]
                rank[i], rank[j] = rank[j], rank[i]
    user321 = [float(XYZ[0]), float(XYZ[1]), float(XYZ[2])]
    more = [0, 0]
    # "XYZ[0] *= 2." on a float array: float * double -> stored back as float
    while XYZ[0] < XYZ[1]:
        if math.sqrt(2.0) * float(XYZ[0]) > float(XYZ[1]):
            break
        more[0] += 1
        XYZ[0] = np.float32(float(XYZ[0]) * 2.0)
    while XYZ[1] < XYZ[2]:
        if math.sqrt(2.0) * float(XYZ[1]) > float(XYZ[2]):
            break
        more[1] += 1
        XYZ[1] = np.float32(float(XYZ[1]) * 2.0)
    total = 64 - 2 * more[0] - more[1]
    base, left = total // 3, total % 3
    b3 = base
    b2 = b3 + more[0]
    b1 = b2 + more[1]
    while left > 0:
        if b3 < b2:
            b3 += 1
        elif b2 < b1:
            b2 += 1
        else:
            b1 += 1
        left -= 1
    bits = [b3, b2, b1]
    l3 = user321[0] / 2.0 ** 16 / 2.0 ** b3
    l2 = user321[1] / 2.0 ** 16 / 2.0 ** b2
    l1 = user321[2] / 2.0 ** 16 / 2.0 ** b1
    l = max(l3, l2, l1)
    nv = [0, 0, 0]
    for pos, axis in enumerate(rank):
        nv[axis] = bits[pos]
    voxel = float(1 << 16) * l
    return nv[0], nv[1], nv[2], l, voxel


def box_domain(x, y, z):
    """InstructBoxDomainDimension(x,y,z) (src/DEM/APIPublic.cpp:845-872): user box and the 20%-enlarged target box."""
    f = np.float32
    umin = np.array([f(-x / 2.0), f(-y / 2.0), f(-z / 2.0)], "f4")
    umax = np.array([f(x / 2.0), f(y / 2.0), f(z / 2.0)], "f4")
    enl = np.array([f(x * 0.2 / 2.0), f(y * 0.2 / 2.0), f(z * 0.2 / 2.0)], "f4")
    return umin, umax, (umin - enl).astype("f4"), (umax + enl).astype("f4")


def material_tables(mats, pairs=None):
    """equipMaterials (src/DEM/APIPrivate.cpp:1877-2026): diagonal from each material, off-diagonal = mean unless set.
    mats: list of dicts with E, nu, CoR, mu, Crr; pairs: {(prop,i,j): value}."""
    n = len(mats)
    out = {"E": np.array([m.get("E", 0.0) for m in mats], "f4"), "nu": np.array([m.get("nu", 0.0) for m in mats], "f4")}
    for prop in ("CoR", "mu", "Crr"):
        t = np.zeros((n, n), "f4")
        for i, m in enumerate(mats):
            t[i, i] = np.float32(m.get(prop, 0.0))
        for i in range(n):
            for j in range(n):
                if i != j:
                    t[i, j] = np.float32((float(t[i, i]) + float(t[j, j])) / 2.0)
        for (p, i, j), v in (pairs or {}).items():
            if p == prop:
                t[i, j] = t[j, i] = np.float32(v)
        out[prop] = t.reshape(-1)
    return out


def world_from_flat(f, contact_capacity=None):
    """Build an oracle World from the flattened arrays a scene produces (pyapi.scenes.flatten)."""
    nTri = int(getattr(f, "nTri", 0))
    w = World(f.nOwners, f.nSpheres, f.nComp, f.nMassProps, f.nMat, f.nAnal, nTri, contact_capacity)
    for name in ("nvXp2", "nvYp2", "nvZp2", "l", "voxelSize", "integrator", "force_model"):
        setattr(w, name, getattr(f, name))
    w.LBF[:] = f.LBF
    w.G[:] = f.G
    w.h = np.float32(f.h)
    w.beta, w.approxMaxVel = np.float32(f.beta), np.float32(f.approxMaxVel)
    w.expSafetyMulti, w.expSafetyAdder = np.float32(f.expSafetyMulti), np.float32(f.expSafetyAdder)
    w.userBoxMin, w.userBoxMax = f.userBoxMin.copy(), f.userBoxMax.copy()
    groups = [(_OWNER_ARRAYS, f.nOwners), (_SPHERE_ARRAYS, f.nSpheres), (_COMP_ARRAYS, f.nComp),
              (_MASS_ARRAYS, f.nMassProps), (_ANAL_ARRAYS, f.nAnal)]
    for arrs, n in groups:
        for name, dt in arrs:
            if hasattr(f, name) and n > 0:
                getattr(w, name)[:n] = np.asarray(getattr(f, name))[:n]
    for name, dt in _TRI_ARRAYS:
        if nTri:
            src = np.asarray(getattr(f, name)).reshape(-1)
            getattr(w, name)[:len(src)] = src
    for name in ("E", "nu", "CoR", "mu", "Crr"):
        getattr(w, name)[:] = getattr(f, name)
    w.familyMasks[:] = f.familyMasks
    w.familyExtraMarginSize[:] = f.familyExtraMarginSize
    w.prescriptions[:] = f.prescriptions
    w.cd_update_freq = f.cd_update_freq
    return w
