"""ctypes binding of libdemcore.so (the C ABI of include/dem_b200.h).

Thin host-side plumbing used by tests/, bench.py and __graft_entry__.py: every call goes straight through the
C ABI into the hand-written sm_100a kernels.  There is NO CPU fallback: if the shared library is missing or no GPU is
present, construction raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(os.path.dirname(_HERE), "libdemcore.so")

DEM_OK, DEM_ERR_INVALID, DEM_ERR_CUDA, DEM_ERR_CAPACITY, DEM_ERR_VELOCITY, DEM_ERR_NO_GPU = 0, -1, -2, -3, -4, -5
FORWARD_EULER, CENTERED_DIFFERENCE, EXTENDED_TAYLOR = 0, 1, 2
HERTZIAN, HERTZIAN_FRICTIONLESS = 0, 1
REDUCE_MAX_ABSV, REDUCE_MAX_Z, REDUCE_MIN_Z, REDUCE_KINETIC_ENERGY, REDUCE_TOTAL_MASS = 0, 1, 2, 3, 4

EXPORTED_SYMBOLS = [
    "dem_abi_version", "dem_host_figure_out_nv", "dem_host_box_domain", "dem_host_encode_positions",
    "dem_ctx_create", "dem_ctx_destroy", "dem_last_error", "dem_set_stream", "dem_set_params",
    "dem_upload_templates", "dem_upload_materials", "dem_upload_analytical", "dem_upload_families",
    "dem_upload_owners", "dem_upload_spheres", "dem_upload_triangles", "dem_update_triangle_nodes", "dem_host_partition_owners", "dem_debug_download", "dem_profile_binning", "dem_initialize", "dem_set_contacts",
    "dem_do_dynamics", "dem_step", "dem_step_async", "dem_sync", "dem_rebuild_contacts", "dem_update_step_size",
    "dem_download_owner_state", "dem_download_positions", "dem_upload_owner_state", "dem_download_contacts",
    "dem_get_stats", "dem_set_sim_time", "dem_download_contact_records", "dem_reduce", "dem_reduce_many", "dem_profile_steps", "dem_profile_rebuild", "dem_set_option",
    "dem_mgpu_unique_id", "dem_mgpu_init", "dem_mgpu_info", "dem_host_slab_bounds",
    "dem_mgpu_init_local", "dem_mgpu_barrier", "dem_device_count", "dem_ctx_create_group", "dem_set_family_material", "dem_group_step_async", "dem_group_sync", "dem_group_gather", "dem_add_owner_acc", "dem_host_figure_out_nv_exact",
]


class DemSimParams(C.Structure):
    _fields_ = [
        ("nvXp2", C.c_uint32), ("nvYp2", C.c_uint32), ("nvZp2", C.c_uint32), ("integrator", C.c_uint32),
        ("force_model", C.c_uint32), ("cd_update_freq", C.c_uint32),
        ("l", C.c_double), ("voxelSize", C.c_double),
        ("LBF", C.c_float * 3), ("G", C.c_float * 3), ("userBoxMin", C.c_float * 3), ("userBoxMax", C.c_float * 3),
        ("h", C.c_float), ("beta", C.c_float), ("approxMaxVel", C.c_float), ("expSafetyMulti", C.c_float),
        ("expSafetyAdder", C.c_float), ("errOutVel", C.c_float),
        ("record_contact_forces", C.c_uint32), ("pad_", C.c_uint32),
    ]


class DemStats(C.Structure):
    _fields_ = [
        ("n_steps", C.c_uint64), ("n_rebuilds", C.c_uint64), ("n_contacts_ss", C.c_uint64),
        ("n_contacts_sa", C.c_uint64), ("n_contacts_st", C.c_uint64), ("contact_capacity", C.c_uint64),
        ("kernel_launches", C.c_uint64), ("device_bytes", C.c_uint64), ("sim_time", C.c_double),
        ("max_margin", C.c_float), ("cell_size", C.c_float), ("n_cells", C.c_uint32 * 3), ("overflow", C.c_uint32),
        ("cd_update_freq", C.c_uint32), ("n_contacts_ss_touching", C.c_uint64),
    ]


PRESC_DTYPE = np.dtype([
    ("used", "u1"), ("linVelPrescribed", "u1", 3), ("rotVelPrescribed", "u1", 3), ("linPosPrescribed", "u1", 3),
    ("rotPosPrescribed", "u1"), ("hasLinVel", "u1", 3), ("hasRotVel", "u1", 3), ("hasLinPos", "u1", 3),
    ("hasAcc", "u1", 3), ("hasAngAcc", "u1", 3), ("pad_", "u1", 2), ("linVel", "f4", 3), ("rotVel", "f4", 3),
    ("linPos", "f4", 3), ("acc", "f4", 3), ("angAcc", "f4", 3)])
assert PRESC_DTYPE.itemsize == 88

_lib = None


def load_library(path=None):
    """dlopen libdemcore.so. Loading needs no GPU; creating a context does."""
    global _lib
    if _lib is None:
        p = path or _LIB_PATH
        if not os.path.exists(p):
            raise RuntimeError("libdemcore.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'`"
                               " -- there is no CPU fallback" % p)
        _lib = C.CDLL(p)
        _lib.dem_last_error.restype = C.c_char_p
        _lib.dem_last_error.argtypes = [C.c_void_p]
        _lib.dem_do_dynamics.argtypes = [C.c_void_p, C.c_double]
        _lib.dem_set_sim_time.argtypes = [C.c_void_p, C.c_double]
        _lib.dem_host_box_domain.argtypes = [C.c_float, C.c_float, C.c_float] + [C.c_void_p] * 4
        for name in ("dem_step", "dem_step_async"):
            getattr(_lib, name).argtypes = [C.c_void_p, C.c_uint64]
        _lib.dem_initialize.argtypes = [C.c_void_p, C.c_uint64]
        _lib.dem_update_step_size.argtypes = [C.c_void_p, C.c_float]
        _lib.dem_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def host_figure_out_nv(box_min, box_max, exact_dir=None):
    """figureOutNV; exact_dir = 0 / 1 / 2 makes the world span the box exactly along X / Y / Z."""
    lib = load_library()
    mn = np.ascontiguousarray(box_min, "f4")
    mx = np.ascontiguousarray(box_max, "f4")
    nv = np.zeros(3, "u4")
    l, vs = C.c_double(), C.c_double()
    if exact_dir is None:
        rc = lib.dem_host_figure_out_nv(_p(mn), _p(mx), _p(nv), C.byref(l), C.byref(vs))
    else:
        rc = lib.dem_host_figure_out_nv_exact(_p(mn), _p(mx), C.c_int(exact_dir), _p(nv), C.byref(l), C.byref(vs))
    assert rc == 0
    return int(nv[0]), int(nv[1]), int(nv[2]), l.value, vs.value


def host_box_domain(x, y, z):
    lib = load_library()
    out = [np.zeros(3, "f4") for _ in range(4)]
    rc = lib.dem_host_box_domain(x, y, z, *[_p(o) for o in out])
    assert rc == 0
    return out


def params_from_flat(f):
    """DemSimParams of a flattened scene (scenes.flatten)."""
    p = DemSimParams()
    p.nvXp2, p.nvYp2, p.nvZp2 = f.nvXp2, f.nvYp2, f.nvZp2
    p.integrator, p.force_model, p.cd_update_freq = f.integrator, f.force_model, f.cd_update_freq
    p.l, p.voxelSize = f.l, f.voxelSize
    for k in range(3):
        p.LBF[k], p.G[k] = float(f.LBF[k]), float(f.G[k])
        p.userBoxMin[k], p.userBoxMax[k] = float(f.userBoxMin[k]), float(f.userBoxMax[k])
    p.h, p.beta = float(f.h), float(f.beta)
    p.approxMaxVel, p.expSafetyMulti, p.expSafetyAdder = float(f.approxMaxVel), float(f.expSafetyMulti), float(f.expSafetyAdder)
    p.errOutVel = float(getattr(f, "errOutVel", 1e3))
    p.record_contact_forces = int(getattr(f, "record_contact_forces", 0))
    return p


def host_slab_bounds(p, world, rank):
    """x-slab (LBF-relative) of `rank`: dem_host_slab_bounds."""
    lo, hi = C.c_float(), C.c_float()
    rc = load_library().dem_host_slab_bounds(C.byref(p), int(world), int(rank), C.byref(lo), C.byref(hi))
    assert rc == 0
    return lo.value, hi.value


def host_partition_owners(p, world, rank, halo, voxelID, locX):
    """(role, send) per owner on `rank`: dem_host_partition_owners."""
    voxelID = np.ascontiguousarray(voxelID, "u8")
    locX = np.ascontiguousarray(locX, "u2")
    n = len(voxelID)
    role, send = np.zeros(n, "u1"), np.zeros(n, "u1")
    lib = load_library()
    lib.dem_host_partition_owners.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_uint64] + [C.c_void_p] * 4
    rc = lib.dem_host_partition_owners(C.byref(p), int(world), int(rank), float(halo), n, _p(voxelID), _p(locX), _p(role), _p(send))
    assert rc == 0
    return role, send


class DemError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dem_b200 error %d: %s" % (code, msg))
        self.code = code


class Engine:
    """One DemCtx. Mirrors the call order of DEMSolver::Initialize / DoDynamics of the reference."""

    def __init__(self, device=0, devices=None):
        """device: one GPU; devices=[...]: ONE context driving a group of GPUs (dem_ctx_create_group)"""
        self.lib = load_library()
        self.ctx = C.c_void_p()
        if devices is not None:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            rc = self.lib.dem_ctx_create_group(C.byref(self.ctx), arr, len(devices))
        else:
            rc = self.lib.dem_ctx_create(C.byref(self.ctx), int(device))
        if rc != 0:
            raise DemError(rc, "dem_ctx_create failed (no CUDA device? there is no CPU fallback)")
        self.params = None

    def close(self):
        if self.ctx:
            self.lib.dem_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise DemError(rc, self.lib.dem_last_error(self.ctx).decode())

    # ---- set-up ----
    def set_params(self, p: DemSimParams):
        self.params = p
        self._ck(self.lib.dem_set_params(self.ctx, C.byref(p)))

    def set_stream(self, cuda_stream_ptr):
        self._ck(self.lib.dem_set_stream(self.ctx, C.c_void_p(cuda_stream_ptr)))

    def load_flat(self, f, contact_capacity=0):
        """f: FlatWorld-like object (see scenes.flatten) with the reference's SoA arrays."""
        p = params_from_flat(f)
        self.set_params(p)
        lib = self.lib
        self._ck(lib.dem_upload_templates(self.ctx, C.c_uint32(f.nComp), _p(f.Radii), _p(f.CDRelPosX), _p(f.CDRelPosY),
                                          _p(f.CDRelPosZ), C.c_uint32(f.nMassProps), _p(f.MassProperties), _p(f.moiX),
                                          _p(f.moiY), _p(f.moiZ)))
        self._ck(lib.dem_upload_materials(self.ctx, C.c_uint32(f.nMat), _p(f.E), _p(f.nu), _p(f.CoR), _p(f.mu), _p(f.Crr)))
        self._ck(lib.dem_upload_analytical(self.ctx, C.c_uint32(f.nAnal), _p(f.objOwner), _p(f.objType), _p(f.objMaterial),
                                           _p(f.objNormal), _p(f.objRelPosX), _p(f.objRelPosY), _p(f.objRelPosZ),
                                           _p(f.objRotX), _p(f.objRotY), _p(f.objRotZ), _p(f.objSize1), _p(f.objSize2),
                                           _p(f.objSize3), _p(f.objMass)))
        presc = np.ascontiguousarray(f.prescriptions)
        assert presc.dtype.itemsize == 88
        self._ck(lib.dem_upload_families(self.ctx, _p(f.familyMasks), _p(f.familyExtraMarginSize), _p(presc)))
        self._ck(lib.dem_upload_owners(self.ctx, C.c_uint32(f.nOwners), _p(f.voxelID), _p(f.locX), _p(f.locY), _p(f.locZ),
                                       _p(f.oriQw), _p(f.oriQx), _p(f.oriQy), _p(f.oriQz), _p(f.vX), _p(f.vY), _p(f.vZ),
                                       _p(f.omgBarX), _p(f.omgBarY), _p(f.omgBarZ), _p(f.familyID),
                                       _p(f.inertiaPropOffsets)))
        self._ck(lib.dem_upload_spheres(self.ctx, C.c_uint32(f.nSpheres), _p(f.ownerClumpBody), _p(f.clumpComponentOffset),
                                        _p(f.sphereMaterialOffset)))
        nTri = int(getattr(f, "nTri", 0))
        if nTri:
            self._ck(lib.dem_upload_triangles(self.ctx, C.c_uint32(nTri), _p(f.ownerMesh), _p(f.relPosNode1), _p(f.relPosNode2),
                                              _p(f.relPosNode3), _p(f.triMaterialOffset)))
        else:
            self._ck(lib.dem_upload_triangles(self.ctx, C.c_uint32(0), None, None, None, None, None))
        self._ck(lib.dem_initialize(self.ctx, int(contact_capacity)))
        self.nOwners, self.nSpheres = int(f.nOwners), int(f.nSpheres)

    def set_contacts(self, idA, idB, ctype, wildcards4=None):
        idA = np.ascontiguousarray(idA, "u4"); idB = np.ascontiguousarray(idB, "u4")
        ctype = np.ascontiguousarray(ctype, "u1")
        wc = None if wildcards4 is None else np.ascontiguousarray(wildcards4, "f4")
        self._ck(self.lib.dem_set_contacts(self.ctx, C.c_uint64(len(idA)), _p(idA), _p(idB), _p(ctype), _p(wc)))

    def update_triangle_nodes(self, first, n1, n2, n3):
        """New owner-frame node positions (n x 3 each) of the facets starting at `first` (deforming mesh)."""
        n1, n2, n3 = (np.ascontiguousarray(a, "f4").reshape(-1, 3) for a in (n1, n2, n3))
        self._ck(self.lib.dem_update_triangle_nodes(self.ctx, C.c_uint32(first), C.c_uint32(len(n1)), _p(n1), _p(n2), _p(n3)))

    def update_families(self, masks, extra, presc):
        presc = np.ascontiguousarray(presc)
        self._ck(self.lib.dem_upload_families(self.ctx, _p(masks), _p(extra), _p(presc)))

    # ---- hot loop ----
    def step(self, n=1):
        self._ck(self.lib.dem_step(self.ctx, int(n)))

    def step_async(self, n=1):
        self._ck(self.lib.dem_step_async(self.ctx, int(n)))

    def sync(self):
        self._ck(self.lib.dem_sync(self.ctx))

    def do_dynamics(self, t):
        self._ck(self.lib.dem_do_dynamics(self.ctx, float(t)))

    def rebuild_contacts(self):
        self._ck(self.lib.dem_rebuild_contacts(self.ctx))

    def update_step_size(self, h):
        self._ck(self.lib.dem_update_step_size(self.ctx, float(h)))

    # ---- state ----
    def owner_state(self, first=0, n=None):
        n = self.nOwners - first if n is None else n
        out = {"voxelID": np.zeros(n, "u8"), "locX": np.zeros(n, "u2"), "locY": np.zeros(n, "u2"),
               "locZ": np.zeros(n, "u2"), "oriQ": np.zeros((n, 4), "f4"), "vel": np.zeros((n, 3), "f4"),
               "omg": np.zeros((n, 3), "f4"), "acc": np.zeros((n, 3), "f4"), "angacc": np.zeros((n, 3), "f4"),
               "family": np.zeros(n, "u1")}
        self._ck(self.lib.dem_download_owner_state(self.ctx, C.c_uint32(first), C.c_uint32(n), _p(out["voxelID"]),
                                                   _p(out["locX"]), _p(out["locY"]), _p(out["locZ"]), _p(out["oriQ"]),
                                                   _p(out["vel"]), _p(out["omg"]), _p(out["acc"]), _p(out["angacc"]),
                                                   _p(out["family"])))
        return out

    def positions(self, first=0, n=None, f64=True):
        n = self.nOwners - first if n is None else n
        x32 = np.zeros((n, 3), "f4")
        x64 = np.zeros((n, 3), "f8")
        self._ck(self.lib.dem_download_positions(self.ctx, C.c_uint32(first), C.c_uint32(n), _p(x32), _p(x64)))
        return x64 if f64 else x32

    def upload_owner_state(self, first, pos=None, oriQ=None, vel=None, omg=None, family=None):
        arrs = [None if a is None else np.ascontiguousarray(a, dt) for a, dt in
                ((pos, "f4"), (oriQ, "f4"), (vel, "f4"), (omg, "f4"), (family, "u1"))]
        n = next(len(a.reshape(-1, w)) for a, w in zip(arrs, (3, 4, 3, 3, 1)) if a is not None)
        self._ck(self.lib.dem_upload_owner_state(self.ctx, C.c_uint32(first), C.c_uint32(n), *[_p(a) for a in arrs]))

    def add_owner_acc(self, first, acc=None, angacc_local=None):
        """AddOwnerNextStepAcc / AddOwnerNextStepAngAcc: extra (angular) acceleration of consecutive owners, next step only."""
        arrs = [None if a is None else np.ascontiguousarray(a, "f4").reshape(-1, 3) for a in (acc, angacc_local)]
        n = next(len(a) for a in arrs if a is not None)
        self._ck(self.lib.dem_add_owner_acc(self.ctx, C.c_uint32(first), C.c_uint32(n), *[_p(a) for a in arrs]))

    def contacts(self, with_force=False):
        n = C.c_uint64(0)
        self._ck(self.lib.dem_download_contacts(self.ctx, C.c_uint64(0), C.byref(n), None, None, None, None, None))
        m = int(n.value)
        idA, idB, ct = np.zeros(m, "u4"), np.zeros(m, "u4"), np.zeros(m, "u1")
        wc, fr = np.zeros((m, 4), "f4"), np.zeros((m, 3), "f4")
        if m:
            self._ck(self.lib.dem_download_contacts(self.ctx, C.c_uint64(m), C.byref(n), _p(idA), _p(idB), _p(ct), _p(wc),
                                                    _p(fr) if with_force else None))
        return (idA, idB, ct, wc, fr) if with_force else (idA, idB, ct, wc)

    def contact_records(self):
        """(idA, idB, type, wildcards, force on A, contact point in the world frame) of every listed contact."""
        n = C.c_uint64(0)
        self._ck(self.lib.dem_download_contacts(self.ctx, C.c_uint64(0), C.byref(n), None, None, None, None, None))
        m = int(n.value)
        idA, idB, ct = np.zeros(m, "u4"), np.zeros(m, "u4"), np.zeros(m, "u1")
        wc, fr, pt = np.zeros((m, 4), "f4"), np.zeros((m, 3), "f4"), np.zeros((m, 3), "f4")
        if m:
            self._ck(self.lib.dem_download_contact_records(self.ctx, C.c_uint64(m), C.byref(n), _p(idA), _p(idB), _p(ct),
                                                           _p(wc), _p(fr), _p(pt)))
        return idA, idB, ct, wc, fr, pt

    def stats(self):
        s = DemStats()
        self._ck(self.lib.dem_get_stats(self.ctx, C.byref(s)))
        return s

    def reduce(self, kind):
        out = C.c_double(0)
        self._ck(self.lib.dem_reduce(self.ctx, int(kind), C.byref(out)))
        return out.value

    def reduce_many(self, kinds):
        """Several reductions in one pass and one read-back (dem_reduce_many); returns {kind: value}."""
        mask = 0
        for k in kinds:
            mask |= 1 << int(k)
        out = (C.c_double * 5)()
        self._ck(self.lib.dem_reduce_many(self.ctx, C.c_uint32(mask), out))
        return {int(k): out[int(k)] for k in kinds}

    def profile_binning(self, repeats=10):
        out = (C.c_float * 3)()
        self._ck(self.lib.dem_profile_binning(self.ctx, int(repeats), out))
        return {"keys_us": out[0], "sort_us": out[1], "total_us": out[2]}

    def debug_download(self, what, n=None):
        """Raw device scratch of the last rebuild / step (dem_debug_download)."""
        dt = "f4" if what == "sphere_pos" else "u4"
        if n is None:
            n = 4 * self.nSpheres if what == "sphere_pos" else self.nSpheres
        out = np.zeros(max(int(n), 1), dt)
        got = C.c_uint64()
        self.lib.dem_debug_download.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_uint64, C.c_void_p]
        self._ck(self.lib.dem_debug_download(self.ctx, what.encode(), _p(out), int(n), C.byref(got)))
        return out[: got.value]

    # ---- multi-GPU ----
    @staticmethod
    def mgpu_unique_id():
        out = np.zeros(128, "u1")
        rc = load_library().dem_mgpu_unique_id(_p(out))
        if rc != 0:
            raise DemError(rc, "dem_mgpu_unique_id failed (NCCL not loadable?)")
        return out

    def mgpu_init(self, rank, world, unique_id):
        uid = np.ascontiguousarray(unique_id, "u1")
        self._ck(self.lib.dem_mgpu_init(self.ctx, int(rank), int(world), _p(uid)))

    def mgpu_barrier(self):
        self._ck(self.lib.dem_mgpu_barrier(self.ctx))

    def mgpu_info(self):
        out = np.zeros(6, "u8")
        self._ck(self.lib.dem_mgpu_info(self.ctx, _p(out)))
        d = dict(zip(["n_own", "n_active", "n_send_left", "n_send_right", "halo_bytes_per_step", "world"], out.tolist()))
        d["peer_memory_exchange"] = bool(d["world"] >> 32)
        d["world"] &= 0xffffffff
        return d

    def set_option(self, name, value):
        self._ck(self.lib.dem_set_option(self.ctx, name.encode(), float(value)))

    def profile_rebuild(self):
        out = (C.c_float * 8)()
        self._ck(self.lib.dem_profile_rebuild(self.ctx, out))
        names = ["prep_us", "sort_us", "cellscan_us", "gather_us", "sweep_us", "counts_us", "redistribute_us", "total_us"]
        return dict(zip(names, [float(x) for x in out]))

    def profile_steps(self, n):
        out = (C.c_float * 8)()
        self._ck(self.lib.dem_profile_steps(self.ctx, C.c_uint64(n), out))
        return {"force_ss_us": out[0], "force_sa_us": out[1], "integrate_us": out[2], "rebuild_us_per_step": out[3],
                "step_us": out[4], "halo_exchange_us": out[5]}


class EngineGroup:
    """Slab decomposition over several GPUs of THIS process (dem_mgpu_init_local): one Engine per device, all loaded with
    the same flattened scene, stepped together (dem_group_step_async / dem_group_sync)."""

    def __init__(self, flat, devices, contact_capacity=0, options=None):
        self.lib = load_library()
        self.engines = [Engine(d) for d in devices]
        for e in self.engines:
            for k, v in (options or {}).items():
                e.set_option(k, v)
            e.load_flat(flat, contact_capacity)
        self.world = len(self.engines)
        self._arr = (C.c_void_p * self.world)(*[e.ctx for e in self.engines])
        self._ck(self.lib.dem_mgpu_init_local(self._arr, self.world))

    def _ck(self, rc):
        if rc != 0:
            msgs = [self.lib.dem_last_error(e.ctx).decode() for e in self.engines]
            raise DemError(rc, " | ".join(m for m in msgs if m))

    def step_async(self, n):
        self.lib.dem_group_step_async.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        self._ck(self.lib.dem_group_step_async(self._arr, self.world, int(n)))

    def sync(self):
        self._ck(self.lib.dem_group_sync(self._arr, self.world))

    def step(self, n):
        self.step_async(n)
        self.sync()

    def gather(self):
        """merge the owners of all ranks into engines[0] (positions(), owner_state() of engines[0] then see everything)"""
        self._ck(self.lib.dem_group_gather(self._arr, self.world))
        return self.engines[0]

    def close(self):
        for e in self.engines:
            e.close()
