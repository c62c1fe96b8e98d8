"""CPU-only: properties of the host-side world sizing (dem_host_figure_out_nv / dem_host_box_domain /
dem_host_encode_positions, the counterparts of DEMSolver::figureOutNV APIPrivate.cpp:373-487,
InstructBoxDomainDimension APIPublic.cpp:845-872 and positionToVoxelID DEMHelperKernels.cuh:137-159), checked over
random boxes with hypothesis."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from pyapi import demb200

dims = st.floats(min_value=1e-3, max_value=1e3, allow_nan=False, allow_infinity=False)


@settings(max_examples=200, deadline=None)
@given(dims, dims, dims)
def test_figure_out_nv_properties(x, y, z):
    size = np.array([x, y, z], "f4")
    lo = -size / 2
    nx, ny, nz, l, vs = demb200.host_figure_out_nv(lo, lo + size)
    bits = np.array([nx, ny, nz])
    ext = (lo + size) - lo                                   # float subtraction, as the host routine sees it
    # the 64 bits of the voxel id are all handed out, and the world covers the box in every direction
    assert bits.sum() == 64 and (bits >= 1).all()
    assert vs == l * 65536.0
    world = vs * (2.0 ** bits)
    assert (world >= ext.astype("f8") * (1 - 1e-12)).all()
    # l is the smallest length unit that achieves that with these bits (one direction is tight)
    assert np.isclose((world / ext.astype("f8")).min(), 1.0, rtol=1e-9)
    # a longer side never gets fewer bits than a shorter one
    for i in range(3):
        for j in range(3):
            if ext[i] > ext[j]:
                assert bits[i] >= bits[j]
    # and the split is the reference's rule (APIPrivate.cpp:373-487): one extra bit per factor 2 between consecutive
    # sides (rounded at sqrt 2), the remaining bits shared out evenly, left-overs to the losers first
    s3 = np.sort(ext.astype("f4"), kind="stable").astype("f4")
    more = [0, 0]
    a, b, c = float(s3[0]), float(s3[1]), float(s3[2])
    while a < b and not (math.sqrt(2.0) * a > b):
        more[0] += 1
        a *= 2.0
    while b < c and not (math.sqrt(2.0) * b > c):
        more[1] += 1
        b *= 2.0
    total = 64 - 2 * more[0] - more[1]
    b3 = total // 3
    b2, left = b3 + more[0], total % 3
    b1 = b2 + more[1]
    while left > 0:
        if b3 < b2:
            b3 += 1
        elif b2 < b1:
            b2 += 1
        else:
            b1 += 1
        left -= 1
    assert sorted(bits.tolist()) == sorted([b3, b2, b1])
    l_ref = max(float(s3[0]) / 65536.0 / 2.0 ** b3, float(s3[1]) / 65536.0 / 2.0 ** b2, float(s3[2]) / 65536.0 / 2.0 ** b1)
    assert l == l_ref


@settings(max_examples=200, deadline=None)
@given(dims, dims, dims, st.integers(min_value=0, max_value=2))
def test_figure_out_nv_exact_direction(x, y, z, exact):
    """InstructBoxDomainDimension(..., dir_exact) (APIPrivate.cpp:442-476 of the reference): the world spans the box EXACTLY
    along the chosen axis, covers it along the other two, and all 64 bits are handed out."""
    size = np.array([x, y, z], "f4")
    lo = -size / 2
    ext = ((lo + size) - lo).astype("f8")
    nx, ny, nz, l, vs = demb200.host_figure_out_nv(lo, lo + size, exact_dir=exact)
    bits = np.array([nx, ny, nz], "i8")
    assert bits.sum() == 64 and vs == l * 65536.0
    world = vs * (2.0 ** bits.astype("f8"))
    assert world[exact] == ext[exact] or np.isclose(world[exact], ext[exact], rtol=1e-15)   # exact up to the power-of-two scaling
    assert (world >= ext * (1 - 1e-12)).all()
    # the reference's rule restated step by step (APIPrivate.cpp:378-476): rank the sides, split the bits, then let the
    # exact axis lend bits to the other two, the longer one first
    e32 = ((lo + size) - lo).astype("f4")
    order = [0, 1, 2]
    v = [float(e32[0]), float(e32[1]), float(e32[2])]
    for i in range(2):
        for j in range(i + 1, 3):
            if v[i] > v[j]:
                v[i], v[j] = v[j], v[i]
                order[i], order[j] = order[j], order[i]
    user = list(v)
    more = [0, 0]
    while v[0] < v[1] and not (math.sqrt(2.0) * v[0] > v[1]):
        more[0] += 1
        v[0] = float(np.float32(v[0] * 2.0))
    while v[1] < v[2] and not (math.sqrt(2.0) * v[1] > v[2]):
        more[1] += 1
        v[1] = float(np.float32(v[1] * 2.0))
    total = 64 - 2 * more[0] - more[1]
    b = [total // 3, 0, 0]
    b[1] = b[0] + more[0]
    b[2] = b[1] + more[1]
    for _ in range(total % 3):
        if b[0] < b[1]:
            b[0] += 1
        elif b[1] < b[2]:
            b[1] += 1
        else:
            b[2] += 1
    e = order.index(exact)
    others = {0: (1, 2), 1: (0, 2), 2: (0, 1)}[e]
    unit = lambda: user[e] / 65536.0 / 2.0 ** b[e]
    l_ref = unit()
    for k in (others[1], others[0]):
        while l_ref * 65536.0 * 2.0 ** b[k] < user[k]:
            b[e] -= 1
            b[k] += 1
            l_ref = unit()
    assert l == l_ref and [int(bits[order[p]]) for p in range(3)] == b
    # the axis lent no more bits than needed: with one bit back, one of the other two axes would fall short -- unless it
    # never lent any (then the split is the plain one of the rule above)
    nx0, ny0, nz0, _, _ = demb200.host_figure_out_nv(lo, lo + size)
    plain = np.array([nx0, ny0, nz0], "i8")
    lent = int(plain[exact] - bits[exact])
    assert lent >= 0
    if lent > 0:
        l_back = ext[exact] / 65536.0 / 2.0 ** (bits[exact] + 1)
        others = [k for k in range(3) if k != exact]
        short = [l_back * 65536.0 * 2.0 ** (bits[k] - 1) < ext[k] for k in others] + \
                [l_back * 65536.0 * 2.0 ** bits[k] < ext[k] for k in others]
        assert any(short)


@settings(max_examples=100, deadline=None)
@given(dims, dims, dims, st.integers(min_value=0, max_value=2 ** 31 - 1))
def test_position_code_round_trip(x, y, z, seed):
    """encode -> decode returns the position to within one length unit l, and never above it (truncating encode)."""
    umin, umax, tmin, tmax = demb200.host_box_domain(float(np.float32(x)), float(np.float32(y)), float(np.float32(z)))
    # the target box is the user box enlarged by 20 % (10 % on each side)
    assert np.allclose(tmax - tmin, (umax - umin) * 1.2, rtol=1e-5)
    nx, ny, nz, l, vs = demb200.host_figure_out_nv(tmin, tmax)
    p = demb200.DemSimParams()
    p.nvXp2, p.nvYp2, p.nvZp2, p.l, p.voxelSize = nx, ny, nz, l, vs
    for k in range(3):
        p.LBF[k] = float(tmin[k])
    rng = np.random.RandomState(seed)
    n = 64
    pts = (umin + rng.rand(n, 3).astype("f4") * (umax - umin)).astype("f4")
    vox, lx, ly, lz = np.zeros(n, "u8"), np.zeros(n, "u2"), np.zeros(n, "u2"), np.zeros(n, "u2")
    import ctypes as C
    lib = demb200.load_library()
    assert lib.dem_host_encode_positions(C.byref(p), pts.ctypes.data_as(C.c_void_p), C.c_uint64(n),
                                         vox.ctypes.data_as(C.c_void_p), lx.ctypes.data_as(C.c_void_p),
                                         ly.ctypes.data_as(C.c_void_p), lz.ctypes.data_as(C.c_void_p)) == 0
    vxs = [vox & np.uint64((1 << nx) - 1), (vox >> np.uint64(nx)) & np.uint64((1 << ny) - 1), vox >> np.uint64(nx + ny)]
    for k, (v, sub) in enumerate(zip(vxs, (lx, ly, lz))):
        back = v.astype("f8") * vs + sub.astype("f8") * l
        want = (pts[:, k] - np.float32(tmin[k])).astype("f8")   # float subtraction of LBF, as dT.cpp does
        assert (back <= want + 1e-9 * abs(vs)).all()
        assert (want - back < l * (1 + 1e-6) + 1e-12 * np.abs(want)).all()
