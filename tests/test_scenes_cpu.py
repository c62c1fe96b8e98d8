"""CPU-only: the synthetic inputs of BASELINE.json's configurations (pyapi/scenes.py) are what they claim to be -- a
wrong input would make every GPU parity / bench number meaningless."""
import math

import numpy as np

from pyapi import scenes


def test_drum_mesh_is_closed_and_faces_inwards():
    R, L = 0.3, 0.36
    v, f = scenes.drum_mesh(R, L, 48, 9, 8)
    v = v.astype("f8")
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    n = np.cross(b - a, c - a)
    area = 0.5 * np.linalg.norm(n, axis=1)
    assert (area > 1e-9).all()                                     # no degenerate facet
    # watertight: after welding coincident vertices (the caps carry their own rim ring), every edge is shared by exactly
    # two facets, once in each direction
    _, weld = np.unique(np.round(v, 6), axis=0, return_inverse=True)
    fw = weld.reshape(-1)[f]
    e = np.concatenate([fw[:, [0, 1]], fw[:, [1, 2]], fw[:, [2, 0]]])
    fwd = {}
    for p, q in e.tolist():
        fwd[(p, q)] = fwd.get((p, q), 0) + 1
    assert all(cnt == 1 for cnt in fwd.values())
    assert all((q, p) in fwd for (p, q) in fwd)
    # right-hand-rule normals point INTO the drum (towards the axis on the mantle, towards the middle on the caps)
    cen = (a + b + c) / 3.0
    inward = -cen.copy()
    assert ((n * inward).sum(1) > 0).all()
    # surface area of the tessellation approaches that of the cylinder from below
    exact = 2 * math.pi * R * L + 2 * math.pi * R * R
    assert 0.97 * exact < area.sum() <= exact * (1 + 1e-6)
    # signed volume (inward normals => negative) matches the cylinder's
    vol = (a * np.cross(b, c)).sum() / 6.0
    assert vol < 0
    assert 0.97 * math.pi * R * R * L < abs(vol) <= math.pi * R * R * L * (1 + 1e-6)


def test_config4_inputs():
    sc = scenes.config4_drum(20000, 3000, omega=3.0, spacing=2.9)
    assert len(sc.clump_type) == 20000 and set(np.unique(sc.clump_type).tolist()) == {0, 1, 2}
    nt = sum(len(m["faces"]) for m in sc.meshes)
    assert 0.8 * 3000 <= nt <= 1.25 * 3000
    # every clump starts inside the drum, clear of the mantle and the caps by its own circumscribed radius
    rr = np.hypot(sc.clump_xyz[:, 0], sc.clump_xyz[:, 2])
    reach = 1.46 * 0.004                                            # circumscribed radius of the largest template
    assert rr.max() + reach < sc.drum_radius and np.abs(sc.clump_xyz[:, 1]).max() + reach < sc.drum_length / 2
    q = sc.clump_quat.astype("f8")
    assert np.abs(np.sqrt((q * q).sum(1)) - 1).max() < 1e-6
    # no two clumps start closer than the lattice spacing
    from scipy.spatial import cKDTree
    d, _ = cKDTree(sc.clump_xyz).query(sc.clump_xyz, k=2)
    assert d[:, 1].min() > 2.9 * 0.004 * (1 - 1e-4)
    assert sc.prescribed[10]["angvel"] == (0.0, 3.0, 0.0)


def test_config5_inputs_and_partition():
    n = 200000
    full = scenes.config5_spheres(n)
    side = full.box[0] / 1.02
    assert len(full.clump_xyz) == n
    assert abs(n * 4.0 / 3.0 * math.pi * 1e-9 / side ** 3 - 0.5) < 1e-6            # 50 % packing
    assert np.abs(full.clump_xyz).max() <= side / 2 * (1 + 1e-6)
    # the per-GPU shares are a partition of the same cloud
    parts = [scenes.config5_spheres(n, x_range=(r / 4, (r + 1) / 4)) for r in range(4)]
    assert sum(len(p.clump_xyz) for p in parts) == n
    allp = np.concatenate([p.clump_xyz for p in parts])
    assert np.array_equal(np.sort(allp.view("f4,f4,f4"), axis=0), np.sort(full.clump_xyz.view("f4,f4,f4"), axis=0))
    for r, p in enumerate(parts):
        fr = p.clump_xyz[:, 0] / side + 0.5
        assert fr.min() >= r / 4 - 1e-6 and fr.max() < (r + 1) / 4 + 1e-6


def test_config2_inputs():
    sc = scenes.config2_clumps(10, 10, 10, scale=0.005, spacing=2.7)
    assert len(sc.clump_type) == 1000 and sc.bounding == "top_open"
    f = scenes.flatten(sc)
    assert f.nSpheres == 3000 and np.allclose(np.asarray(f.Radii), 0.004)
    # Mixer-demo mass properties of the 3-sphere clump at 5 mm (DEMdemo_Mixer.cpp:68-72)
    assert np.isclose(float(np.asarray(f.MassProperties).ravel()[0]), 2.6e3 * 5.5886717 * 0.005 ** 3, rtol=1e-6)
    assert np.allclose([f.moiX[0], f.moiY[0], f.moiZ[0]], np.array([2.928, 2.6029, 3.9908]) * 2.6e3 * 0.005 ** 5, rtol=1e-6)
    # random orientations from a fixed seed: reproducible, unit, and not all alike
    sc2 = scenes.config2_clumps(10, 10, 10, scale=0.005, spacing=2.7)
    assert np.array_equal(sc.clump_quat, sc2.clump_quat)
    assert np.abs(np.linalg.norm(sc.clump_quat.astype("f8"), axis=1) - 1).max() < 1e-6
    assert np.unique(np.round(sc.clump_quat, 3), axis=0).shape[0] > 900
