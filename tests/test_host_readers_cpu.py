"""CPU-only: the facade's checkpoint readers (ReadClump*FromCsv, ReadContactPairsFromCsv,
ReadContactWildcardsFromCsv -- reference API.h:1124-1250) parse the files the facade / the reference write."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dem-engine_b200", "host")


def test_checkpoint_readers_round_trip(built, tmp_path):
    subprocess.run(["make", "-C", HOST], check=True, stdout=subprocess.DEVNULL)
    exe = str(tmp_path / "csv_readers_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(HOST, "include"), "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "host", "csv_readers_check.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "dem-engine_b200"), "-ldeme_b200", "-ldemcore",
                    "-Wl,-rpath," + os.path.join(ROOT, "dem-engine_b200")], check=True)
    rng = np.random.RandomState(5)
    n = 7
    types = ["three_sphere" if i % 3 else "ball" for i in range(n)]
    cols = rng.normal(size=(n, 13)).astype("f4")
    clumps = tmp_path / "clumps.csv"
    with open(clumps, "w") as f:
        # the column order of WriteClumpFile (XYZ, QUAT, clump_type, VEL, ANG_VEL, FAMILY)
        f.write("X,Y,Z,Qw,Qx,Qy,Qz,clump_type,v_x,v_y,v_z,w_x,w_y,w_z,family\n")
        for i in range(n):
            c = cols[i]
            f.write(",".join("%.9g" % v for v in c[:7]) + "," + types[i] + "," + ",".join("%.9g" % v for v in c[7:]) + ",0\n")
    contacts = tmp_path / "contacts.csv"
    rows = [("SS", 0, 2, 1, 7, rng.normal(size=7).astype("f4")), ("SA", 1, 9, 4, 0, rng.normal(size=7).astype("f4")),
            ("SS", 1, 3, 5, 11, rng.normal(size=7).astype("f4")), ("SM", 2, 10, 8, 3, rng.normal(size=7).astype("f4"))]
    with open(contacts, "w") as f:
        f.write("contact_type,A,B,geoA,geoB,f_x,f_y,f_z,delta_tan_x,delta_tan_y,delta_tan_z,delta_time\n")
        for t, a, b, ga, gb, v in rows:
            f.write("%s,%d,%d,%d,%d," % (t, a, b, ga, gb) + ",".join("%.9g" % x for x in v) + "\n")
    out = subprocess.run([exe, str(clumps), str(contacts)], capture_output=True, text=True, check=True).stdout.splitlines()
    got = {}
    for line in out:
        p = line.split()
        if p[0] == "clump":
            got.setdefault(p[1], []).append([float(x) for x in p[3:]])
    for name in ("three_sphere", "ball"):
        want = cols[[i for i in range(n) if types[i] == name]]
        assert np.allclose(np.array(got[name], "f4"), want, rtol=1e-6, atol=0), name
    assert "pairs 2 wildcards 4" in out and "sa_pairs 1" in out
    pr = [l.split() for l in out if l.startswith("pair ")]
    ss = [r for r in rows if r[0] == "SS"]
    for l, r in zip(pr, ss):
        assert (int(l[1]), int(l[2])) == (r[3], r[4])                       # geometry ids, not owner ids
        assert np.allclose([float(x) for x in l[3:]], r[5][3:], rtol=1e-6)   # the four history wildcards


def test_point_samplers(built, tmp_path):
    """Samplers the demo scripts generate their input with (reference src/DEM/utils/Samplers.hpp): Poisson-disk
    (minimum distance respected, maximal, reproducible), HCP / grid lattices, cylinder-surface shell."""
    from scipy.spatial import cKDTree
    exe = str(tmp_path / "samplers_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(HOST, "include"), "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "host", "samplers_check.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "dem-engine_b200"), "-ldeme_b200", "-ldemcore",
                    "-Wl,-rpath," + os.path.join(ROOT, "dem-engine_b200")], check=True)
    lines = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
    sets, i = {}, 0
    while i < len(lines):
        name, n = lines[i].split()
        n = int(n)
        sets[name] = np.array([[float(x) for x in l.split()] for l in lines[i + 1:i + 1 + n]], "f8").reshape(n, 3)
        i += 1 + n

    def min_dist(p):
        d, _ = cKDTree(p).query(p, k=2)
        return d[:, 1].min()

    pd = sets["pd_box"]
    c, h = np.array([0.1, -0.2, 0.3]), np.array([0.5, 0.4, 0.3])
    assert (np.abs(pd - c) <= h + 1e-6).all()
    assert min_dist(pd) >= 0.05 * (1 - 1e-5)
    # maximal: random probes inside the box all have a sample within the separation (a few per mille may sit in a gap
    # the dart throwing did not reach); density of a maximal Poisson-disk set is 0.65-0.9 points per separation^3
    probes = c + (np.random.RandomState(0).rand(4000, 3) * 2 - 1) * (h - 0.05)
    d, _ = cKDTree(pd).query(probes)
    assert (d < 0.05).mean() > 0.99
    dens = len(pd) * 0.05 ** 3 / np.prod(2 * h)
    assert 0.5 < dens < 1.0, dens
    assert np.array_equal(pd, sets["pd_box_again"])                       # seeded with 0: reproducible
    cyl = sets["pd_cylz"]
    assert (np.hypot(cyl[:, 0], cyl[:, 1]) <= 0.3 + 1e-6).all() and (np.abs(cyl[:, 2]) <= 0.2 + 1e-6).all()
    assert min_dist(cyl) >= 0.04 * (1 - 1e-5) and len(cyl) > 1000
    assert abs(min_dist(sets["hcp_box"]) - 0.05) < 1e-5 and abs(min_dist(sets["grid_box"]) - 0.05) < 1e-5
    assert len(sets["hcp_box"]) > len(sets["grid_box"])                    # HCP packs denser than the cubic grid
    shell = sets["cyl_surf"]
    assert np.allclose(np.hypot(shell[:, 0], shell[:, 1]), 0.5, atol=1e-5)
    assert shell[:, 2].min() >= 0.5 - 1e-5 and shell[:, 2].max() <= 1.5 + 1e-5 and len(shell) > 1000
    ball = sets["pd_sphere"]
    assert (np.linalg.norm(ball - np.array([1, 2, 3]), axis=1) <= 0.3 + 1e-6).all()
    assert min_dist(ball) >= 0.04 * (1 - 1e-5) and len(ball) > 800
    hb = sets["hcp_sphere"]
    assert (np.linalg.norm(hb, axis=1) <= 0.3 + 1e-6).all() and abs(min_dist(hb) - 0.05) < 1e-5
    assert len(hb) < len(sets["hcp_box"]) * 0.6                           # a ball fills 52 % of its bounding cube
    lay = sets["pd_layers"]
    zs = np.unique(np.round(lay[:, 2], 6))
    assert np.allclose(np.diff(zs), 0.04 * 1.05, atol=1e-6) and len(zs) == 5 and zs[0] == pytest.approx(0.4)
    assert (np.abs(lay[:, 0]) <= 0.3 + 1e-6).all() and (np.abs(lay[:, 1]) <= 0.2 + 1e-6).all()
    for z in zs:                                                           # Poisson-disk inside every layer
        assert min_dist(lay[np.abs(lay[:, 2] - z) < 1e-6]) >= 0.04 * 1.05 * (1 - 1e-5)
    assert np.array_equal(sets["hcp_box_vec"], sets["hcp_box"]) and np.array_equal(sets["grid_box_vec"], sets["grid_box"])
    tr = sets["truncated"][:, 0]
    assert tr.min() >= 0.8 and tr.max() <= 1.2 and tr.std() > 0.05


def test_prescription_expression_parser(tmp_path):
    """Prescription strings (expressions of t as the reference's demos write them) parsed and evaluated by the facade's
    TimeExpression, against Python's own evaluation of the same formulas."""
    import math
    exe = str(tmp_path / "expression_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(HOST, "include"),
                    os.path.join(ROOT, "tests", "host", "expression_check.cpp"), "-o", exe], check=True)
    cases = [  # (C expression, python expression or None for a syntax error, is constant)
        ("0.04 * sin(200 * t)", "0.04 * math.sin(200 * t)", False),
        ("-3.14 / 4", "-3.14 / 4", True),
        ("(t > 1.0) ? 2.0 * sin(5.0 * deme::PI * (t - 1.0)) : 0", "(2.0 * math.sin(5.0 * math.pi * (t - 1.0))) if t > 1.0 else 0.0", False),
        ("1.0f + pow(t, 2) - fmin(t, 0.5) * 3", "1.0 + t ** 2 - min(t, 0.5) * 3", False),
        ("2 - 3 - 4 * 2 / 8", "2 - 3 - 4 * 2 / 8", True),
        ("-(1 + 2) * -t", "-(1 + 2) * -t", False),
        ("t >= 0.5 && t < 1.5 ? 1 : 0", "1.0 if (t >= 0.5 and t < 1.5) else 0.0", False),
        ("!(t < 1) || t == 0.5", "1.0 if ((not (t < 1)) or t == 0.5) else 0.0", False),
        ("std::sqrt(fabs(t - 1)) + exp(-t) + atan2(t, 2) + erf(t)", "math.sqrt(abs(t - 1)) + math.exp(-t) + math.atan2(t, 2) + math.erf(t)", False),
        ("(float)3 / 2 + M_PI", "1.5 + math.pi", True),
        ("to_the_moon(t)", None, False),
        ("1 +", None, False),
        ("(1 + 2", None, False),
    ]
    src = tmp_path / "expr.txt"
    src.write_text("\n".join(c[0] for c in cases) + "\n")
    times = [0.0, 0.5, 1.3, 2.75]
    out = subprocess.run([exe, str(src)] + [repr(t) for t in times], capture_output=True, text=True, check=True).stdout.splitlines()
    assert len(out) == len(cases)
    for (cexpr, pyexpr, const), line in zip(cases, out):
        if pyexpr is None:
            assert line == "ERROR", (cexpr, line)
            continue
        p = line.split()
        assert int(p[0]) == int(const), cexpr
        for t, v in zip(times, p[1:]):
            want = float(eval(pyexpr, {"math": math, "t": t, "min": min, "abs": abs}))
            assert abs(float(v) - want) <= 1e-12 * max(1.0, abs(want)), (cexpr, t, v, want)


def test_wavefront_mesh_loading_and_transforms(built, tmp_path):
    """DEMMeshConnected on the host (reference BdrsAndObjs.h:222-520): OBJ loading incl. quads and v/vt/vn corners,
    Scale (mass ~ s^3, MOI ~ s^5), Move (rotate, then translate), Mirror (reflect and flip the winding)."""
    exe = str(tmp_path / "mesh_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(HOST, "include"), "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "host", "mesh_check.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "dem-engine_b200"), "-ldeme_b200", "-ldemcore",
                    "-Wl,-rpath," + os.path.join(ROOT, "dem-engine_b200")], check=True)
    obj = tmp_path / "shape.obj"
    obj.write_text("# a unit square (one quad) and a triangle above it\n"
                   "v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 0.5 1\n"
                   "vn 0 0 1\nvt 0 0\n"
                   "f 1/1/1 2/1/1 3/1/1 4/1/1\n"
                   "f 1//1 2//1 5//1\n"
                   "f -3 -2 -1\n")
    out = subprocess.run([exe, str(obj)], capture_output=True, text=True, check=True).stdout.splitlines()
    sets, cur = {}, None
    for l in out:
        p = l.split()
        if p[0] in ("loaded", "scaled", "moved", "mirrored"):
            cur = p[0]
            sets[cur] = dict(nv=int(p[1]), nt=int(p[2]), mass=float(p[4]), moi=[float(x) for x in p[6:9]], v=[], f=[])
        elif p[0] == "v":
            sets[cur]["v"].append([float(x) for x in p[1:]])
        elif p[0] == "f":
            sets[cur]["f"].append([int(x) for x in p[1:]])
    assert "missing 0" in out
    L = sets["loaded"]
    assert L["nv"] == 5 and L["nt"] == 4                       # the quad became two triangles
    V = np.array(L["v"])
    F = np.array(L["f"])
    assert F.min() == 0 and F.max() == 4                       # zero-based
    assert [2, 3, 4] in F.tolist()                             # negative indices count from the end
    tri_area = lambda v, f: 0.5 * np.linalg.norm(np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]]), axis=1)
    assert np.isclose(tri_area(V, F)[:2].sum(), 1.0)          # the two halves of the unit square
    S = sets["scaled"]
    assert np.allclose(np.array(S["v"]), 2 * V) and np.isclose(S["mass"], 2 * 8) and np.allclose(S["moi"], np.array([1, 2, 3]) * 32)
    M = np.array(sets["moved"]["v"])
    want = np.stack([-V[:, 1], V[:, 0], V[:, 2]], 1) + np.array([1, 2, 3])   # 90 degrees about z, then the shift
    assert np.allclose(M, want, atol=1e-6)
    R = sets["mirrored"]
    assert np.allclose(np.array(R["v"]), V * np.array([-1, 1, 1]))
    n0 = np.cross(V[F[:, 1]] - V[F[:, 0]], V[F[:, 2]] - V[F[:, 0]])
    Rv, Rf = np.array(R["v"]), np.array(R["f"])
    n1 = np.cross(Rv[Rf[:, 1]] - Rv[Rf[:, 0]], Rv[Rf[:, 2]] - Rv[Rf[:, 0]])
    assert np.allclose(n1, n0 * np.array([-1, 1, 1]))          # normals are reflected, not inverted


def test_facade_host_side_pieces(built, tmp_path):
    """Facade logic that never touches the device (region expressions of inspectors, the per-contact read-out container, the
    vector-form setters and wildcard bookkeeping of a clump batch, the force-model handle, frame / quaternion helpers):
    tests/host/facade_host_check.cpp asserts each against the behaviour the reference documents."""
    exe = str(tmp_path / "facade_host_check")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I" + os.path.join(HOST, "include"), "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "host", "facade_host_check.cpp"), "-o", exe,
                    "-L" + os.path.join(ROOT, "dem-engine_b200"), "-ldeme_b200", "-ldemcore",
                    "-Wl,-rpath," + os.path.join(ROOT, "dem-engine_b200")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.split("\n")[:7] == ["ok regions", "ok contact_info", "ok clump_batch", "ok force_model", "ok helpers",
                                        "ok mesh_and_templates", "ok input_files"]
