"""CPU-only: the deme::DEMSolver facade driven over a RECORDING TEST DOUBLE of the device-touching C ABI
(tests/host/fake_core.cpp: it keeps what the facade uploads, answers read-backs from that and from tables the test injects,
and computes no physics -- it is not a CPU path of the product, and nothing outside this test links it).  What runs is the
facade's own host logic, under AddressSanitizer / UBSan: the flattening Initialize() does (owner / sphere / facet / family
tables as the reference's dT::populateEntityArrays lays them out, dT.cpp:638-1024), the step-count rule of DoDynamics, the
clump / sphere / contact / mesh file writers with the reference's columns (dT.cpp:1254-1936), the detailed contact read-out
(normals, owners, families), contact-wildcard edits, persistent-contact marks, region inspectors, trackers, material
pair tables, analytical components and the bounding box, restart contacts, UpdateClumps (state and contact list carried over,
owners renumbered)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dem-engine_b200", "host")
LIBDIR = os.path.join(ROOT, "dem-engine_b200")


def test_facade_host_logic_over_the_recording_fake(built, tmp_path):
    fake = str(tmp_path / "libfake_demcore.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", os.path.join(ROOT, "tests", "host", "fake_core.cpp"), "-o", fake,
                    "-L" + LIBDIR, "-ldemcore", "-Wl,-rpath," + LIBDIR], check=True)
    exe = str(tmp_path / "facade_fake_check")
    # the facade's source is compiled INTO the test (sanitised) instead of taking libdeme_b200.so; the fake comes before the
    # real core in the link order, so its definitions of the device-touching entry points win and dem_host_* stay real
    subprocess.run(["g++", "-O0", "-std=c++17", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                    "-I" + os.path.join(HOST, "include"), "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "host", "facade_fake_check.cpp"), os.path.join(HOST, "src", "API.cpp"), "-o", exe,
                    "-Wl,--no-as-needed", "-L" + str(tmp_path), "-lfake_demcore", "-L" + LIBDIR, "-ldemcore",
                    "-Wl,-rpath," + str(tmp_path), "-Wl,-rpath," + LIBDIR], check=True)
    r = subprocess.run([exe, str(tmp_path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr[-3000:]
    assert r.stdout.split() == ["ok", "initialize", "ok", "stepping", "ok", "contact_readout", "ok", "files", "ok",
                                "wildcards_persistence", "ok", "inspectors", "ok", "controls", "ok", "second_solver"]
    assert "ERROR: AddressSanitizer" not in r.stderr and "runtime error" not in r.stderr, r.stderr[-3000:]
    # the contact file as a post-processing script would read it
    with open(tmp_path / "fake_contacts.csv") as fh:
        assert fh.readline().strip() == ("contact_type,A,B,geoA,geoB,f_x,f_y,f_z,X,Y,Z,n_x,n_y,n_z,"
                                         "delta_tan_x,delta_tan_y,delta_tan_z,delta_time")


def test_fake_core_is_test_infrastructure_only():
    """The double lives under tests/ and the product never refers to it."""
    for base, _, files in os.walk(os.path.join(ROOT, "dem-engine_b200")):
        for fn in files:
            if fn.endswith((".cpp", ".cu", ".cuh", ".h", ".hpp", ".py", "Makefile")):
                with open(os.path.join(base, fn), errors="replace") as fh:
                    assert "fake_core" not in fh.read(), os.path.join(base, fn)
    for fn in ("bench.py", "__graft_entry__.py"):
        with open(os.path.join(ROOT, fn)) as fh:
            assert "fake_core" not in fh.read(), fn


REF = "/root/reference"
QUICK_DEMOS = ["BallDrop", "Repose", "TestPack", "ContactChain", "WheelDPSimplified", "Plow", "FlexibleMesh"]


def _cuda_present():
    from conftest import _cuda_device_count
    return _cuda_device_count() > 0


import pytest  # noqa: E402


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src", "demo")), reason="the reference tree is not present on this machine")
@pytest.mark.parametrize("name", QUICK_DEMOS)
def test_reference_demo_host_side_over_the_recording_fake(built, name, tmp_path):
    """The reference's own demo script (compiled unmodified, `make refdemos`) started with the recording fake preloaded in front
    of the real core: no physics happens, but everything the script does on the HOST side does -- samplers, template and mesh
    loading from the reference's data files (Windows line endings, files without a radius column), flattening at Initialize(),
    trackers, inspectors, family changes, the clump / mesh / contact files it writes every frame -- and it must reach its
    "exiting" line without an exception.  (This is how two reader bugs were found.)"""
    exe = os.path.join(HOST, "refdemo", "DEMdemo_" + name)
    if not os.path.exists(exe):
        pytest.skip("refdemo/DEMdemo_%s not built" % name)
    fake = str(tmp_path / "libfake_demcore.so")
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", os.path.join(ROOT, "tests", "host", "fake_core.cpp"), "-o", fake,
                    "-L" + LIBDIR, "-ldemcore", "-Wl,-rpath," + LIBDIR], check=True)
    # some scripts address their input as ../data/... (they expect to run from <build>/bin)
    work = tmp_path / "bin"
    work.mkdir()
    os.symlink(os.path.join(REF, "data"), str(tmp_path / "data"))
    env = dict(os.environ, LD_PRELOAD=fake, DEME_DATA_PATH=os.path.join(REF, "data"))
    r = subprocess.run(["timeout", "120", exe], cwd=str(work), env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       errors="replace")
    tail = r.stdout[-1500:]
    assert r.returncode == 0, (name, r.returncode, tail)
    assert "exiting" in tail and "what():" not in r.stdout and "terminate called" not in r.stdout, tail
