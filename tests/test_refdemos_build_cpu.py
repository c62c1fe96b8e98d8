"""The reference's own demo scripts against the facade, the part that needs no GPU: every one of the 27
/root/reference/src/demo/DEMdemo_*.cpp compiles and links UNMODIFIED (`make refdemos`), and a binary started on a machine
without a CUDA device stops with the core's own error -- there is no CPU path it could silently take."""
import os
import subprocess

import pytest

from conftest import _cuda_device_count

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "dem-engine_b200", "host")
REF_DEMOS = "/root/reference/src/demo"


def _reference_demo_names():
    return sorted(f[len("DEMdemo_"):-len(".cpp")] for f in os.listdir(REF_DEMOS)
                  if f.startswith("DEMdemo_") and f.endswith(".cpp"))


@pytest.mark.skipif(not os.path.isdir(REF_DEMOS), reason="the reference tree is not present on this machine")
def test_every_reference_demo_compiles_and_links_unmodified(built):
    names = _reference_demo_names()
    assert len(names) >= 27
    listed = subprocess.run(["make", "-s", "-C", HOST, "--eval", "show: ; @echo $(REFDEMOS)", "show"],
                            stdout=subprocess.PIPE, text=True, check=True).stdout.split()
    assert sorted(listed) == names, "host/Makefile's REFDEMOS must list every demo of the reference"
    for n in names:
        exe = os.path.join(HOST, "refdemo", "DEMdemo_" + n)
        assert os.path.exists(exe), exe
        # the binary is newer than the script it was compiled from, and it resolves its libraries
        assert os.path.getmtime(exe) >= os.path.getmtime(os.path.join(REF_DEMOS, "DEMdemo_%s.cpp" % n))
        ldd = subprocess.run(["ldd", exe], stdout=subprocess.PIPE, text=True).stdout
        assert "libdeme_b200.so" in ldd and "libdemcore.so" in ldd and "not found" not in ldd, ldd


@pytest.mark.skipif(_cuda_device_count() > 0, reason="a CUDA device is present: the demos would run")
@pytest.mark.parametrize("name", ["BallDrop", "SolarSystem"])
def test_demo_without_a_gpu_fails_loudly(built, name, tmp_path):
    exe = os.path.join(HOST, "refdemo", "DEMdemo_" + name)
    if not os.path.exists(exe):
        pytest.skip("refdemo/DEMdemo_%s not built" % name)
    r = subprocess.run(["timeout", "60", exe], cwd=str(tmp_path), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       env=dict(os.environ, DEME_DATA_PATH="/root/reference/data"))
    assert r.returncode != 0
    assert "no usable CUDA device" in r.stdout and "no CPU fallback" in r.stdout, r.stdout[-500:]
