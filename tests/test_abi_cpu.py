"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/dem_b200.h
declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from pyapi import demb200, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "dem_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|const char\*)\s+(dem_[a-z0-9_]+)\s*\(", hdr, re.M))
    assert len(declared) >= 30
    lib = demb200.load_library()
    for sym in sorted(declared):
        assert getattr(lib, sym) is not None, sym
    assert declared == set(demb200.EXPORTED_SYMBOLS)
    assert lib.dem_abi_version() == 1


def test_struct_sizes_match_header(built):
    assert C.sizeof(demb200.DemSimParams) == 120  # static_assert in csrc/dem_core.cu
    assert demb200.PRESC_DTYPE.itemsize == 88


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(demb200.DemError) as e:
        demb200.Engine(0)
    assert e.value.code == demb200.DEM_ERR_NO_GPU


def test_product_never_imports_oracle():
    """The product path (csrc + pyapi) must not reference the oracle."""
    pkg = os.path.join(ROOT, "dem-engine_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".cu", ".cuh", ".h", ".py", ".cpp", ".hpp")):
                txt = open(os.path.join(dirpath, fn)).read()
                assert "liboracle" not in txt and "dem_oracle" not in txt and "pyoracle" not in txt, fn


def test_flatten_layout(built):
    sc = scenes.config2_clumps(3, 3, 2, spacing=2.7)
    f = scenes.flatten(sc)
    assert f.nClumps == 18 and f.nSpheres == 54 and f.nOwners == 19 and f.nAnal == 5
    assert f.nvXp2 + f.nvYp2 + f.nvZp2 == 64
    assert np.all(f.ownerClumpBody == np.repeat(np.arange(18), 3))
    assert np.all(f.clumpComponentOffset == np.tile(np.arange(3), 18))
    assert f.familyID[-1] == 255 and f.inertiaPropOffsets[-1] == 1
    # encode/decode round trip within one length unit
    from oracle import pyoracle
    w = pyoracle.world_from_flat(f)
    back = w.positions_f64()[:18]
    assert np.abs(back - sc.clump_xyz.astype("f8")).max() < 1e-6
