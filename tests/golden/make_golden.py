#!/usr/bin/env python3
"""Generates tests/golden/golden_*.npz by running the reference's own kernel text (oracle/_ref/libdemref.so, built from
/root/reference by oracle/Makefile) on small seeded scenes. Run in the build container (needs the reference tree):
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

from oracle import pyoracle  # noqa: E402
from pyapi import scenes  # noqa: E402
from test_oracle_vs_ref import GOLDEN_CASES, STATE, _scene  # noqa: E402

NSTEPS = 5000
EARLY = 200  # a second snapshot before collisions have amplified round-off: what the GPU path is held against directly
for kind in GOLDEN_CASES:
    f = scenes.flatten(_scene(kind))
    w = pyoracle.world_from_flat(f)
    assert pyoracle.ref() is not None, "build oracle/_ref first (make -C oracle)"
    we = w.copy()
    we.step(EARLY, cd_every=f.cd_update_freq, use_ref=True)
    w.step(NSTEPS, cd_every=f.cd_update_freq, use_ref=True)
    out = {name: getattr(w, name)[: w.nOwners].copy() for name in STATE}
    out.update(early_nsteps=EARLY, early_pos=we.positions_f64()[: f.nClumps].copy(),
               early_vel=np.stack([we.vX, we.vY, we.vZ], 1)[: f.nClumps].copy(),
               early_quat=np.stack([we.oriQw, we.oriQx, we.oriQy, we.oriQz], 1)[: f.nClumps].copy())
    out.update(nsteps=NSTEPS, nContacts=w.nContacts, idGeometryA=w.idGeometryA[: w.nContacts].copy(),
               idGeometryB=w.idGeometryB[: w.nContacts].copy(), contactType=w.contactType[: w.nContacts].copy(),
               wildcards=np.stack([c[: w.nContacts] for c in w.contactWildcards], 1))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_%s.npz" % kind)
    np.savez_compressed(path, **out)
    print(kind, "owners", w.nOwners, "contacts", w.nContacts, "->", path, os.path.getsize(path), "bytes")
