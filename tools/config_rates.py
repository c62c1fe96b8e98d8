#!/usr/bin/env python3
"""Steps/s of the other BASELINE.json configurations on one GPU (not part of the product; bench.py measures C2):
C1 10k frictionless spheres, C4 500k polydisperse clumps in a rotating 50k-facet drum, C5 binning + sort of 5M spheres."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    sys.path.insert(0, p)
from pyapi import demb200, scenes  # noqa: E402


def rate(eng, n):
    eng.sync()
    t0 = time.perf_counter()
    eng.step(n)
    eng.sync()
    return n / (time.perf_counter() - t0)


def rnd(r):
    return json.dumps({k: round(v, 1) for k, v in r.items()})


# ---- C1
f = scenes.flatten(scenes.config1_spheres(n_side=22))
eng = demb200.Engine(0)
eng.load_flat(f)
eng.step(20000)
st = eng.stats()
print("C1: %d spheres, ss %d sa %d: %.0f steps/s   kernels %s" % (f.nSpheres, st.n_contacts_ss, st.n_contacts_sa,
                                                                  rate(eng, 20000), rnd(eng.profile_steps(400))), flush=True)
eng.close()
# ---- C4
sc = scenes.config4_drum(500000, 50000, omega=3.0, init_vel=(0.0, 0.0, -1.5), spacing=2.7)
f = scenes.flatten(sc)
eng = demb200.Engine(0)
eng.load_flat(f)
eng.step(20000)
st = eng.stats()
print("C4: %d clumps (%d spheres) + %d facets, ss %d (touching %d) st %d: %.0f steps/s   kernels %s   rebuild %s" % (
    f.nClumps, f.nSpheres, f.nTri, st.n_contacts_ss, st.n_contacts_ss_touching, st.n_contacts_st, rate(eng, 2000),
    rnd(eng.profile_steps(200)), rnd(eng.profile_rebuild())), flush=True)
eng.close()
