#!/usr/bin/env python3
"""Settles the C2 bed, then prints the per-stage device times of a contact-list rebuild (dem_profile_rebuild), the
per-kernel times of a step (dem_profile_steps) and the list sizes.  Not part of the product."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from pyapi import demb200, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clumps", type=int, default=1000000)
ap.add_argument("--settle-steps", type=int, default=80000)
ap.add_argument("--cd-update-freq", type=int, default=20)
ap.add_argument("--spacing", type=float, default=2.7)
ap.add_argument("--repeats", type=int, default=3)
args = ap.parse_args()
sc, dims = bench.build_scene(args.clumps, args.cd_update_freq, args.spacing)
f = scenes.flatten(sc)
eng = demb200.Engine(0)
eng.load_flat(f)
eng.step(args.settle_steps)
for r in range(args.repeats):
    print("rebuild", json.dumps({k: round(v, 1) for k, v in eng.profile_rebuild().items()}))
    eng.step(args.cd_update_freq)
print("steps", json.dumps({k: round(float(v), 2) for k, v in eng.profile_steps(200).items()}))
st = eng.stats()
print("contacts ss %d (touching %d) sa %d; cells %s; margin %.3e; overflow %d; capacity %d" % (
    st.n_contacts_ss, st.n_contacts_ss_touching, st.n_contacts_sa, list(st.n_cells), st.max_margin, st.overflow, st.contact_capacity))
