#!/usr/bin/env python3
"""GPU tuning / profiling helper (not part of the product): settles the C2 bed, then times kernel variants."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from pyapi import demb200, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clumps", type=int, default=1000000)
ap.add_argument("--settle-steps", type=int, default=12000)
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--cd-update-freq", type=int, default=20)
ap.add_argument("--spacing", type=float, default=2.7)
args = ap.parse_args()

sc, dims = bench.build_scene(args.clumps, args.cd_update_freq, args.spacing)
f = scenes.flatten(sc)
eng = demb200.Engine(0)
eng.load_flat(f)
t0 = time.time()
done = 0
while done < args.settle_steps:
    n = min(10000, args.settle_steps - done)
    eng.step(n)
    done += n
    st = eng.stats()
    print("  step %d: max|v| %.3f m/s  KE %.4e J  max z %.4f  ss %d (touching %d) sa %d  wall %.1f s" % (
        done, eng.reduce(demb200.REDUCE_MAX_ABSV), eng.reduce(demb200.REDUCE_KINETIC_ENERGY),
        eng.reduce(demb200.REDUCE_MAX_Z), st.n_contacts_ss, st.n_contacts_ss_touching, st.n_contacts_sa, time.time() - t0), flush=True)
print("settled %d steps in %.1f s" % (args.settle_steps, time.time() - t0), flush=True)
st = eng.stats()
print("contacts ss %d sa %d cells %s cs %.5f margin %.6f" % (st.n_contacts_ss, st.n_contacts_sa, list(st.n_cells), st.cell_size, st.max_margin))
print("touching", st.n_contacts_ss_touching)
for ctas in (2, 3, 4):
    eng.set_option("ctas_per_sm", ctas)
    eng.profile_steps(20)
    r = eng.profile_steps(args.steps)
    print("ctas_per_sm=%d  %s" % (ctas, json.dumps({k: round(v, 1) for k, v in r.items()})), flush=True)
eng.set_option("ctas_per_sm", 3)
for sm in (0, 1):
    eng.set_option("sort_mode", sm)
    for i in range(2):
        print("sort_mode", sm, "rebuild", json.dumps({k: round(v, 1) for k, v in eng.profile_rebuild().items()}), flush=True)
t0 = time.time()
eng.step(2000)
print("2000 steps wall %.3f s -> %.1f steps/s" % (time.time() - t0, 2000 / (time.time() - t0)))
