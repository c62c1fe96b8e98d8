#!/usr/bin/env python3
"""GPU tuning helper (not part of the product): settles the C2 bed once, then times kernel variants and the
contact-list update frequency on the same state."""
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
ap.add_argument("--settle-steps", type=int, default=80000)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--spacing", type=float, default=2.7)
ap.add_argument("--freqs", default="")
ap.add_argument("--late-options", default="", help="like --options, timed last (for settings that disturb the bed)")
ap.add_argument("--due", type=int, default=0, help="force_opts value to print the due-candidate counts of a cycle for")
ap.add_argument("--pairs", type=int, default=0, help="print the owner-pair statistics of the settled bed")
ap.add_argument("--options", default="b_agg=0;b_agg=1;b_agg=0;b_agg=1;fast_encode=3;fast_encode=1")
args = ap.parse_args()

sc, dims = bench.build_scene(args.clumps, 20, args.spacing)
f = scenes.flatten(sc)
eng = demb200.Engine(0)
eng.load_flat(f)
t0 = time.time()
eng.step(args.settle_steps)
st = eng.stats()
print("settled %d steps in %.1f s: ss %d (touching %d) sa %d, max|v| %.4f" % (
    args.settle_steps, time.time() - t0, st.n_contacts_ss, st.n_contacts_ss_touching, st.n_contacts_sa,
    eng.reduce(demb200.REDUCE_MAX_ABSV)), flush=True)


import numpy as np  # noqa: E402
if args.pairs:
    idA, idB, ct, wc = eng.contacts()
    ss = ct == 1
    own = np.asarray(f.ownerClumpBody)
    oa, ob = own[idA[ss]].astype("u8"), own[idB[ss]].astype("u8")
    alive = np.abs(wc[ss]).sum(axis=1) > 0
    print("sphere pairs %d, alive %d; distinct owner pairs: all %d, alive %d" % (
        ss.sum(), alive.sum(), len(np.unique(oa * (1 << 32) + ob)), len(np.unique(oa[alive] * (1 << 32) + ob[alive]))), flush=True)


def rnd(r):
    return json.dumps({k: round(v, 1) for k, v in r.items()})


for opt in [o for o in args.options.split(";") if o]:
    name, val = opt.split("=")
    eng.set_option(name, float(val))
    eng.profile_steps(20)
    print("%-16s %s" % (opt, rnd(eng.profile_steps(args.steps))), flush=True)

if args.due:
    # how many candidates are due (not skipped) at each step of a cycle
    eng.set_option("force_opts", float(args.due))
    eng.step(40)
    for k in range(21):
        eng.step(1)
        st = eng.stats()
        nN = st.n_contacts_ss - st.n_contacts_ss_touching
        first = eng.debug_download("sn_due", (nN + 3) // 4).view("u1")[:nN]
        cyc = int(eng.debug_download("flags", 8)[6])
        due = first <= cyc
        print("cycle step %2d: %8d of %8d candidates due (%.1f %%), due byte: median %d max %d; max margin %.3e" % (
            cyc, int(due.sum()), nN, 100.0 * due.mean(), int(np.median(first)), int(first.max()), st.max_margin), flush=True)

for k in [int(x) for x in args.freqs.split(",") if x]:
    eng.params.cd_update_freq = k
    eng.set_params(eng.params)
    eng.step(3 * k)
    eng.sync()
    n = (1200 // k) * k
    t0 = time.perf_counter()
    eng.step(n)
    eng.sync()
    dt = time.perf_counter() - t0
    st = eng.stats()
    print("cd_update_freq=%-3d %8.1f steps/s   ss %d (touching %d)  margin %.6f  rebuild %s" % (
        k, n / dt, st.n_contacts_ss, st.n_contacts_ss_touching, st.max_margin, rnd(eng.profile_rebuild())), flush=True)

for opt in [o for o in args.late_options.split(";") if o]:
    name, val = opt.split("=")
    eng.set_option(name, float(val))
    eng.profile_steps(20)
    print("%-16s %s" % (opt, rnd(eng.profile_steps(args.steps))), flush=True)
