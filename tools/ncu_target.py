#!/usr/bin/env python3
"""ncu target (not part of the product): settles the C2 bed unprofiled, then runs a few steps between
cudaProfilerStart/Stop so that `ncu --profile-from-start off` sees exactly those launches.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/ncu_target.py --steps 45
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import bench  # noqa: E402
from pyapi import demb200, scenes  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--clumps", type=int, default=1000000)
ap.add_argument("--settle-steps", type=int, default=80000)
ap.add_argument("--steps", type=int, default=45)
ap.add_argument("--cd-update-freq", type=int, default=20)
ap.add_argument("--spacing", type=float, default=2.7)
ap.add_argument("--ctas-per-sm", type=int, default=4)
args = ap.parse_args()
sc, dims = bench.build_scene(args.clumps, args.cd_update_freq, args.spacing)
f = scenes.flatten(sc)
eng = demb200.Engine(0)
eng.load_flat(f)
eng.set_option("ctas_per_sm", args.ctas_per_sm)
eng.step(args.settle_steps)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.step(args.steps)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
st = eng.stats()
print("profiled %d steps; contacts ss %d sa %d" % (args.steps, st.n_contacts_ss, st.n_contacts_sa))
