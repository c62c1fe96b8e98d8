#!/usr/bin/env python3
"""Times the sphere--sphere force kernel of the settled C2 bed with the candidate skip / lazy kinematics fetch switched
on and off (dem_set_option("force_opts")).  Not part of the product."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402
from pyapi import demb200, scenes  # noqa: E402

sc, dims = bench.build_scene(1000000, 20, 2.7)
f = scenes.flatten(sc)
eng = demb200.Engine(0)
eng.load_flat(f)
eng.step(int(sys.argv[1]) if len(sys.argv) > 1 else 80000)
for ctas, opts in ((4, 3), (3, 3), (3, 0), (4, 0), (4, 4), (4, 8 | 3), (3, 4), (3, 8 | 3)):
    eng.set_option("ctas_per_sm", ctas)
    eng.set_option("force_opts", opts)
    eng.step(40)
    print("ctas", ctas, "force_opts", opts, json.dumps({k: round(float(v), 2) for k, v in eng.profile_steps(200).items()}), flush=True)
eng.set_option("force_opts", 3)
for r in range(2):
    print("rebuild", json.dumps({k: round(v, 1) for k, v in eng.profile_rebuild().items()}))
