#!/usr/bin/env python3
"""Times the UNMODIFIED reference (projectchrono/DEM-Engine built by baseline/build_ref.sh into baseline/_ref) on the GPU(s)
of this box, on the SAME settled bed bench.py times our engine on:

    python tools/run_reference_gpu.py [--clumps 1000000] [--settle-steps 80000] [--steps 400] [--gpus 1,2]

Our engine is only used to produce the settled bed (the reference would need minutes for that); the timed part is the
reference's own DEMSolver(nGPUs) -> DoDynamicsThenSync (src/DEM/APIPublic.cpp:2446-2479) through baseline/run_ref.cpp.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

RUN_REF = os.path.join(ROOT, "baseline", "_ref", "run_ref")


def reference_available():
    return os.path.exists(RUN_REF) and os.path.isdir(os.path.join(ROOT, "baseline", "_ref", "build", "kernel"))


def dump_settled_scene(eng, sc, f, path):
    """scene file of the bed as it stands in `eng` (positions as the reference's own getters report them: float)"""
    from pyapi import scenes
    n = f.nClumps
    st = eng.owner_state(0, n)
    pos = eng.positions(0, n)
    scenes.write_scene_file(path, sc, xyz=pos.astype("f4"), quat=st["oriQ"], vel=st["vel"], omg=st["omg"])
    return path


def run_reference(scene_path, n_gpus, steps, warmup, cd_update_freq=0, timeout=900, visible=None, exe=None, mode="bench"):
    """one run of the driver script baseline/run_ref.cpp (exe: baseline/_ref/run_ref = the unmodified reference, or
    dem-engine_b200/host/run_b200 = the same script against this repository's facade); returns its JSON line (dict) or
    {"unavailable": why}"""
    if exe is None:
        if not reference_available():
            return {"unavailable": "baseline/_ref/run_ref not built (baseline/build_ref.sh needs /root/reference)"}
        exe = RUN_REF
    elif not os.path.exists(exe):
        return {"unavailable": "%s not built" % os.path.relpath(exe, ROOT)}
    env = dict(os.environ)
    if exe != RUN_REF:  # (the reference has its data directory baked in at configure time; the facade reads this one)
        env.setdefault("DEME_DATA_PATH", os.path.join(ROOT, "dem-engine_b200", "host", "data"))
    if visible is not None:
        env["CUDA_VISIBLE_DEVICES"] = visible
    cmd = [exe, mode, scene_path, str(n_gpus), str(steps), str(warmup), str(cd_update_freq)]
    try:
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    except subprocess.TimeoutExpired:
        return {"unavailable": "run_ref timed out after %d s" % timeout}
    for line in r.stdout.splitlines():
        if line.startswith("{"):
            try:
                d = json.loads(line)
                d["stats_tail"] = r.stdout[-1500:]
                return d
            except ValueError:
                pass
    return {"unavailable": "run_ref failed rc=%d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:].replace("\n", " | "))}


def main():
    import bench
    from pyapi import demb200, scenes
    ap = argparse.ArgumentParser()
    ap.add_argument("--clumps", type=int, default=1000000)
    ap.add_argument("--settle-steps", type=int, default=80000)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--cd-update-freq", type=int, default=20)
    ap.add_argument("--spacing", type=float, default=2.7)
    ap.add_argument("--gpus", default="1,2")
    args = ap.parse_args()
    sc, dims = bench.build_scene(args.clumps, args.cd_update_freq, args.spacing)
    f = scenes.flatten(sc)
    eng = demb200.Engine(0)
    eng.load_flat(f)
    eng.step(args.settle_steps)
    path = os.path.join(tempfile.gettempdir(), "dem_c2_settled_%d.bin" % args.clumps)
    dump_settled_scene(eng, sc, f, path)
    eng.close()
    import torch
    have = torch.cuda.device_count()
    for g in [int(v) for v in args.gpus.split(",")]:
        if g > have:
            print(json.dumps({"n_gpus": g, "unavailable": "box has %d GPU(s)" % have}))
            continue
        print(json.dumps(run_reference(path, g, args.steps, args.warmup)), flush=True)


if __name__ == "__main__":
    main()
