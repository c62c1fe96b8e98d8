"""Print (no asserts) what the GPU tests that have not yet run on hardware would measure: the CUDA path against the committed
golden fixtures at step 200 and the sphere-level reductions.  Writes progressively so that a cut-off run still leaves output."""
import os, sys, time
t0 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dem-engine_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
from pyapi import demb200, scenes
from test_oracle_vs_ref import GOLDEN_CASES, _scene

out = open(os.path.join(ROOT, "gpurun_out", "last_check.log"), "w") if "--stdout" not in sys.argv else sys.stdout
def say(s):
    out.write("[%5.1fs] %s\n" % (time.time() - t0, s)); out.flush()

say("imports done")
if "--dry" in sys.argv:
    for kind in GOLDEN_CASES:
        f = scenes.flatten(_scene(kind)); say("flattened " + kind)
    sys.exit(0)
for kind in GOLDEN_CASES:
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "golden_%s.npz" % kind))
        f = scenes.flatten(_scene(kind))
        eng = demb200.Engine(0); eng.load_flat(f); eng.step(int(g["early_nsteps"]))
        nC = f.nClumps
        st = eng.owner_state()
        err_x = np.abs(eng.positions()[:nC] - g["early_pos"]).max()
        err_v = np.abs(st["vel"][:nC] - g["early_vel"]).max()
        q, qg = st["oriQ"][:nC].astype("f8"), g["early_quat"].astype("f8")
        q /= np.linalg.norm(q, axis=1, keepdims=True); qg /= np.linalg.norm(qg, axis=1, keepdims=True)
        err_q = (1.0 - np.abs((q * qg).sum(1))).max()
        say("golden %s: step %d |dx| %.3e m |dv| %.3e m/s 1-|q.qg| %.3e" % (kind, int(g["early_nsteps"]), err_x, err_v, err_q))
        eng.close()
    except Exception as e:
        say("golden %s: EXCEPTION %r" % (kind, e))
try:
    sys.argv = sys.argv[:1]
    import test_gpu_zz_golden as tz
    tz.test_sphere_level_reductions_match_numpy(True)
    say("sphere-level reductions: PASS")
except BaseException as e:
    say("sphere-level reductions: FAIL %r" % (e,))
say("done")
