import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dem-engine_b200")); sys.path.insert(0, ROOT)
import numpy as np
from pyapi import demb200, scenes
from oracle import pyoracle as po

omega = float(sys.argv[1]) if len(sys.argv) > 1 else 6.0
sc = scenes.config4_drum(2000, 1500, omega=omega, init_vel=(0.2, 0.0, -1.0), cd_update_freq=10, spacing=2.7)
f = scenes.flatten(sc)
eng = demb200.Engine(0); eng.load_flat(f)
w = po.world_from_flat(f)
wp = w.copy()
for name in ("vX", "vY", "vZ", "omgBarX", "omgBarY", "omgBarZ"):
    a = getattr(wp, name); a[:] = (a.astype("f8") * (1.0 + 1e-6)).astype("f4")
nC = f.nClumps; d = f.nOwners - 1
def touching(w):
    n = w.nContacts
    F = w.contactForces[:3*n].reshape(-1, 3); on = np.abs(F).max(1) > 0
    return set(zip(w.idGeometryA[:n][on].tolist(), w.idGeometryB[:n][on].tolist(), w.contactType[:n][on].tolist()))
for cp in range(100, 2600, 100):
    eng.step(100); w.step(100, cd_every=f.cd_update_freq); wp.step(100, cd_every=f.cd_update_freq)
    pw = w.positions_f64()[:nC]
    sens = np.abs(wp.positions_f64()[:nC] - pw).max()
    pe = eng.positions()[:nC]
    err = np.abs(pe - pw).max()
    st = eng.owner_state()
    q = st["oriQ"][d]; qo = np.array([w.oriQw[d], w.oriQx[d], w.oriQy[d], w.oriQz[d]])
    idA, idB, ct, wc = eng.contacts()
    alive = np.abs(wc).max(1) > 0
    mine = set(zip(idA[alive].tolist(), idB[alive].tolist(), ct[alive].tolist()))
    th = touching(w)
    worst = int(np.abs(pe - pw).max(1).argmax())
    print("step %4d err %.2e sens %.2e  dq %.2e  touching oracle %d (tri %d) device-alive %d  only-oracle %d only-device %d worst clump %d" % (
        cp, err, sens, np.abs(q - qo).max(), len(th), sum(1 for t in th if t[2] == 2), len(mine), len(th - mine), len(mine - th), worst))
    if len(th - mine) and cp <= 1600:
        print("   only oracle:", sorted(th - mine)[:6])
    if len(mine - th) and cp <= 1600:
        print("   only device:", sorted(mine - th)[:6])
