"""The CUDA path against the COMMITTED golden vectors (tests/golden/golden_*.npz), without any oracle call: the fixtures were
dumped from the reference's own kernel text compiled for the host (tests/golden/make_golden.py); the CPU suite checks that the
oracle reproduces them bit for bit.  (Named to run last in the GPU suite.)"""
import numpy as np
import pytest

from pyapi import demb200, scenes
from test_gpu_parity import EARLY_BOUND_V, EARLY_BOUND_X, _mk

pytestmark = pytest.mark.gpu


def _golden_cases():
    from test_oracle_vs_ref import GOLDEN_CASES
    return GOLDEN_CASES


@pytest.mark.parametrize("kind", _golden_cases())
def test_early_state_matches_golden_fixture(built, kind):
    """The CUDA path against the COMMITTED golden vectors (tests/golden/golden_*.npz, dumped from the reference's own kernel
    text by tests/golden/make_golden.py; the oracle reproduces them bit for bit on the CPU side): state after 200 steps,
    before collisions have amplified round-off, within the fixed bounds 1e-7 m / 1e-3 m/s.  No oracle call on this path."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_%s.npz" % kind))
    f = scenes.flatten(_mk(kind))
    eng = demb200.Engine(0)
    eng.load_flat(f)
    eng.step(int(g["early_nsteps"]))
    nC = f.nClumps
    err_x = np.abs(eng.positions()[:nC] - g["early_pos"]).max()
    st = eng.owner_state()
    err_v = np.abs(st["vel"][:nC] - g["early_vel"]).max()
    # orientation: compare as rotations (q and -q are the same orientation)
    q, qg = st["oriQ"][:nC].astype("f8"), g["early_quat"].astype("f8")
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    qg /= np.linalg.norm(qg, axis=1, keepdims=True)
    err_q = (1.0 - np.abs((q * qg).sum(1))).max()
    print("%s: after %d steps |dx| %.2e m, |dv| %.2e m/s, 1 - |q.q_golden| %.2e" % (kind, int(g["early_nsteps"]), err_x, err_v, err_q))
    assert err_x <= EARLY_BOUND_X and err_v <= EARLY_BOUND_V and err_q <= 1e-6, (err_x, err_v, err_q)
    eng.close()
