"""GPU known-answer tests: the same analytic checks as tests/test_oracle_analytic_cpu.py, on the CUDA path through the C
ABI (static weight on a plane and Hertz depth, two-sphere collision, rolling vs slipping down an incline)."""
import math

import numpy as np
import pytest

from pyapi import demb200, scenes
from test_oracle_analytic_cpu import _ball_scene

pytestmark = pytest.mark.gpu


def test_static_weight_on_plane_gpu(built):
    R, E, nu = 0.01, 1e8, 0.3
    sc, t, mass = _ball_scene(radius=R, E=E, nu=nu, CoR=0.2)
    sc.add_clumps(t, [[0.0, 0.0, -0.5 + R - 1e-6]])
    f = scenes.flatten(sc)
    eng = demb200.Engine(0)
    eng.load_flat(f)
    eng.step(30000)
    st = eng.owner_state()
    assert abs(float(st["vel"][0, 2])) < 1e-5
    idA, idB, ct, wc, fr, pt = eng.contact_records()
    sa = ct != 1
    assert sa.sum() >= 1
    assert fr[sa][:, 2].sum() == pytest.approx(mass * 9.81, rel=2e-3)
    Estar = E / (2.0 * (1.0 - nu * nu))
    depth = (mass * 9.81 / (4.0 / 3.0 * Estar * math.sqrt(R))) ** (2.0 / 3.0)
    z = eng.positions()[0, 2]
    assert (-0.5 + R) - z == pytest.approx(depth, rel=5e-3)
    # the contact point of the record lies under the ball, on the plane (to within the overlap)
    touching = sa & (np.abs(fr).max(1) > 0)
    assert np.abs(pt[touching][:, :2]).max() < 1e-6 and abs(pt[touching][0, 2] + 0.5) < 2 * depth
    eng.close()


def test_two_sphere_collision_gpu(built):
    sc = scenes.Scene()
    m = sc.load_material(E=1e8, nu=0.3, CoR=0.7, mu=0.0, Crr=0.0)
    rho, r1, r2 = 2600.0, 0.01, 0.015
    m1, m2 = (rho * 4.0 / 3.0 * math.pi * r ** 3 for r in (r1, r2))
    t1, t2 = sc.load_sphere_type(m1, r1, m), sc.load_sphere_type(m2, r2, m)
    gap = 1e-4
    sc.add_clumps(t1, [[-(r1 + gap / 2), 0, 0]], vel=(0.5, 0, 0))
    sc.add_clumps(t2, [[(r2 + gap / 2), 0, 0]], vel=(-0.3, 0, 0))
    sc.box, sc.G, sc.h, sc.cd_update_freq = (1.0, 1.0, 1.0), (0, 0, 0), 1e-6, 10
    f = scenes.flatten(sc)
    eng = demb200.Engine(0)
    eng.load_flat(f)
    touching = 0
    for k in range(250):
        eng.step(10)
        p = eng.positions()
        touching += 10 * int(p[1, 0] - p[0, 0] < r1 + r2)
    v = eng.owner_state()["vel"]
    v1, v2 = float(v[0, 0]), float(v[1, 0])
    assert m1 * v1 + m2 * v2 == pytest.approx(m1 * 0.5 - m2 * 0.3, rel=1e-5)
    e = (v2 - v1) / 0.8
    assert 0.45 < e < 0.85
    mstar, Rstar, Estar = m1 * m2 / (m1 + m2), r1 * r2 / (r1 + r2), 1e8 / (2.0 * (1.0 - 0.09))
    t_hertz = 2.87 * (mstar ** 2 / (Rstar * Estar ** 2 * 0.8)) ** 0.2
    assert touching * 1e-6 == pytest.approx(t_hertz, rel=0.17)
    eng.close()


@pytest.mark.parametrize("mu,regime", [(0.5, "rolls"), (0.05, "slips")])
def test_ball_on_incline_gpu(built, mu, regime):
    theta, g, R = math.radians(20.0), 9.81, 0.01
    sc, t, mass = _ball_scene(radius=R, mu=mu, CoR=0.3, h=5e-6, G=(g * math.sin(theta), 0.0, -g * math.cos(theta)),
                              box=(2.0, 0.5, 0.5))
    Estar = 1e8 / (2.0 * (1.0 - 0.09))
    depth = (mass * g * math.cos(theta) / (4.0 / 3.0 * Estar * math.sqrt(R))) ** (2.0 / 3.0)
    sc.add_clumps(t, [[-0.8, 0.0, -0.25 + R - depth]])
    f = scenes.flatten(sc)
    eng = demb200.Engine(0)
    eng.load_flat(f)
    T = 0.1
    eng.step(int(round(T / 5e-6)))
    st = eng.owner_state()
    v, om = float(st["vel"][0, 0]), float(st["omg"][0, 1])
    if regime == "rolls":
        assert v == pytest.approx(5.0 / 7.0 * g * math.sin(theta) * T, rel=0.02)
        assert om * R == pytest.approx(v, rel=0.02)
    else:
        assert v == pytest.approx(g * (math.sin(theta) - mu * math.cos(theta)) * T, rel=0.02)
        assert om == pytest.approx(2.5 * mu * g * math.cos(theta) / R * T, rel=0.03)
    eng.close()
