"""CPU-only: analytic / known-answer checks of the oracle (and of the reference's own host-compiled kernels, oracle/_ref,
when present) -- the checks the reference's demos imply (SURVEY.md 8c ii): static weight on a plane, momentum in a
two-sphere collision and Hertzian contact time, rolling vs slipping down an incline
(DEMdemo_TestPack.cpp:98-203 set-up), energy decay.  They pin the PHYSICS of the parity anchor, next to the bit-exact
comparison with the reference's kernel text (test_oracle_vs_ref.py)."""
import math

import numpy as np
import pytest

from oracle import pyoracle
from pyapi import demb200, scenes

USE_REF = [False] + ([True] if pyoracle.ref() is not None else [])


def _ball_scene(radius=0.01, rho=2600.0, E=1e8, nu=0.3, CoR=0.5, mu=0.5, Crr=0.0, h=1e-5, G=(0, 0, -9.81), box=(1.0, 1.0, 1.0)):
    sc = scenes.Scene()
    m = sc.load_material(E=E, nu=nu, CoR=CoR, mu=mu, Crr=Crr)
    mass = rho * 4.0 / 3.0 * math.pi * radius ** 3
    t = sc.load_sphere_type(mass, radius, m)
    sc.box, sc.G, sc.h, sc.cd_update_freq = box, G, h, 10
    sc.bounding, sc.bounding_mat = "only_bottom", m
    sc.record_contact_forces = 1
    return sc, t, mass


@pytest.mark.parametrize("use_ref", USE_REF)
def test_static_weight_on_plane(use_ref):
    """A ball at rest on the bottom plane: the plane carries exactly its weight, the ball sinks by the Hertz depth."""
    R, E, nu = 0.01, 1e8, 0.3
    sc, t, mass = _ball_scene(radius=R, E=E, nu=nu, CoR=0.2)
    z0 = -0.5 + R - 1e-6
    sc.add_clumps(t, [[0.0, 0.0, z0]])
    f = scenes.flatten(sc)
    w = pyoracle.world_from_flat(f)
    w.step(30000, cd_every=10, use_ref=use_ref)       # 0.3 s: the damped bounce has died out
    assert abs(float(w.vZ[0])) < 1e-5
    n = w.nContacts
    F = w.contactForces[: 3 * n].reshape(-1, 3)
    sa = w.contactType[:n] != 1
    assert sa.sum() >= 1
    assert F[sa][:, 2].sum() == pytest.approx(mass * 9.81, rel=2e-3)
    # Hertz: F = 4/3 E* sqrt(R) d^1.5 with E* = E / (2 (1 - nu^2)) for equal materials (wall radius -> infinity)
    Estar = E / (2.0 * (1.0 - nu * nu))
    depth = (mass * 9.81 / (4.0 / 3.0 * Estar * math.sqrt(R))) ** (2.0 / 3.0)
    z = w.positions_f64()[0, 2]
    assert (-0.5 + R) - z == pytest.approx(depth, rel=5e-3)


@pytest.mark.parametrize("use_ref", USE_REF)
def test_two_sphere_collision_momentum_and_contact_time(use_ref):
    """Head-on collision of two different balls without gravity: momentum is conserved to round-off, kinetic energy
    drops (CoR < 1), and the contact lasts the Hertzian contact time 2.87 (m*^2 / (R* E*^2 v))^(1/5) within 15 %
    (the damped contact is a little longer than the elastic one)."""
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
    w = pyoracle.world_from_flat(f)
    p0 = m1 * 0.5 - m2 * 0.3
    ke0 = 0.5 * m1 * 0.25 + 0.5 * m2 * 0.09
    touching = []
    for k in range(2500):
        w.step(1, cd_every=10, use_ref=use_ref)
        d = w.positions_f64()[1, 0] - w.positions_f64()[0, 0]
        touching.append(d < r1 + r2)
    v1, v2 = float(w.vX[0]), float(w.vX[1])
    assert m1 * v1 + m2 * v2 == pytest.approx(p0, rel=1e-5)
    assert v2 - v1 > 0                                             # they separate again
    ke1 = 0.5 * m1 * v1 * v1 + 0.5 * m2 * v2 * v2
    assert 0.3 * ke0 < ke1 < ke0
    e = (v2 - v1) / 0.8
    assert 0.45 < e < 0.85                                         # the reference's Hertzian damping realises CoR 0.7 roughly
    t_contact = sum(touching) * 1e-6
    mstar, Rstar, Estar = m1 * m2 / (m1 + m2), r1 * r2 / (r1 + r2), 1e8 / (2.0 * (1.0 - 0.09))
    t_hertz = 2.87 * (mstar ** 2 / (Rstar * Estar ** 2 * 0.8)) ** 0.2
    assert t_contact == pytest.approx(t_hertz, rel=0.15), (t_contact, t_hertz)


@pytest.mark.parametrize("use_ref", USE_REF)
@pytest.mark.parametrize("mu,regime", [(0.5, "rolls"), (0.05, "slips")])
def test_ball_on_incline_rolls_or_slips(use_ref, mu, regime):
    """A ball released on a 20-degree incline (gravity tilted instead of the plane).  With mu >= 2/7 tan(theta) it rolls
    without slipping, a = 5/7 g sin(theta) and v = omega R; below that it slips, a = g (sin(theta) - mu cos(theta)) and
    the spin follows the friction torque, alpha = 5/2 mu g cos(theta) / R."""
    theta, g, R = math.radians(20.0), 9.81, 0.01
    assert (mu >= 2.0 / 7.0 * math.tan(theta)) == (regime == "rolls")
    sc, t, mass = _ball_scene(radius=R, mu=mu, CoR=0.3, h=5e-6, G=(g * math.sin(theta), 0.0, -g * math.cos(theta)),
                              box=(2.0, 0.5, 0.5))
    Estar = 1e8 / (2.0 * (1.0 - 0.09))
    depth = (mass * g * math.cos(theta) / (4.0 / 3.0 * Estar * math.sqrt(R))) ** (2.0 / 3.0)
    sc.add_clumps(t, [[-0.8, 0.0, -0.25 + R - depth]])            # starts in static equilibrium normal to the plane
    f = scenes.flatten(sc)
    w = pyoracle.world_from_flat(f)
    T = 0.1
    w.step(int(round(T / 5e-6)), cd_every=10, use_ref=use_ref)
    v, om = float(w.vX[0]), float(w.omgBarY[0])
    if regime == "rolls":
        assert v == pytest.approx(5.0 / 7.0 * g * math.sin(theta) * T, rel=0.02)
        assert om * R == pytest.approx(v, rel=0.02)               # no slip at the contact point
    else:
        assert v == pytest.approx(g * (math.sin(theta) - mu * math.cos(theta)) * T, rel=0.02)
        assert om == pytest.approx(2.5 * mu * g * math.cos(theta) / R * T, rel=0.03)
        assert om * R < 0.5 * v                                     # sliding much faster than it spins
    assert abs(float(w.vY[0])) < 1e-6 and abs(float(w.vZ[0])) < 2e-3


def test_bed_energy_decays_and_quaternions_stay_unit():
    """A small clump bed dropped into a box: kinetic + gravitational energy never exceeds its initial value (the elastic
    energy stored in the contacts only borrows from it, damping and friction only remove), it has dropped markedly by
    the end, and orientations stay unit quaternions."""
    # lattice spacing 3.2 clump scales: wider than two circumscribed radii (2 x 1.458), so no clump starts in overlap
    # with a neighbour (an initial overlap would release elastic energy)
    sc = scenes.config2_clumps(4, 4, 3, cd_update_freq=10, spacing=3.2, init_vel=(0.2, 0.1, -1.0))
    f = scenes.flatten(sc)
    w = pyoracle.world_from_flat(f)
    nC = f.nClumps
    mass = float(np.asarray(f.MassProperties).ravel()[0])
    moi = np.array([f.moiX[0], f.moiY[0], f.moiZ[0]], "f8")

    def energy():
        v = np.stack([w.vX, w.vY, w.vZ], 1)[:nC].astype("f8")
        om = np.stack([w.omgBarX, w.omgBarY, w.omgBarZ], 1)[:nC].astype("f8")
        z = w.positions_f64()[:nC, 2]
        return 0.5 * mass * (v * v).sum() + 0.5 * (om * om * moi).sum() + mass * 9.81 * z.sum()

    e0 = energy()
    ke0 = 0.5 * mass * nC * (0.2 ** 2 + 0.1 ** 2 + 1.0 ** 2)
    for k in range(10):
        w.step(400, cd_every=10)
        e = energy()
        assert e <= e0 + 1e-6 * ke0
        q = np.stack([w.oriQw, w.oriQx, w.oriQy, w.oriQz], 1)[:nC].astype("f8")
        assert np.abs(np.sqrt((q * q).sum(1)) - 1.0).max() < 1e-5
    assert e0 - e > 0.3 * ke0      # the impacts have dissipated a good part of the initial kinetic energy
