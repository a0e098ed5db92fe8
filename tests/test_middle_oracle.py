"""LangevinMiddle scheme in the CPU oracle: the reference's platform-test
known answers (TestReferenceMmvtLangevinMiddleIntegrator.cpp has the same cases as the non-Middle
file) and the bounce semantics of langevinMiddle.cu:41,202-214."""
import math

import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2
from test_oracle_known_answers import _bond_force


def _plain(masses, precision, T, gamma, dt, elber=False):
    system = s2.System()
    for m in masses:
        system.addParticle(m)
    integ = (s2.ElberLangevinMiddleIntegrator if elber else s2.MmvtLangevinMiddleIntegrator)(T, gamma, dt, "")
    return H.oracle_for(integ, system, precision), integ


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
@pytest.mark.parametrize("elber", [False, True])
def test_single_bond_middle(precision, elber):
    masses = np.array([2.0, 2.0])
    o, integ = _plain(masses, precision, 0.0, 0.1, 0.01, elber)
    assert o.cfg.scheme == L.SCHEME_MIDDLE
    hb = H.make_host_buffers(precision, np.array([[-1.0, 0, 0], [1.0, 0, 0]]), np.zeros((2, 3)), masses)
    rnd = np.zeros((32, 4), dtype=np.float32)
    freq = math.sqrt(1 - 0.05 * 0.05)
    kind = L.KIND_ELBER if elber else L.KIND_MMVT
    for i in range(1000):
        t = o.state().time
        dist = 1.5 + 0.5 * math.exp(-0.05 * t) * math.cos(freq * t)
        x = hb.posq[:2, :3].astype(np.float64)
        np.testing.assert_allclose(x[0], [-0.5 * dist, 0, 0], atol=0.02)
        np.testing.assert_allclose(x[1], [0.5 * dist, 0, 0], atol=0.02)
        _bond_force(hb)
        assert o.step(kind, hb.posq, hb.corr, hb.velm, hb.force, rnd, 0) == 0


def test_temperature_middle():
    n, temp = 8, 100.0
    masses = np.full(n, 2.0)
    o, _ = _plain(masses, L.MIXED, temp, 2.0, 0.01)
    pos = np.array([[(2 if i % 2 == 0 else -2), (2 if i % 4 < 2 else -2), (2 if i < 4 else -2)] for i in range(n)], dtype=np.float64)
    hb = H.make_host_buffers(L.MIXED, pos, np.zeros((n, 3)), masses)
    x0 = hb.posq.copy()
    pool = H.gaussian_pool(H.SEED_BASE + 50, 20000 * 32)
    assert o.run(10000, hb.posq, hb.corr, hb.velm, hb.force, pool, 0, x0, 10.0) == 0
    ke = 0.0
    for i in range(10000):
        v = hb.velm[:n, :3]
        ke += 0.5 * (masses[:, None] * v * v).sum()
        o.run(1, hb.posq, hb.corr, hb.velm, hb.force, pool, (10000 + i) * 32, x0, 10.0)
    expected = 0.5 * n * 3 * s2.BOLTZ * temp
    assert abs(ke / 10000 - expected) / expected < 6 / math.sqrt(10000.0)


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
def test_middle_bounce_negates_force_kicked_velocity(precision):
    """oldVelm is saved in Part2 BEFORE the O step, i.e. after the force kick only (langevinMiddle.cu:41);
    the bounce writes -oldVelm and the pre-Part3 position (:202-214)."""
    syn = S.ballistic_toy(speed=0.5, dt=0.01)
    syn.positions[0, 0] = 1.2 + 0.998
    syn.friction, syn.temperature = 1.0, 300.0
    integ = H.make_mmvt_integrator(syn, "", middle=True)
    before = H.make_host_buffers(precision, syn.positions, syn.velocities, syn.masses,
                                 forces=np.array([[250.0, -30.0, 12.0]]))
    hb = before.copy()
    o = H.oracle_for(integ, syn.system, precision)
    pool = H.gaussian_pool(3, 32)
    assert o.step(L.KIND_MMVT, hb.posq, hb.corr, hb.velm, hb.force, pool, 0) == 0
    assert len(o.records()) == 1
    np.testing.assert_array_equal(hb.posq[:1], before.posq[:1])
    mixed = H.mixed_dtype(precision)
    fs = mixed(0.01) / mixed(4294967296.0)
    fw = fs * before.velm[0, 3]
    # fma(fw, F, v) computed exactly in wider precision then rounded once
    v1 = mixed(np.longdouble(fw) * np.longdouble(mixed(before.force[0, 0])) + np.longdouble(before.velm[0, 0]))
    assert hb.velm[0, 0] == -v1


def test_middle_differs_from_langevin_and_reference_variant():
    syn = H.toy_system()
    pool = H.gaussian_pool(9, 200 * 32)
    out = {}
    for name, middle, variant in (("lang", False, "cuda"), ("mid", True, "cuda"), ("mid_ref", True, "reference")):
        integ = H.make_mmvt_integrator(syn, "", variant, middle=middle)
        rc, o, hb = S.run_oracle(syn, integ, L.MIXED, 200, pool)
        out[name] = hb.posq[0, :3].copy()
    assert (out["lang"] != out["mid"]).any()
