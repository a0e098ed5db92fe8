"""Pins the CPU oracle's Langevin arithmetic on the known-answer cases of the reference's own platform
tests (seekr2plugin/platforms/reference/tests/TestReferenceMmvtLangevinIntegrator.cpp:63-284), restated
for a harness force model because OpenMM's force fields are not available."""
import math

import numpy as np
import pytest

import helpers as H
from helpers import L, s2


def _bond_force(hb, r0=1.5, k=1.0):
    """HarmonicBondForce(0,1,r0,k) on the float/double positions, as int64 fixed point."""
    x = hb.posq[:2, :3].astype(np.float64)
    d = x[1] - x[0]
    r = np.linalg.norm(d)
    f0 = k * (r - r0) * d / r
    f = np.zeros((2, 3))
    f[0], f[1] = f0, -f0
    hb.force[:, :2] = np.rint(f.T * 4294967296.0).astype(np.int64)


def _plain_oracle(n, masses, precision, T, gamma, dt, kind=L.KIND_MMVT):
    system = s2.System()
    for m in masses:
        system.addParticle(m)
    integ = s2.MmvtLangevinIntegrator(T, gamma, dt, "") if kind == L.KIND_MMVT else s2.ElberLangevinIntegrator(T, gamma, dt, "")
    return H.oracle_for(integ, system, precision), integ


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
@pytest.mark.parametrize("kind", [L.KIND_MMVT, L.KIND_ELBER])
def test_single_bond(precision, kind):
    """testSingleBond (:63-106): damped harmonic oscillator vs the analytic solution, tol 0.02."""
    masses = np.array([2.0, 2.0])
    o, _ = _plain_oracle(2, masses, precision, 0.0, 0.1, 0.01, kind)
    hb = H.make_host_buffers(precision, np.array([[-1.0, 0, 0], [1.0, 0, 0]]), np.zeros((2, 3)), masses)
    rnd = np.zeros((32, 4), dtype=np.float32)
    freq = math.sqrt(1 - 0.05 * 0.05)
    for i in range(1000):
        t = o.state().time
        dist = 1.5 + 0.5 * math.exp(-0.05 * t) * math.cos(freq * t)
        speed = -0.5 * math.exp(-0.05 * t) * (0.05 * math.cos(freq * t) + freq * math.sin(freq * t))
        x = hb.posq[:2, :3].astype(np.float64) + (hb.corr[:2, :3] if hb.corr is not None else 0)
        v = hb.velm[:2, :3].astype(np.float64)
        np.testing.assert_allclose(x[0], [-0.5 * dist, 0, 0], atol=0.02)
        np.testing.assert_allclose(x[1], [0.5 * dist, 0, 0], atol=0.02)
        np.testing.assert_allclose(v[0], [-0.5 * speed, 0, 0], atol=0.02)
        np.testing.assert_allclose(v[1], [0.5 * speed, 0, 0], atol=0.02)
        _bond_force(hb)
        assert o.step(kind, hb.posq, hb.corr, hb.velm, hb.force, rnd, 0) == 0
    assert o.state().step_count == 1000


def test_single_bond_energy_conservation():
    """Second half of testSingleBond (:94-105): friction 0 conserves energy to 0.01."""
    masses = np.array([2.0, 2.0])
    o, _ = _plain_oracle(2, masses, L.DOUBLE, 0.0, 0.0, 0.01)
    hb = H.make_host_buffers(L.DOUBLE, np.array([[-1.0, 0, 0], [1.0, 0, 0]]), np.zeros((2, 3)), masses)
    rnd = np.zeros((32, 4), dtype=np.float32)
    assert o.params()[1] == 0.01          # fscale = stepSize when friction == 0

    def energy():
        x = hb.posq[:2, :3]
        r = np.linalg.norm(x[1] - x[0])
        _bond_force(hb)
        f = hb.force[:, :2].T / 4294967296.0
        # kinetic energy with the half-step shift of computeKineticEnergy (CudaSeekr2Kernels.cpp:368-370)
        v = hb.velm[:2, :3] + f * (0.5 * 0.01 / 2.0)
        return 0.5 * 1.0 * (r - 1.5) ** 2 + 0.5 * 2.0 * (v * v).sum()

    e0 = energy()
    for _ in range(1000):
        assert abs(energy() - e0) < 0.01
        _bond_force(hb)
        o.step(L.KIND_MMVT, hb.posq, hb.corr, hb.velm, hb.force, rnd, 0)


def test_temperature():
    """testTemperature (:108-142): <KE> = 3/2 N kT within 6/sqrt(10^4) (relative), harness tethers
    instead of the charged LJ particles."""
    n, temp = 8, 100.0
    masses = np.full(n, 2.0)
    o, _ = _plain_oracle(n, masses, L.MIXED, temp, 2.0, 0.01)
    pos = np.array([[(2 if i % 2 == 0 else -2), (2 if i % 4 < 2 else -2), (2 if i < 4 else -2)] for i in range(n)], dtype=np.float64)
    hb = H.make_host_buffers(L.MIXED, pos, np.zeros((n, 3)), masses)
    x0 = hb.posq.copy()
    steps = 20000
    pool = H.gaussian_pool(H.SEED_BASE, steps * 32)
    assert o.run(10000, hb.posq, hb.corr, hb.velm, hb.force, pool, 0, x0, 10.0) == 0
    ke = 0.0
    for i in range(10000):
        v = hb.velm[:n, :3]
        ke += 0.5 * (masses[:, None] * v * v).sum()
        o.run(1, hb.posq, hb.corr, hb.velm, hb.force, pool, (10000 + i) * 32, x0, 10.0)
    ke /= 10000
    expected = 0.5 * n * 3 * s2.BOLTZ * temp
    assert abs(ke - expected) / expected < 6 / math.sqrt(10000.0)


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
def test_massless_particles_do_not_move(precision):
    """testConstrainedMasslessParticles (:194-225): a massless particle keeps velocity exactly 0
    (pins the invMass == 0 skip, langevin.cu:21,81)."""
    masses = np.array([0.0, 0.0])
    o, _ = _plain_oracle(2, masses, precision, 300.0, 2.0, 0.01)
    hb = H.make_host_buffers(precision, np.array([[-1.0, 0, 0], [1.0, 0, 0]]), np.zeros((2, 3)), masses,
                             forces=np.array([[5.0, 1, 2], [3, 4, 5]]))
    before = hb.copy()
    pool = H.gaussian_pool(3, 32)
    assert o.step(L.KIND_MMVT, hb.posq, hb.corr, hb.velm, hb.force, pool, 0) == 0
    assert hb.velm[0, 0] == 0.0
    H.assert_state_equal(before, hb, 2)


def test_random_seed_reproducibility():
    """testRandomSeed (:227-284): same numbers => identical trajectories, different => different."""
    n = 8
    masses = np.full(n, 2.0)
    pos = np.array([[(2 if i % 2 == 0 else -2), (2 if i % 4 < 2 else -2), (2 if i < 4 else -2)] for i in range(n)], dtype=np.float64)
    out = []
    for seed in (5, 5, 10, 10):
        o, _ = _plain_oracle(n, masses, L.MIXED, 100.0, 2.0, 0.01)
        hb = H.make_host_buffers(L.MIXED, pos, np.zeros((n, 3)), masses)
        o.run(10, hb.posq, hb.corr, hb.velm, hb.force, H.gaussian_pool(seed, 10 * 32), 0, hb.posq.copy(), 10.0)
        out.append(hb.posq[:n, :3].copy())
    np.testing.assert_array_equal(out[0], out[1])
    np.testing.assert_array_equal(out[2], out[3])
    assert (out[0] != out[2]).all()


def test_parameters_formula():
    """CudaSeekr2Kernels.cpp:188-203."""
    o, _ = _plain_oracle(1, np.array([1.0]), L.MIXED, 300.0, 1.0, 0.002)
    vs, fs, ns = o.params()
    assert vs == math.exp(-0.002)
    assert fs == (1 - vs) / 1.0
    assert ns == math.sqrt(s2.BOLTZ * 300.0 * (1 - vs * vs))
    assert abs(s2.BOLTZ - 0.00831446261815324) < 1e-15
