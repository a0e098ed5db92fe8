"""The reference's platform tests restated on the B200 path through the harness
(TestCudaMmvtLangevinIntegrator.cpp:63-284; run for single / mixed / double like cuda/tests/CMakeLists.txt:20-22)."""
import math

import numpy as np
import pytest

import helpers as H
from helpers import L, s2

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("decision_mode")]


@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
@pytest.mark.parametrize("cls", [s2.MmvtLangevinIntegrator, s2.ElberLangevinIntegrator])
def test_single_bond(cuda_device, precision, cls):
    """testSingleBond (:63-106): damped harmonic oscillator vs the analytic solution, tol 0.02; then
    friction 0 conserves energy to 0.01."""
    system = s2.System()
    system.addParticle(2.0)
    system.addParticle(2.0)
    integrator = cls(0, 0.1, 0.01, "")
    bonds = s2.HarmonicBondForce()
    bonds.addBond(0, 1, 1.5, 1)
    system.addForce(bonds)
    context = s2.Context(system, integrator, {"CudaPrecision": precision})
    positions = np.array([[-1.0, 0, 0], [1.0, 0, 0]])
    context.setPositions(positions)
    freq = math.sqrt(1 - 0.05 * 0.05)
    for i in range(1000):
        state = context.getState(getPositions=True, getVelocities=True)
        time = state.getTime()
        assert abs(time - i * 0.01) < 1e-9
        dist = 1.5 + 0.5 * math.exp(-0.05 * time) * math.cos(freq * time)
        speed = -0.5 * math.exp(-0.05 * time) * (0.05 * math.cos(freq * time) + freq * math.sin(freq * time))
        np.testing.assert_allclose(state.getPositions()[0], [-0.5 * dist, 0, 0], atol=0.02)
        np.testing.assert_allclose(state.getPositions()[1], [0.5 * dist, 0, 0], atol=0.02)
        np.testing.assert_allclose(state.getVelocities()[0], [-0.5 * speed, 0, 0], atol=0.02)
        np.testing.assert_allclose(state.getVelocities()[1], [0.5 * speed, 0, 0], atol=0.02)
        integrator.step(1)
    integrator.setFriction(0.0)
    context.setPositions(positions)
    context.setVelocities(np.zeros((2, 3)))

    def energy():
        st = context.getState(getPositions=True, getVelocities=True)
        x, v = st.getPositions(), st.getVelocities()
        r = np.linalg.norm(x[1] - x[0])
        f0 = (r - 1.5) * (x[1] - x[0]) / r
        vs = v + np.stack([f0, -f0]) * (0.5 * 0.01 / 2.0)      # computeKineticEnergy's half-step shift
        return 0.5 * (r - 1.5) ** 2 + 0.5 * 2.0 * (vs * vs).sum()

    e0 = energy()
    for i in range(1000):
        assert abs(energy() - e0) < 0.01
        integrator.step(1)


def test_temperature(cuda_device):
    """testTemperature (:108-142) with harness tethers and the in-kernel Philox generator:
    <KE> = 3/2 N kT within 6/sqrt(10^4)."""
    n, temp = 8, 100.0
    system = s2.System()
    for _ in range(n):
        system.addParticle(2.0)
    pos = np.array([[(2 if i % 2 == 0 else -2), (2 if i % 4 < 2 else -2), (2 if i < 4 else -2)] for i in range(n)], dtype=np.float64)
    system.addForce(s2.TetherForce(10.0, pos))
    integrator = s2.MmvtLangevinIntegrator(temp, 2.0, 0.01, "")
    integrator.setRandomNumberSeed(7)
    integrator.setUseGraph(True, 100)
    context = s2.Context(system, integrator, {"CudaPrecision": "mixed"}, numAnchors=16)
    context.setPositions(pos)
    integrator.step(10000)
    ke = 0.0
    for i in range(100):
        ke += context.getState(getEnergy=True).getKineticEnergy().mean()
        integrator.step(100)
    ke /= 100
    expected = 0.5 * n * 3 * s2.BOLTZ * temp
    assert abs(ke - expected) / expected < 6 / math.sqrt(10000.0)


def test_massless_particles_and_random_seed(cuda_device):
    """testConstrainedMasslessParticles (:194-225): massless particles keep velocity exactly 0.
    testRandomSeed (:227-284): same seed => same trajectory, different seed => different."""
    system = s2.System()
    system.addParticle(0.0)
    system.addParticle(0.0)
    integrator = s2.MmvtLangevinIntegrator(300.0, 2.0, 0.01, "")
    context = s2.Context(system, integrator, {"CudaPrecision": "mixed"})
    context.setPositions(np.array([[-1.0, 0, 0], [1.0, 0, 0]]))
    integrator.step(1)
    assert context.getState(getVelocities=True).getVelocities()[0][0] == 0.0
    n = 8
    pos = np.array([[(2 if i % 2 == 0 else -2), (2 if i % 4 < 2 else -2), (2 if i < 4 else -2)] for i in range(n)], dtype=np.float64)
    out = []
    for seed in (5, 5, 10, 10):
        system = s2.System()
        for _ in range(n):
            system.addParticle(2.0)
        system.addForce(s2.TetherForce(10.0, pos))
        integrator = s2.MmvtLangevinIntegrator(100.0, 2.0, 0.01, "")
        integrator.setRandomNumberSeed(seed)
        context = s2.Context(system, integrator, {"CudaPrecision": "mixed"})
        context.setPositions(pos)
        context.setVelocities(np.zeros((n, 3)))
        integrator.step(10)
        out.append(context.getState(getPositions=True).getPositions())
    np.testing.assert_allclose(out[0], out[1], atol=1e-6)
    np.testing.assert_allclose(out[2], out[3], atol=1e-6)
    assert (out[0] != out[2]).all()
