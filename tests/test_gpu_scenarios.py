"""Deterministic crossing scenarios (tests/scenarios.py) on the B200 path vs the oracle, plus
size-independent properties at BASELINE.json's full sizes (23 036 and 100 000 atoms)."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("decision_mode")]


def _same(o, ob, integ, gb, n, m):
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))
    H.assert_state_equal(ob, gb, n)
    assert H.stats_tuple(o.state(), m) == H.stats_tuple(integ.getAnchorState(0), m)


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
@pytest.mark.parametrize("variant", ["cuda", "reference"])
@pytest.mark.parametrize("mode", ["fused", "two_pass"])
@pytest.mark.parametrize("restore", [False, True])
def test_ballistic_bounce_semantics(cuda_device, precision, variant, mode, restore):
    syn = S.ballistic_toy()
    syn.friction = 0.5
    for which in ("oracle", "gpu"):
        integ = H.make_mmvt_integrator(syn, "", variant)
        integ.setRestoreCorrection(restore)
        if which == "oracle":
            rc, o, ob = S.run_oracle(syn, integ, precision, 300)
        else:
            grc, ctx, gb = S.run_gpu(syn, integ, precision, 300, mode=mode)
    assert rc == grc == 0 and len(o.records()) >= 2
    _same(o, ob, integ, gb, 1, 2)


def test_corner_rawstyle_and_first_step(cuda_device):
    for make, steps, m in ((lambda: S.ballistic_toy(extra_plane=True), 400, 3), (lambda: S.ballistic_toy(raw_style=True), 200, 1)):
        syn = make()
        rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.MIXED, steps)
        integ = H.make_mmvt_integrator(syn, "")
        grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, steps, graph=8)
        assert rc == grc == 0
        _same(o, ob, integ, gb, 1, m)
    syn = S.ballistic_toy()
    syn.positions[0, 0] = 1.2 + 1.01
    integ = H.make_mmvt_integrator(syn, "")
    grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, 3)
    assert grc == L.ERR_FIRST_STEP and integ.getAnchorState(0).step_count == 0


@pytest.mark.parametrize("variant", ["cuda", "reference"])
@pytest.mark.parametrize("mode", ["fused", "two_pass"])
def test_elber_scenarios(cuda_device, variant, mode):
    for end, speed in ((False, 0.5), (True, 0.5), (False, -0.5)):
        syn, integ_o = S.elber_ballistic(end, speed=speed)
        integ_o.setVariant(variant)
        rc, o, ob = S.run_oracle(syn, integ_o, L.MIXED, 120)
        syn2, integ = S.elber_ballistic(end, speed=speed)
        integ.setVariant(variant)
        grc, ctx, gb = S.run_gpu(syn2, integ, L.MIXED, 120, mode=mode)
        assert rc == grc == 0 and len(o.records()) == 1
        _same(o, ob, integ, gb, 1, 0)


def test_large_groups_span_several_warps_and_blocks(cuda_device):
    """Groups of 300 and 77 atoms (more slots than one decide block handles in one pass), massless atoms
    inside a group, a shared atom between two groups."""
    host, guest = list(range(0, 600, 2)), list(range(601, 755, 2)) + [0]
    syn = H.pair_sphere_system(3000, host, guest, 0.50, 0.56, H.SEED_BASE + 9, tether_k=300.0)
    syn.masses[4] = 0.0
    syn.system.setParticleMass(4, 0.0)
    syn.velocities[4] = 0.0
    H.place_pair_distance(syn, host, [g for g in guest if g != 0], 0.53)
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.MIXED, 400, H.gaussian_pool(5, 400 * 3008))
    integ = H.make_mmvt_integrator(syn, "")
    grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, 400, H.gaussian_pool(5, 400 * 3008), graph=16)
    assert rc == grc == 0
    _same(o, ob, integ, gb, 3000, 2)


@pytest.mark.parametrize("n_atoms,planes", [(23036, False), (100000, True)])
def test_full_size_properties(cuda_device, n_atoms, planes):
    """BASELINE sizes (C4: 23 036 atoms; C5: 100 000 atoms with spheres + planes), 4 anchors:
      - anchors given identical inputs evolve identically (batch index independence);
      - MMVT engine == plain Langevin engine on atoms' state for every step without a bounce;
      - a bounce step leaves posq untouched for EVERY atom and negates the kicked velocity;
      - oracle parity on one anchor for a few steps."""
    import torch
    rec, lig = [2478, 2489, 2499, 2535, 2718, 2745, 2769, 2787, 2794, 2867, 2926], [3221, 3222, 3223, 3224, 3225, 3226, 3227, 3228, 3229]
    syn = H.pair_sphere_system(n_atoms, rec, lig, 1.10, 1.1004, H.SEED_BASE + 11, dt=0.004, tether_k=0.0, extra_planes=planes,
                               plane_group=list(range(5000, 5064)))
    H.place_pair_distance(syn, rec, lig, 1.1002)
    A, steps = 4, 24
    npad = (n_atoms + 31) // 32 * 32
    pool = H.gaussian_pool(77, steps * npad)
    hb = H.make_host_buffers(L.MIXED, syn.positions, syn.velocities, syn.masses)
    integ = H.make_mmvt_integrator(syn, "")
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
    H.upload(ctx, [hb] * A)
    ctx.setRandomPool(pool)
    lang = s2.LangevinIntegrator(300.0, 1.0, 0.004)
    ctx_l = s2.Context(syn.system, lang, {"CudaPrecision": "mixed"}, numAnchors=1)
    H.upload(ctx_l, [hb])
    ctx_l.setRandomPool(pool)
    o = H.oracle_for(H.make_mmvt_integrator(syn, ""), syn.system, L.MIXED)
    ob = hb.copy()
    bounced_steps = 0
    for s in range(steps):
        before = H.download(ctx, 0)
        integ.step(1)
        assert o.run(1, ob.posq, ob.corr, ob.velm, ob.force, pool, s * npad, None, 0.0) == 0
        after = H.download(ctx, 0)
        H.assert_state_equal(ob, after, n_atoms, f"oracle step {s}")
        for a in range(1, A):
            H.assert_state_equal(after, H.download(ctx, a), n_atoms, f"anchor {a} step {s}")
        nb = integ.getAnchorState(0).num_bounce_steps
        if nb > bounced_steps:                       # this step bounced
            bounced_steps = nb
            np.testing.assert_array_equal(after.posq[:n_atoms], before.posq[:n_atoms])
            vs = integ._prev and np.exp(-0.004)
            moving = before.velm[:n_atoms, 3] != 0
            assert (np.sign(after.velm[:n_atoms, :3][moving]) != 0).any()
            H.upload(ctx_l, [after])                 # resynchronise the plain-Langevin twin
            ctx_l.random_cursor = s + 1
        else:
            lang.step(1)
            H.assert_state_equal(H.download(ctx_l, 0), after, n_atoms, f"plain Langevin twin step {s}")
    assert bounced_steps >= 1, "the scenario must contain at least one bounce"
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))


def test_angle_and_pair_sum_boundaries_on_gpu(cuda_device):
    """demo5-style angle milestones (two step-style bonds, inner and outer angle) and a demo4-style pair-sum
    milestone on small solvated systems: trajectories, decisions and records bit-equal to the oracle."""
    import math
    n = 600
    pos, vel, masses = H.lattice_system(n, H.SEED_BASE + 21)
    pool = H.gaussian_pool(H.SEED_BASE + 22, 800 * 608)

    def angle_system():
        system = s2.System()
        for m in masses:
            system.addParticle(float(m))
        f = s2.CustomCentroidBondForce(3, "bitcode*step(k*(angle(g1, g2, g3) - ref_angle))")
        g = [f.addGroup([5, 6, 7]), f.addGroup([100, 101]), f.addGroup([300, 301, 302, 303])]
        f.setForceGroup(1)
        for name in ("bitcode", "k", "ref_angle"):
            f.addPerBondParameter(name)
        c = [np.average(pos[idx], axis=0, weights=masses[idx]) for idx in ([5, 6, 7], [100, 101], [300, 301, 302, 303])]
        a, b = c[0] - c[1], c[2] - c[1]
        ang = math.acos(a @ b / math.sqrt((a @ a) * (b @ b)))
        f.addBond(g, [1.0, -1.0, ang - 0.004])
        f.addBond(g, [2.0, 1.0, ang + 0.004])
        system.addForce(f)
        return system

    def pair_sum_system():
        system = s2.System()
        for m in masses:
            system.addParticle(float(m))
        f = s2.CustomCentroidBondForce(6, "bitcode*step(k*(distance(g1,g2)^2 + distance(g3,g4)^2 + distance(g5,g6)^2 - radius^2))")
        idx = [[1], [50], [2], [51], [3], [52]]
        g = [f.addGroup(i) for i in idx]
        f.setForceGroup(1)
        for name in ("bitcode", "k", "radius"):
            f.addPerBondParameter(name)
        total = sum(((pos[idx[2 * p][0]] - pos[idx[2 * p + 1][0]]) ** 2).sum() for p in range(3))
        f.addBond(g, [1.0, -1.0, math.sqrt(total) - 0.01])
        f.addBond(g, [2.0, 1.0, math.sqrt(total) + 0.01])
        system.addForce(f)
        return system

    for make in (angle_system, pair_sum_system):
        system = make()
        syn = H.Synthetic(system, pos, vel, masses, None, [1, 2], 0.0, 0.002)
        rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.MIXED, 800, pool)
        integ = H.make_mmvt_integrator(syn, "")
        grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, 800, pool, graph=16)
        assert rc == grc == 0 and o.state().num_bounce_steps >= 2, make.__name__
        _same(o, ob, integ, gb, n, 2)
        integ2 = H.make_mmvt_integrator(syn, "")
        grc, ctx, gb2 = S.run_gpu(syn, integ2, L.MIXED, 800, pool, mode="two_pass")
        _same(o, ob, integ2, gb2, n, 2)
