"""Pins the oracle's crossing bookkeeping against the text of the reference
(CudaSeekr2Kernels.cpp:205-365 MMVT, :517-594 Elber; ReferenceSeekr2Kernels.cpp for VARIANT_REFERENCE)
on deterministic ballistic scenarios.  The reference ships no test that exercises a crossing, so these
expectations are derived by hand from the cited lines (parity unpinned otherwise)."""
import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2


def _records(o):
    return [(r.kind, r.index, r.counter, r.num_surfaces, r.step, round(r.time, 9)) for r in o.records()]


def test_ballistic_bounces_between_two_spheres():
    """0.005 nm per step from r = 0.9: reaches r = 1.0 on the step whose NEW position is >= 1.0; the
    bounce restores the previous position and reverses the velocity, then 0.2 nm inwards, and so on."""
    syn = S.ballistic_toy()
    integ = H.make_mmvt_integrator(syn, "")
    rc, o, hb = S.run_oracle(syn, integ, L.DOUBLE, 200)
    assert rc == 0
    recs = _records(o)
    assert [r[1] for r in recs] == [1, 0, 1, 0, 1]                 # outer, inner, outer, ...
    assert [r[2] for r in recs] == [0, 1, 2, 3, 4]                 # bounceCounter on the line (:281,:309)
    assert all(r[3] == 1 for r in recs)
    steps = [r[4] for r in recs]
    assert steps[0] in (19, 20) and all(b - a in (39, 40, 41) for a, b in zip(steps, steps[1:]))
    # the logged time is the context time BEFORE the step's increment (:281 vs :362)
    assert all(abs(t - s * 0.01) < 1e-9 for (_, _, _, _, s, t) in recs)
    st = o.state()
    assert st.bounce_counter == 5 and st.previous_milestone == 1 and st.num_bounce_steps == 5
    # CUDA variant: N_alpha_beta skips the very first crossing (:294-296)
    assert list(st.N_alpha_beta[:2]) == [2, 2]
    nij = lambda i, j: st.Nij_alpha[i * L.MAX_MILESTONES + j]
    assert (nij(1, 0), nij(0, 1), nij(0, 0), nij(1, 1)) == (2, 2, 0, 0)
    # incubation time accumulates stepSize every step and is reset at a crossing to a new milestone
    # (:297-305,364): each leg takes (steps between crossings) * dt
    assert abs(st.Ri_alpha[1] - sum((steps[i + 1] - steps[i]) * 0.01 for i in (0, 2))) < 1e-9
    assert abs(st.Ri_alpha[0] - sum((steps[i + 1] - steps[i]) * 0.01 for i in (1, 3))) < 1e-9
    assert abs(st.T_alpha - (steps[-1] - steps[0]) * 0.01) < 1e-9
    assert abs(st.time - 2.0) < 1e-9 and st.step_count == 200


def test_reference_variant_counts_first_crossing_and_restores_start_velocity():
    syn = S.ballistic_toy()
    integ = H.make_mmvt_integrator(syn, "", variant="reference")
    rc, o, hb = S.run_oracle(syn, integ, L.DOUBLE, 200)
    assert list(o.state().N_alpha_beta[:2]) == [2, 3]              # unconditional increment (Reference :290)


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
def test_bounce_restores_position_and_negates_kicked_velocity(precision):
    """One step that crosses: posq == old posq, velm == -(post-kick velm) (langevin.cu:167-179), and in
    mixed precision the correction keeps Part2's NEW value unless FLAG_RESTORE_CORRECTION."""
    for restore in (False, True):
        syn = S.ballistic_toy(speed=0.5, dt=0.01)
        syn.positions[0, 0] = 1.2 + 0.998                          # 0.005 nm from the outer sphere: crosses at once
        syn.friction = 1.0                                          # so that v(kick) != v(start)
        integ = H.make_mmvt_integrator(syn, "")
        integ.setRestoreCorrection(restore)
        before = H.make_host_buffers(precision, syn.positions, syn.velocities, syn.masses)
        rc, o, hb = S.run_oracle(syn, integ, precision, 1)
        assert rc == 0 and len(o.records()) == 1
        np.testing.assert_array_equal(hb.posq[:1], before.posq[:1])
        vs = o.params()[0]
        mixed = H.mixed_dtype(precision)
        expect = -(mixed(vs) * before.velm[0, 0])
        assert hb.velm[0, 0] == expect and hb.velm[0, 3] == before.velm[0, 3]
        if precision == L.MIXED:
            if restore:
                np.testing.assert_array_equal(hb.corr[:1], before.corr[:1])
            else:
                X = np.float64(before.posq[0, 0]) + np.float64(before.corr[0, 0]) + np.float64(0.01) * (-np.float64(expect))
                assert hb.corr[0, 0] == np.float32(X - np.float64(np.float32(X)))
        # reference variant negates the velocity at the START of the step instead
        integ2 = H.make_mmvt_integrator(syn, "", variant="reference")
        rc, o2, hb2 = S.run_oracle(syn, integ2, precision, 1)
        assert hb2.velm[0, 0] == -before.velm[0, 0]


def test_corner_bounce_logs_both_surfaces():
    """Two boundaries crossed in one step: two lines at the same time, counter advances twice,
    num_surfaces == 2 (state would not be saved, :264-283)."""
    syn = S.ballistic_toy(extra_plane=True)
    integ = H.make_mmvt_integrator(syn, "")
    rc, o, hb = S.run_oracle(syn, integ, L.MIXED, 30)
    recs = _records(o)
    assert len(recs) == 2
    (k0, i0, c0, n0, s0, t0), (k1, i1, c1, n1, s1, t1) = recs
    assert (i0, i1) == (1, 2) and (c0, c1) == (0, 1) and n0 == n1 == 2 and s0 == s1 and t0 == t1
    st = o.state()
    assert st.bounce_counter == 2 and st.num_bounce_steps == 1 and st.previous_milestone == 2
    assert st.Nij_alpha[1 * L.MAX_MILESTONES + 2] == 1            # the second line sees previous == 1


def test_first_step_behind_boundary_is_an_error():
    syn = S.ballistic_toy()
    syn.positions[0, 0] = 1.2 + 1.01
    rc, o, hb = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.MIXED, 3)
    assert rc == L.ERR_FIRST_STEP and o.state().step_count == 0


def test_boundary_value_counts_as_crossed():
    """step(0) == 1: a centroid exactly ON the surface is a crossing.  Uses binary-exact
    numbers (centre 1.25, radius 1, position 2.25) so that u is exactly 0."""
    system = s2.System()
    system.addParticle(1.0)
    f = s2.CustomCentroidBondForce(1, "bitcode*step(k*((x1-centerx)^2 + (y1-centery)^2 + (z1-centerz)^2 - radius^2))")
    g = f.addGroup([0])
    f.setForceGroup(1)
    for name in ("bitcode", "k", "centerx", "centery", "centerz", "radius"):
        f.addPerBondParameter(name)
    f.addBond([g], [2.0, 1.0, 1.25, 1.25, 1.25, 1.0])
    system.addForce(f)
    integ = s2.MmvtLangevinIntegrator(0.0, 0.0, 0.01, "")
    integ.addMilestoneGroup(1); integ.addMilestoneGroup(2)
    o = H.oracle_for(integ, system, L.DOUBLE)
    for x, expect in ((2.25, 2.0), (2.25 - 2.0 ** -30, 0.0), (2.25 + 2.0 ** -30, 2.0)):
        hb = H.make_host_buffers(L.DOUBLE, np.array([[x, 1.25, 1.25]]), np.zeros((1, 3)), np.array([1.0]))
        assert o.group_value(hb.posq, hb.corr, 1) == expect
        assert o.group_value(hb.posq, hb.corr, 2) == 0.0          # other force groups are not looked at


def test_old_style_expression_bounces_without_logging():
    """Non-step() expressions give value > 0 but (int) value == 0: bounce, no line, counter unchanged
    (CudaSeekr2Kernels.cpp:262-281)."""
    syn = S.ballistic_toy(raw_style=True)
    rc, o, hb = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.DOUBLE, 60)
    st = o.state()
    assert rc == 0 and st.num_bounce_steps >= 1 and len(o.records()) == 0 and st.bounce_counter == 0


def test_elber_source_reset_then_destination():
    """endOnSrcMilestone = false: crossing the source resets the clock to 0 and marks it crossed; the
    destination crossing then ends the run with a plain label; nothing is logged afterwards."""
    syn, integ = S.elber_ballistic(False)
    integ.setCrossingCounter(3)
    rc, o, hb = S.run_oracle(syn, integ, L.DOUBLE, 80)
    recs = _records(o)
    assert len(recs) == 1
    kind, index, counter, nsurf, step, time = recs[0]
    assert (kind, index, counter, nsurf) == (L.REC_ELBER_DEST, 1, 3, 1)
    st = o.state()
    assert st.crossed_src == 1 and st.end_simulation == 1 and st.crossing_counter == 4
    # source crossed around step 10 (0.95 -> 1.0), clock reset; destination 0.1 nm later = 20 more steps
    assert 0.15 < time < 0.25 and step in range(28, 33)
    assert st.step_count == 80                                     # integration continues after the end (:524)


def test_elber_end_on_source():
    syn, integ = S.elber_ballistic(True)
    rc, o, hb = S.run_oracle(syn, integ, L.DOUBLE, 80)
    recs = _records(o)
    assert len(recs) == 1 and recs[0][0] == L.REC_ELBER_SRC and recs[0][1] == 0
    assert o.state().end_simulation == 1 and o.state().crossed_src == 0
    assert abs(recs[0][5] - recs[0][4] * 0.01) < 1e-9             # no clock reset in this mode


def test_elber_destination_without_source_gets_asterisk():
    syn, integ = S.elber_ballistic(False, speed=-0.5)              # moves inwards: hits r = 0.9 first
    rc, o, hb = S.run_oracle(syn, integ, L.DOUBLE, 40)
    recs = _records(o)
    assert len(recs) == 1 and recs[0][0] == L.REC_ELBER_DEST_STAR and recs[0][1] == 0


def test_elber_no_crossing_at_time_zero():
    """CUDA platform: a value change is ignored while context time == 0 (:533,:558); the Reference
    platform has no such guard (:472,:498).  Baselines are sampled AFTER the first step (:528-531)."""
    for variant, expect in (("cuda", 0), ("reference", 0)):
        syn, integ = S.elber_ballistic(True, start=0.999)          # crosses the source on step 1
        integ.setVariant(variant)
        rc, o, hb = S.run_oracle(syn, integ, L.DOUBLE, 1)
        # first sample IS the post-step value, so no change can be seen on step 1 in either variant
        assert len(o.records()) == expect


def test_log_files_append_and_restart(tmp_path):
    """Existing file: no second header, lines appended, counter seeded by setBounceCounter (:151,:158-171,:258)."""
    path = str(tmp_path / "mmvt.txt")
    syn = S.ballistic_toy()
    integ = H.make_mmvt_integrator(syn, "")
    rc, o, hb = S.run_oracle(syn, integ, L.MIXED, 100)
    o.write_log(path, syn.milestone_groups)
    first = open(path).read()
    assert first.startswith('#"Bounced boundary ID","bounce index","total time (ps)"\n2,0,0.')
    integ2 = H.make_mmvt_integrator(syn, "")
    integ2.setBounceCounter(o.state().bounce_counter)
    rc, o2, hb2 = S.run_oracle(syn, integ2, L.MIXED, 100)
    o2.write_log(path, syn.milestone_groups)
    text = open(path).read()
    assert text.count("#") == 1 and text.startswith(first)
    second = text[len(first):].splitlines()
    assert second[0].split(",")[1] == str(o.state().bounce_counter)
    stats = str(tmp_path / "stats.txt")
    o.write_statistics(stats, syn.milestone_groups)
    lines = open(stats).read().splitlines()
    assert lines[0].startswith("N_alpha_1: ") and lines[-1].startswith("T_alpha: ") and len(lines) == 2 + 4 + 2 + 1


def _three_particle_system(expr, ngroups, params, names, positions, style_groups=1):
    system = s2.System()
    for _ in range(len(positions)):
        system.addParticle(12.0)
    f = s2.CustomCentroidBondForce(ngroups, expr)
    gs = [f.addGroup([i]) for i in range(ngroups)]
    f.setForceGroup(1)
    for n in names:
        f.addPerBondParameter(n)
    f.addBond(gs, params)
    system.addForce(f)
    integ = s2.MmvtLangevinIntegrator(0.0, 0.0, 0.01, "")
    integ.addMilestoneGroup(1)
    return system, integ


def test_angle_boundary_matches_acos_definition():
    """demo5:24 `angle(g1,g2,g3) - ref_angle`, evaluated through the cosine: same decisions as the acos form."""
    import math
    rng = np.random.default_rng(3)
    for trial in range(200):
        pos = rng.uniform(-1, 1, (3, 3))
        ref = rng.uniform(0.2, 3.0)
        a, b = pos[0] - pos[1], pos[2] - pos[1]
        ang = math.acos(max(-1.0, min(1.0, float(a @ b / math.sqrt((a @ a) * (b @ b))))))
        if abs(ang - ref) < 1e-6:
            continue
        for ksign in (1.0, -1.0):
            system, integ = _three_particle_system("bitcode*step(k*(angle(g1,g2,g3)-ref_angle))", 3, [1.0, ksign, ref],
                                                   ("bitcode", "k", "ref_angle"), pos)
            o = H.oracle_for(integ, system, L.DOUBLE)
            hb = H.make_host_buffers(L.DOUBLE, pos, np.zeros((3, 3)), np.full(3, 12.0))
            expect = 1.0 if ksign * (ang - ref) >= 0 else 0.0
            assert o.group_value(hb.posq, hb.corr, 1) == expect


def test_pair_sum_boundary():
    """demo4:27: distance(g1,g2)^2 + distance(g3,g4)^2 + distance(g5,g6)^2 - radius^2 (old non-step style)."""
    pos = np.array([[0.0, 0, 0], [0.3, 0, 0], [1.0, 0, 0], [1.0, 0.4, 0], [2.0, 0, 0], [2.0, 0, 0.5]])
    total = 0.09 + 0.16 + 0.25
    for radius, ksign in ((0.6, -1.0), (0.8, -1.0), (0.6, 1.0), (0.8, 1.0)):
        system, integ = _three_particle_system("-k*(distance(g1,g2)^2 + distance(g3,g4)^2 + distance(g5,g6)^2 - radius^2)" if ksign < 0
                                               else "k*(distance(g1,g2)^2 + distance(g3,g4)^2 + distance(g5,g6)^2 - radius^2)",
                                               6, [1.0e-9, radius], ("k", "radius"), pos)
        o = H.oracle_for(integ, system, L.DOUBLE)
        hb = H.make_host_buffers(L.DOUBLE, pos, np.zeros((6, 3)), np.full(6, 12.0))
        value = o.group_value(hb.posq, hb.corr, 1)
        assert abs(value - ksign * 1.0e-9 * (total - radius ** 2)) < 1e-20
