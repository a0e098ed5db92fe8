"""Randomised differential test: random boundary sets (centroid-pair spheres, fixed-centre spheres, axis planes; step and
raw style; explicit weights; groups of 1 .. 140 atoms, some spanning several warps, some sharing atoms), random precision,
scheme and variant -- the fused step in BOTH decision modes and the two-pass form must equal the CPU oracle bit for bit
(state, crossing records, statistics).  Thresholds are placed a hair from the starting values so that every case bounces."""
import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2

pytestmark = pytest.mark.gpu


def _random_case(seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(200, 2600))
    pos, vel, masses = H.lattice_system(n, 1000 + seed)
    if rng.random() < 0.3:                                   # a few immobile particles
        for i in rng.choice(n, 3, replace=False):
            masses[i] = 0.0
            vel[i] = 0.0
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    bit = 1.0
    groups_used = 0
    labels = []

    def centroid(idx, w=None):
        ww = masses[idx] if w is None else np.asarray(w)
        return (pos[idx] * ww[:, None]).sum(0) / ww.sum()

    def pick(size):
        idx = rng.choice(n, size, replace=False)
        idx = idx[masses[idx] > 0] if rng.random() < 0.5 else idx
        return [int(i) for i in (idx if len(idx) else [int(np.argmax(masses))])]

    n_forces = int(rng.integers(1, 4))
    for _ in range(n_forces):
        kind = rng.choice(["pair", "fixed", "plane"])
        raw = rng.random() < 0.15
        margin = float(rng.uniform(0.0005, 0.003))
        sizes = [int(rng.choice([1, 3, 11, 33, 70, 140])) for _ in range(2)]
        if groups_used + 2 > 14 or bit > 8:
            break
        if kind == "pair":
            f = s2.CustomCentroidBondForce(2, "k*(distance(g1,g2)^2-radius^2)" if raw else "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
            a, b = pick(sizes[0]), pick(sizes[1])
            wa = [float(x) for x in rng.uniform(0.5, 2.0, len(a))] if rng.random() < 0.4 else None
            ga = f.addGroup(a, wa) if wa else f.addGroup(a)
            gb = f.addGroup(b)
            d = float(np.linalg.norm(centroid(a, wa) - centroid(b)))
            params_in, params_out = ([-1.0, max(d - margin, 1e-4)], [1.0, d + margin])
            names = ("k", "radius") if raw else ("bitcode", "k", "radius")
            for p in names:
                f.addPerBondParameter(p)
            if raw:
                f.addBond([ga, gb], params_out)
            else:
                f.addBond([ga, gb], [bit] + params_in); bit *= 2
                f.addBond([ga, gb], [bit] + params_out); bit *= 2
            groups_used += 2
        elif kind == "fixed":
            f = s2.CustomCentroidBondForce(1, "k*((x1-cx)^2+(y1-cy)^2+(z1-cz)^2-radius^2)" if raw else "bitcode*step(k*((x1-cx)^2+(y1-cy)^2+(z1-cz)^2-radius^2))")
            a = pick(sizes[0])
            ga = f.addGroup(a)
            c = centroid(a) + rng.uniform(-0.3, 0.3, 3)
            d = float(np.linalg.norm(centroid(a) - c))
            names = ("k", "cx", "cy", "cz", "radius") if raw else ("bitcode", "k", "cx", "cy", "cz", "radius")
            for p in names:
                f.addPerBondParameter(p)
            tail = [1.0, float(c[0]), float(c[1]), float(c[2]), d + margin]
            f.addBond([ga], tail if raw else [bit] + tail)
            if not raw:
                bit *= 2
            groups_used += 1
        else:
            axis = "xyz"[int(rng.integers(0, 3))]
            f = s2.CustomCentroidBondForce(1, f"k*({axis}1-bound)" if raw else f"bitcode*step(k*({axis}1-bound))")
            a = pick(sizes[0])
            ga = f.addGroup(a)
            c = centroid(a)["xyz".index(axis)]
            names = ("k", "bound") if raw else ("bitcode", "k", "bound")
            for p in names:
                f.addPerBondParameter(p)
            sign = float(rng.choice([-1.0, 1.0]))
            tail = [sign, float(c + sign * margin)]
            f.addBond([ga], tail if raw else [bit] + tail)
            if not raw:
                bit *= 2
            groups_used += 1
        f.setForceGroup(1)
        system.addForce(f)
    num_bits = max(1, int(np.log2(bit))) if bit > 1 else 1
    labels = list(range(1, num_bits + 1))
    syn = H.Synthetic(system, pos, vel, masses, None, labels, 0.0, 0.002)
    precision = [L.SINGLE, L.MIXED, L.MIXED, L.DOUBLE][int(rng.integers(0, 4))]
    middle = bool(rng.random() < 0.4)
    variant = "reference" if rng.random() < 0.25 else "cuda"
    steps = int(rng.integers(60, 200))
    return syn, precision, middle, variant, steps, n


@pytest.mark.parametrize("seed", range(48))
def test_random_boundary_sets_match_the_oracle(cuda_device, seed):
    syn, precision, middle, variant, steps, n = _random_case(seed)
    npad = (n + 31) // 32 * 32
    pool = H.gaussian_pool(7000 + seed, steps * npad)
    m = len(syn.milestone_groups)
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, "", variant, middle=middle), precision, steps, pool)
    for mode, decision in (("fused", "local"), ("fused", "handoff"), ("two_pass", "auto")):
        integ = H.make_mmvt_integrator(syn, "", variant, middle=middle)
        if decision != "auto":
            try:
                integ.setDecisionMode(decision)
                grc, ctx, gb = S.run_gpu(syn, integ, precision, steps, pool, mode=mode, graph=8 if seed % 2 else 0)
            except s2.Seekr2Error as e:                 # more than 256 padded slots: LOCAL is refused, the hand-off covers it
                assert decision == "local" and "group slots" in str(e)
                continue
        else:
            grc, ctx, gb = S.run_gpu(syn, integ, precision, steps, pool, mode=mode)
        assert rc == grc, (seed, mode, decision)
        if rc != 0:
            continue                                    # e.g. a raw-style term positive at step 0: both sides refuse the start
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0)), (seed, mode, decision)
        H.assert_state_equal(ob, gb, n, f"seed {seed} {mode} {decision}")
        assert H.stats_tuple(o.state(), m) == H.stats_tuple(integ.getAnchorState(0), m), (seed, mode, decision)


def test_fuzz_cases_do_bounce(cuda_device):
    """The generator is only useful if its cases cross boundaries: most of them must."""
    bounced = 0
    for seed in range(16):
        syn, precision, middle, variant, steps, n = _random_case(seed)
        pool = H.gaussian_pool(7000 + seed, steps * ((n + 31) // 32 * 32))
        rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, "", variant, middle=middle), precision, steps, pool)
        bounced += rc == 0 and o.state().num_bounce_steps > 0
    assert bounced >= 8, bounced


def _random_elber_case(seed):
    """Thermal solvated system, 2-4 milestone forces in their own force groups (one source, the rest destinations), each a
    centroid-pair sphere a hair outside / inside the starting distance: the run crosses the source (clock reset or end)
    and usually ends on a destination within a few hundred steps."""
    rng = np.random.default_rng(5000 + seed)
    n = int(rng.integers(300, 2200))
    pos, vel, masses = H.lattice_system(n, 3000 + seed)
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    a = [int(i) for i in rng.choice(n, int(rng.choice([1, 5, 20, 100])), replace=False)]      # 100: more than 256 padded slots
    b = [int(i) for i in rng.choice(n, int(rng.choice([1, 9, 30])), replace=False)]           #      in total -> hand-off only
    ca = (pos[a] * masses[a][:, None]).sum(0) / masses[a].sum()
    cb = (pos[b] * masses[b][:, None]).sum(0) / masses[b].sum()
    d = float(np.linalg.norm(ca - cb))
    k = int(rng.integers(2, 5))
    margins = sorted(float(x) for x in rng.uniform(0.0003, 0.004, k))
    for g in range(k):
        f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
        ga, gb = f.addGroup(a), f.addGroup(b)
        for p in ("bitcode", "k", "radius"):
            f.addPerBondParameter(p)
        sign = 1.0 if g % 2 == 0 else -1.0                     # alternate outside / inside surfaces
        f.addBond([ga, gb], [1.0, sign, max(d + sign * margins[g], 1e-4)])
        f.setForceGroup(g + 1)
        system.addForce(f)
    syn = H.Synthetic(system, pos, vel, masses, None, [], 0.0, 0.002)
    middle = bool(rng.random() < 0.5)
    end_on_src = bool(rng.random() < 0.3)
    precision = [L.SINGLE, L.MIXED, L.DOUBLE][int(rng.integers(0, 3))]
    variant = "reference" if rng.random() < 0.25 else "cuda"
    steps = int(rng.integers(80, 260))

    def make():
        integ = (s2.ElberLangevinMiddleIntegrator if middle else s2.ElberLangevinIntegrator)(300.0, 1.0, 0.002, "")
        integ.addSrcMilestoneGroup(1)
        for g in range(2, k + 1):
            integ.addDestMilestoneGroup(g)
        integ.setEndOnSrcMilestone(end_on_src)
        integ.setVariant(variant)
        return integ
    return syn, make, precision, steps, n


@pytest.mark.parametrize("seed", range(24))
def test_random_elber_cases_match_the_oracle(cuda_device, seed):
    syn, make, precision, steps, n = _random_elber_case(seed)
    pool = H.gaussian_pool(9000 + seed, steps * ((n + 31) // 32 * 32))
    rc, o, ob = S.run_oracle(syn, make(), precision, steps, pool)
    for mode, decision in (("fused", "local"), ("fused", "handoff"), ("two_pass", "auto")):
        integ = make()
        integ.setDecisionMode(decision)
        try:
            grc, ctx, gb = S.run_gpu(syn, integ, precision, steps, pool, mode=mode, graph=8 if seed % 2 else 0)
        except s2.Seekr2Error as e:                     # more than 256 padded slots: LOCAL is refused, the hand-off covers it
            assert decision == "local" and "group slots" in str(e)
            continue
        assert rc == grc == 0, (seed, mode, decision, S.run_gpu.last_error)
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0)), (seed, mode, decision)
        H.assert_state_equal(ob, gb, n, f"elber seed {seed} {mode} {decision}")
        so, sg = o.state(), integ.getAnchorState(0)
        assert (so.time, so.step_count, so.crossing_counter, so.crossed_src, so.end_simulation) == \
               (sg.time, sg.step_count, sg.crossing_counter, sg.crossed_src, sg.end_simulation), (seed, mode, decision)


def test_31_milestone_bits_many_surfaces_at_once(cuda_device):
    """The reference's limit: 31 bit codes (README.md:331-333), here on 8 groups, with thresholds so close to the start that
    steps cross several surfaces at once (corner bounces, bit strings beyond float precision: (int)(float)value is decoded the
    same way on both sides)."""
    n = 400
    pos, vel, masses = H.lattice_system(n, 4242)
    rng = np.random.default_rng(99)
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    groups = [[int(i) for i in rng.choice(n, size, replace=False)] for size in (1, 2, 3, 5, 8, 13, 21, 30)]
    for b in range(31):
        axis = "xyz"[b % 3]
        f = s2.CustomCentroidBondForce(1, f"bitcode*step(k*({axis}1-bound))")
        idx = groups[b % len(groups)]
        g = f.addGroup(idx)
        for p in ("bitcode", "k", "bound"):
            f.addPerBondParameter(p)
        c = float((pos[idx] * masses[idx][:, None]).sum(0)[b % 3] / masses[idx].sum())
        sign = 1.0 if b % 2 else -1.0
        f.addBond([g], [float(2 ** b), sign, c + sign * float(rng.uniform(0.002, 0.012))])
        f.setForceGroup(1)
        system.addForce(f)
    syn = H.Synthetic(system, pos, vel, masses, None, list(range(1, 32)), 0.0, 0.002)
    steps = 300
    pool = H.gaussian_pool(4243, steps * ((n + 31) // 32 * 32))
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.MIXED, steps, pool)
    assert rc == 0 and o.state().num_bounce_steps >= 5
    assert max(r.num_surfaces for r in o.records()) >= 2, "some step must cross several surfaces"
    for mode, decision in (("fused", "local"), ("fused", "handoff"), ("two_pass", "auto")):
        integ = H.make_mmvt_integrator(syn, "")
        integ.setDecisionMode(decision)
        grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, steps, pool, mode=mode, graph=8)
        assert grc == 0
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0)), (mode, decision)
        H.assert_state_equal(ob, gb, n, f"31 bits {mode} {decision}")
        assert H.stats_tuple(o.state(), 31) == H.stats_tuple(integ.getAnchorState(0), 31), (mode, decision)


@pytest.mark.parametrize("sizes", [(128, 128), (224, 32), (97, 33), (256,)])
def test_local_mode_at_its_slot_limit(cuda_device, sizes):
    """256 padded group slots is the most the LOCAL kernel takes (512-thread blocks: 256 streaming + 256 slot threads);
    groups that fill whole warps exactly, a group spanning seven warps, one 256-atom group on a plane."""
    n = 1500
    pos, vel, masses = H.lattice_system(n, 5150 + sum(sizes))
    rng = np.random.default_rng(len(sizes) * 100 + sizes[0])
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    idx = [int(i) for i in rng.permutation(n)[: sum(sizes)]]
    parts = [idx[: sizes[0]], idx[sizes[0]:]] if len(sizes) == 2 else [idx]
    cen = [(pos[p] * masses[p][:, None]).sum(0) / masses[p].sum() for p in parts]
    if len(sizes) == 2:
        f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
        ga, gb = f.addGroup(parts[0]), f.addGroup(parts[1])
        for p in ("bitcode", "k", "radius"):
            f.addPerBondParameter(p)
        d = float(np.linalg.norm(cen[0] - cen[1]))
        f.addBond([ga, gb], [1.0, -1.0, d - 0.0002])
        f.addBond([ga, gb], [2.0, 1.0, d + 0.0002])
    else:
        f = s2.CustomCentroidBondForce(1, "bitcode*step(k*(x1-bound))")
        ga = f.addGroup(parts[0])
        for p in ("bitcode", "k", "bound"):
            f.addPerBondParameter(p)
        f.addBond([ga], [1.0, -1.0, float(cen[0][0]) - 0.0001])
        f.addBond([ga], [2.0, 1.0, float(cen[0][0]) + 0.0001])
    f.setForceGroup(1)
    system.addForce(f)
    syn = H.Synthetic(system, pos, vel, masses, None, [1, 2], 0.0, 0.002)
    steps = 400
    pool = H.gaussian_pool(5151, steps * ((n + 31) // 32 * 32))
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.MIXED, steps, pool)
    assert rc == 0 and o.state().num_bounce_steps >= 2, o.state().num_bounce_steps
    for decision in ("local", "handoff"):
        integ = H.make_mmvt_integrator(syn, "")
        integ.setDecisionMode(decision)
        grc, ctx, gb_ = S.run_gpu(syn, integ, L.MIXED, steps, pool, graph=8)
        assert grc == 0 and integ.getDecisionMode() == decision
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0)), decision
        H.assert_state_equal(ob, gb_, n, f"slot limit {sizes} {decision}")


def test_sixteen_elber_milestones(cuda_device):
    """SEEKR2_B200_MAX_ELBER = 16: one source and fifteen destination milestones (each its own Force and force group, all on
    the same two centroid groups, which are lowered once), thermal run until a destination ends it."""
    n = 900
    pos, vel, masses = H.lattice_system(n, 6161)
    a, b = list(range(3, 14)), list(range(40, 49))
    ca = (pos[a] * masses[a][:, None]).sum(0) / masses[a].sum()
    cb = (pos[b] * masses[b][:, None]).sum(0) / masses[b].sum()
    d = float(np.linalg.norm(ca - cb))
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    for g in range(16):
        f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
        ga, gb = f.addGroup(a), f.addGroup(b)
        for p in ("bitcode", "k", "radius"):
            f.addPerBondParameter(p)
        sign = 1.0 if g % 2 == 0 else -1.0
        f.addBond([ga, gb], [1.0, sign, d + sign * 0.0004 * (g + 1)])
        f.setForceGroup(g + 1)
        system.addForce(f)
    syn = H.Synthetic(system, pos, vel, masses, None, [], 0.0, 0.002)

    def make():
        integ = s2.ElberLangevinIntegrator(300.0, 1.0, 0.002, "")
        integ.addSrcMilestoneGroup(1)
        for g in range(2, 17):
            integ.addDestMilestoneGroup(g)
        integ.setEndOnSrcMilestone(False)
        return integ
    steps = 400
    pool = H.gaussian_pool(6162, steps * ((n + 31) // 32 * 32))
    rc, o, ob = S.run_oracle(syn, make(), L.MIXED, steps, pool)
    assert rc == 0 and o.state().end_simulation == 1 and len(o.records()) >= 1
    for mode, decision in (("fused", "local"), ("fused", "handoff"), ("two_pass", "auto")):
        integ = make()
        integ.setDecisionMode(decision)
        grc, ctx, gb_ = S.run_gpu(syn, integ, L.MIXED, steps, pool, mode=mode, graph=8)
        assert grc == 0
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0)), (mode, decision)
        H.assert_state_equal(ob, gb_, n, f"16 elber milestones {mode} {decision}")
        so, sg = o.state(), integ.getAnchorState(0)
        assert (so.time, so.crossing_counter, so.crossed_src, so.end_simulation) == (sg.time, sg.crossing_counter, sg.crossed_src, sg.end_simulation)
