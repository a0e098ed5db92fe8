"""The two decision modes of the fused step (include/seekr2_b200.h FLAG_DECIDE_LOCAL / FLAG_DECIDE_HANDOFF) are the same
function: LOCAL (every block advances the boundary-group atoms and evaluates the boundaries itself) and HANDOFF (one decide
block per anchor publishes the decision) give bit-identical atoms, crossing records and statistics -- and both equal the CPU
oracle (the other GPU test modules run in both modes through the `decision_mode` fixture)."""
import numpy as np
import pytest

import helpers as H
from helpers import L, s2

pytestmark = pytest.mark.gpu


def _run(syn, A, positions, mode, steps, graph, middle=False, kind="mmvt", pool=None, seed=11):
    if kind == "mmvt":
        integ = H.make_mmvt_integrator(syn, "", middle=middle)
    else:
        integ = (s2.ElberLangevinMiddleIntegrator if middle else s2.ElberLangevinIntegrator)(syn.temperature, syn.friction, syn.dt, "")
        integ.addSrcMilestoneGroup(1)
        integ.setEndOnSrcMilestone(False)
    integ.setDecisionMode(mode)
    integ.setRandomNumberSeed(seed)
    if graph:
        integ.setUseGraph(True, graph)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
    assert integ.getDecisionMode() == mode
    ctx.setPositions(positions)
    ctx.setVelocities(syn.velocities)
    if pool is not None:
        ctx.setRandomPool(pool)
    integ.step(steps)
    recs = [H.records_as_tuples(integ.getRecords(a)) for a in range(A)]
    stats = [H.stats_tuple(integ.getAnchorState(a), 2) for a in range(A)]
    clock = [(integ.getAnchorState(a).time, integ.getAnchorState(a).step_count) for a in range(A)]
    return [H.download(ctx, a) for a in range(A)], recs, stats, clock


def _shell_workload(A, n_atoms, seed):
    rec, lig = list(range(0, 11)), list(range(11, 20))
    syn = H.pair_sphere_system(n_atoms, rec, lig, 0.10, 0.14, seed, tether_k=0.0)
    positions = np.empty((A, n_atoms, 3))
    base = syn.positions.copy()
    for a in range(A):
        rin, rout = 0.10 + 0.04 * a, 0.14 + 0.04 * a
        syn.boundary_force.setAnchorBondParameters(a, 0, [1.0, -1.0, rin])
        syn.boundary_force.setAnchorBondParameters(a, 1, [2.0, 1.0, rout])
        syn.positions = base.copy()
        H.place_pair_distance(syn, rec, lig, 0.5 * (rin + rout))
        positions[a] = syn.positions
    syn.positions = base
    return syn, positions


@pytest.mark.parametrize("kind,middle", [("mmvt", False), ("mmvt", True), ("elber", False), ("elber", True)])
@pytest.mark.parametrize("graph", [0, 8])
def test_local_and_handoff_modes_are_bit_identical(cuda_device, kind, middle, graph):
    """4 batched anchors x 3 000 atoms (12 blocks per anchor), in-kernel Philox normals, 400 steps with bounces."""
    A, n = 4, 3000
    out = {}
    for mode in ("local", "handoff"):
        syn, positions = _shell_workload(A, n, H.SEED_BASE + 61)
        out[mode] = _run(syn, A, positions, mode, 400, graph, middle=middle, kind=kind)
    (gl, rl, sl, cl), (gh, rh, sh, ch) = out["local"], out["handoff"]
    assert rl == rh and sl == sh and cl == ch
    if kind == "mmvt":
        assert sum(len(r) for r in rl) >= 4, "the run must bounce"
    for a in range(A):
        H.assert_state_equal(gl[a], gh[a], n, f"local vs handoff, anchor {a}")


def test_auto_mode_picks_by_launch_size(cuda_device):
    """Default: a launch that is resident at once runs LOCAL, a batched one the hand-off; save-state slots force the hand-off."""
    syn, positions = _shell_workload(1, 3000, H.SEED_BASE + 62)
    integ = H.make_mmvt_integrator(syn, "")
    s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    assert integ.getDecisionMode() == "local"
    syn, positions = _shell_workload(64, 3000, H.SEED_BASE + 62)       # 768 streaming blocks: several tiles per resident block
    integ = H.make_mmvt_integrator(syn, "")
    s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=64)
    assert integ.getDecisionMode() == "local"
    syn, positions = _shell_workload(512, 3000, H.SEED_BASE + 62)      # 6 144 streaming blocks (> 30 x SMs): HBM-bound batch
    integ = H.make_mmvt_integrator(syn, "")
    s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=512)
    assert integ.getDecisionMode() == "handoff"
    syn, positions = _shell_workload(1, 3000, H.SEED_BASE + 62)
    integ = H.make_mmvt_integrator(syn, "")
    integ.setSaveStateFileName("/tmp/seekr2_b200_state")
    s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    assert integ.getDecisionMode() == "handoff"
    integ = H.make_mmvt_integrator(syn, "")
    integ.setSaveStateFileName("/tmp/seekr2_b200_state")
    integ.setDecisionMode("local")
    with pytest.raises(s2.Seekr2Error):
        s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})


@pytest.mark.parametrize("mode", ["local", "handoff"])
def test_two_pass_and_fused_steps_interleave_on_one_random_stream(cuda_device, mode):
    """In-kernel Philox normals are keyed by the step index.  Two-pass steps advance the device clock but not the fused
    kernels' four-phase step-index slots; seekr2_b200_sync_groups (mandatory before the next fused step) refreshes them,
    so a run that switches between the forms equals an all-fused run bit for bit."""
    A, n = 2, 1200
    out = {}
    for plan in ("fused", "mixed"):
        syn, positions = _shell_workload(A, n, H.SEED_BASE + 63)
        integ = H.make_mmvt_integrator(syn, "")
        integ.setDecisionMode(mode)
        integ.setRandomNumberSeed(21)
        ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
        ctx.setPositions(positions)
        ctx.setVelocities(syn.velocities)
        if plan == "fused":
            integ.step(150)
        else:
            for form, steps in (("fused", 37), ("two_pass", 41), ("fused", 30), ("two_pass", 2), ("fused", 40)):
                integ.setStepMode(form)
                integ.step(steps)
        out[plan] = ([H.download(ctx, a) for a in range(A)], [H.records_as_tuples(integ.getRecords(a)) for a in range(A)],
                     [integ.getAnchorState(a).step_count for a in range(A)])
    assert out["fused"][1] == out["mixed"][1] and out["fused"][2] == out["mixed"][2] == [150] * A
    assert sum(len(r) for r in out["fused"][1]) >= 2
    for a in range(A):
        H.assert_state_equal(out["fused"][0][a], out["mixed"][0][a], n, f"all-fused vs interleaved, anchor {a}")


def test_long_run_local_equals_handoff(cuda_device):
    """20 000 steps of 3 tethered anchors under 32-step graph replay (programmatic dependent launch, four-phase step
    index, asynchronous drains pipelined behind the chunks): LOCAL and HANDOFF end in the same bits after hundreds of
    bounces, and the crossing records are the same sequence."""
    A, n, steps = 3, 2000, 20000
    out = {}
    for mode in ("local", "handoff"):
        rec, lig = list(range(0, 11)), list(range(11, 20))
        syn = H.pair_sphere_system(n, rec, lig, 0.185, 0.195, H.SEED_BASE + 64, tether_k=800.0)
        H.place_pair_distance(syn, rec, lig, 0.19)
        out[mode] = _run(syn, A, syn.positions, mode, steps, 32, seed=29)
    (gl, rl, sl, cl), (gh, rh, sh, ch) = out["local"], out["handoff"]
    assert rl == rh and sl == sh and cl == ch
    assert min(len(r) for r in rl) >= 20, [len(r) for r in rl]
    assert all(c == (cl[0][0], steps) for c in cl)
    for a in range(A):
        H.assert_state_equal(gl[a], gh[a], n, f"local vs handoff after {steps} steps, anchor {a}")


KNOB_SETS = [
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "128"},
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "64", "SEEKR2_B200_LOCAL_RESIDENT": "6"},      # one block per anchor walks every tile
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "96", "SEEKR2_B200_LOCAL_RESIDENT": "20"},
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_PACKED": "0"},                                              # aligned layout for one-warp groups
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_PDL": "0"},
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_REGS": "72"},                                         # the throughput-regime instantiation
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_REGS": "72", "SEEKR2_B200_LOCAL_THREADS": "128", "SEEKR2_B200_LOCAL_RESIDENT": "40"},
    # the undo log in shared memory (MMVT): off (every block waits for the decision after its first tile); one or two logged tiles
    # of the many a block walks (log, then one tile from registers behind the decision barrier, then predicated stores); the slot
    # warp joining the undo barrier on bounce steps only
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_UNDO": "0"},
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "64", "SEEKR2_B200_LOCAL_RESIDENT": "6", "SEEKR2_B200_LOCAL_UNDO_TILES": "1"},
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "96", "SEEKR2_B200_LOCAL_RESIDENT": "20", "SEEKR2_B200_LOCAL_UNDO_TILES": "2", "SEEKR2_B200_TUNE": "32"},
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_UNDO_TILES": "1", "SEEKR2_B200_TUNE": "32"},          # the only tile of every block logged
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "64", "SEEKR2_B200_LOCAL_RESIDENT": "6", "SEEKR2_B200_LOCAL_UNDO_TILES": "99"},   # 47 tiles, all logged (144 KB)
    # the one-pair fast path of the packed slot warp (default here: difference of the centroid sums reduced directly, activity read
    # off a comparison): off, and with the general term evaluation
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_PAIR_FAST": "0"},
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_PAIR_FAST": "0", "SEEKR2_B200_PACKED": "0"},               # general sums, aligned layout
    {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_PAIR_FAST": "1", "SEEKR2_B200_LOCAL_THREADS": "160"},
]


@pytest.mark.parametrize("kind,middle,variant,restore", [("mmvt", False, "cuda", False), ("mmvt", True, "cuda", True),
                                                         ("mmvt", False, "reference", True), ("elber", False, "cuda", False)])
def test_kernel_knobs_do_not_change_a_bit(cuda_device, monkeypatch, kind, middle, variant, restore):
    """The LOCAL kernel's launch shape (streaming threads per block, tiles per block), the packed / aligned slot layout, early
    normals, programmatic dependent launch and the register budget of the LOCAL instantiation are tuning knobs: every setting must
    give the bits of the default hand-off kernel (which the other modules pin to the oracle) -- atoms, records, statistics,
    clock -- on 4 anchors x 3 000 atoms over 400 steps with bounces, under 8-step graph replay."""
    A, n = 4, 3000

    def run(env):
        for k in ("SEEKR2_B200_DECIDE", "SEEKR2_B200_LOCAL_THREADS", "SEEKR2_B200_LOCAL_RESIDENT", "SEEKR2_B200_TUNE", "SEEKR2_B200_PACKED", "SEEKR2_B200_PDL", "SEEKR2_B200_LOCAL_REGS",
                  "SEEKR2_B200_LOCAL_UNDO", "SEEKR2_B200_LOCAL_UNDO_TILES", "SEEKR2_B200_PAIR_FAST"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        syn, positions = _shell_workload(A, n, H.SEED_BASE + 65)
        if kind == "mmvt":
            integ = H.make_mmvt_integrator(syn, "", variant=variant, middle=middle)
        else:
            integ = s2.ElberLangevinIntegrator(syn.temperature, syn.friction, syn.dt, "")
            integ.addSrcMilestoneGroup(1)
            integ.setEndOnSrcMilestone(False)
        integ.setRestoreCorrection(restore)
        integ.setRandomNumberSeed(13)
        integ.setUseGraph(True, 8)
        ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
        assert integ.getDecisionMode() == ("local" if env.get("SEEKR2_B200_DECIDE") == "l" else "handoff")
        ctx.setPositions(positions)
        ctx.setVelocities(syn.velocities)
        integ.step(400)
        return ([H.download(ctx, a) for a in range(A)], [H.records_as_tuples(integ.getRecords(a)) for a in range(A)],
                [H.stats_tuple(integ.getAnchorState(a), 2) for a in range(A)],
                [(integ.getAnchorState(a).time, integ.getAnchorState(a).step_count) for a in range(A)])

    ref = run({"SEEKR2_B200_DECIDE": "h"})
    if kind == "mmvt":
        assert sum(len(r) for r in ref[1]) >= 4, "the run must bounce"
    for env in KNOB_SETS:
        got = run(env)
        assert got[1] == ref[1] and got[2] == ref[2] and got[3] == ref[3], env
        for a in range(A):
            H.assert_state_equal(ref[0][a], got[0][a], n, f"{env}, anchor {a}")


@pytest.mark.parametrize("groups", [(11, 9), (40, 9)])        # packed slot layout (one warp) / aligned layout (whole warps per group)
@pytest.mark.parametrize("offset_nm", [0.0, 700.0, 6000.0])
def test_one_pair_fast_path_and_its_fallback_match_the_oracle(cuda_device, groups, offset_nm):
    """The LOCAL kernel's one-pair fast path sums the DIFFERENCE of the two centroid sums and needs every fixed-point term below
    2^47 (packed) / 2^45 (aligned); a system far from the origin exceeds that and must take the general path.  The boundary depends
    on the difference of two centroids only, so the same run translated by 0, 700 and 6 000 nm (|w x| 2^40 up to 2^48.6 for the
    smaller group) is checked against the oracle bit for bit, in mixed and double precision: state, records, statistics."""
    import scenarios as S
    na, nb = groups
    host, guest = list(range(0, na)), list(range(na, na + nb))
    for precision in (L.MIXED, L.DOUBLE):
        syn = H.pair_sphere_system(700, host, guest, 0.198, 0.202, H.SEED_BASE + 71, tether_k=0.0)
        H.place_pair_distance(syn, host, guest, 0.20)
        syn.positions = syn.positions + np.array([offset_nm, -0.5 * offset_nm, 0.25 * offset_nm])
        steps = 240
        pool = H.gaussian_pool(H.SEED_BASE + 72, steps * ((700 + 31) // 32 * 32))
        rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), precision, steps, pool)
        integ = H.make_mmvt_integrator(syn, "")
        integ.setDecisionMode("local")
        grc, ctx, gb = S.run_gpu(syn, integ, precision, steps, pool, mode="fused", graph=8)
        assert rc == grc == 0, (groups, offset_nm, precision, S.run_gpu.last_error)
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0)), (groups, offset_nm, precision)
        H.assert_state_equal(ob, gb, 700, f"groups {groups}, offset {offset_nm} nm, precision {precision}")
        assert H.stats_tuple(o.state(), 2) == H.stats_tuple(integ.getAnchorState(0), 2)
        if offset_nm == 0.0:
            assert len(o.records()) > 0, "the run must bounce"


def test_two_engines_with_different_undo_logs_share_the_kernel(cuda_device, monkeypatch):
    """The dynamic shared memory limit of the LOCAL kernel is an attribute of the FUNCTION: an engine that sizes a small undo log
    must not lower it under an engine of the same process that launches the same instantiation with a large one."""
    A, n = 4, 3000

    def make(env):
        for k in ("SEEKR2_B200_DECIDE", "SEEKR2_B200_LOCAL_THREADS", "SEEKR2_B200_LOCAL_RESIDENT", "SEEKR2_B200_LOCAL_UNDO_TILES"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        syn, positions = _shell_workload(A, n, H.SEED_BASE + 65)
        integ = H.make_mmvt_integrator(syn, "")
        integ.setRandomNumberSeed(13)
        ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
        ctx.setPositions(positions)
        ctx.setVelocities(syn.velocities)
        integ.step(8)                                          # the first launch sizes the grid and the log
        return ctx, integ

    big = {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "64", "SEEKR2_B200_LOCAL_RESIDENT": "6", "SEEKR2_B200_LOCAL_UNDO_TILES": "99"}   # 144 KB
    small = {"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "64", "SEEKR2_B200_LOCAL_RESIDENT": "6", "SEEKR2_B200_LOCAL_UNDO_TILES": "1"}  # 3 KB
    ctx_ref, integ_ref = make(big)
    integ_ref.step(392)
    ctx_a, integ_a = make(big)
    ctx_b, integ_b = make(small)
    for _ in range(49):
        integ_a.step(8)
        integ_b.step(8)
    assert sum(len(integ_ref.getRecords(a)) for a in range(A)) >= 4, "the run must bounce"
    for a in range(A):
        H.assert_state_equal(H.download(ctx_ref, a), H.download(ctx_a, a), n, f"large log, anchor {a}")
        H.assert_state_equal(H.download(ctx_ref, a), H.download(ctx_b, a), n, f"small log, anchor {a}")
        assert H.records_as_tuples(integ_ref.getRecords(a)) == H.records_as_tuples(integ_a.getRecords(a)) == H.records_as_tuples(integ_b.getRecords(a))


@pytest.mark.parametrize("precision", ["single", "double"])
@pytest.mark.parametrize("middle", [False, True])
def test_undo_log_in_single_and_double_precision(cuda_device, monkeypatch, precision, middle):
    """The undo log's entry is (real4, mixed4) of the precision mode -- 32 B per atom in single, 64 B in double precision.  Blocks of
    64 streaming threads walking all 47 tiles of their anchor with two of them logged (then one tile behind the decision barrier,
    then predicated stores), against the default hand-off kernel: atoms, records, statistics, clock, with bounces."""
    A, n = 4, 3000
    keys = ("SEEKR2_B200_DECIDE", "SEEKR2_B200_LOCAL_THREADS", "SEEKR2_B200_LOCAL_RESIDENT", "SEEKR2_B200_LOCAL_UNDO_TILES")

    def run(env):
        for k in keys:
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        syn, positions = _shell_workload(A, n, H.SEED_BASE + 67)
        integ = H.make_mmvt_integrator(syn, "", middle=middle)
        integ.setRandomNumberSeed(17)
        integ.setUseGraph(True, 8)
        ctx = s2.Context(syn.system, integ, {"CudaPrecision": precision}, numAnchors=A)
        ctx.setPositions(positions)
        ctx.setVelocities(syn.velocities)
        integ.step(400)
        return ([H.download(ctx, a) for a in range(A)], [H.records_as_tuples(integ.getRecords(a)) for a in range(A)],
                [H.stats_tuple(integ.getAnchorState(a), 2) for a in range(A)])

    ref = run({"SEEKR2_B200_DECIDE": "h"})
    assert sum(len(r) for r in ref[1]) >= 4, "the run must bounce"
    for tiles in ("2", "99"):
        got = run({"SEEKR2_B200_DECIDE": "l", "SEEKR2_B200_LOCAL_THREADS": "64", "SEEKR2_B200_LOCAL_RESIDENT": "6", "SEEKR2_B200_LOCAL_UNDO_TILES": tiles})
        assert got[1] == ref[1] and got[2] == ref[2], (precision, middle, tiles)
        for a in range(A):
            H.assert_state_equal(ref[0][a], got[0][a], n, f"{precision}, middle={middle}, {tiles} logged tiles, anchor {a}")
