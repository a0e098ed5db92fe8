"""Robustness of the engine's device-resident state: external state changes between steps, ring overflow,
parameter changes mid-run, statistics files, long-run statistical agreement of the in-kernel RNG path."""
import ctypes as C
import math

import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("decision_mode")]


def test_set_positions_mid_run_resyncs_group_slab(cuda_device):
    """Context.setPositions / setVelocities between steps (what a user or a barostat does): the engine
    re-gathers the boundary-group slab, and the continuation equals an oracle run restarted from that state."""
    host, guest = list(range(0, 11)), list(range(11, 20))
    syn = H.pair_sphere_system(500, host, guest, 0.185, 0.195, H.SEED_BASE + 13, tether_k=500.0)
    H.place_pair_distance(syn, host, guest, 0.19)
    pool = H.gaussian_pool(31, 200 * 512)
    integ = H.make_mmvt_integrator(syn, "")
    hb = H.make_host_buffers(L.MIXED, syn.positions, syn.velocities, syn.masses)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    H.upload(ctx, [hb])
    ctx.setRandomPool(pool)
    integ.step(100)
    mid = H.download(ctx)
    # move the guest group by hand (public API) and continue
    x = ctx.positions_host()[0]
    x[guest] += np.array([0.001, -0.0005, 0.0])
    ctx.setPositions(x)
    ctx.setVelocities(ctx.velocities_host()[0] * 0.5)
    restart = H.download(ctx)
    integ.step(100)
    got = H.download(ctx)
    # a fresh oracle started from the modified state (fresh bookkeeping, so only the atoms' evolution is
    # compared; its step-0 check is skipped by giving it the engine's step count)
    o = H.oracle_for(H.make_mmvt_integrator(syn, ""), syn.system, L.MIXED)
    o.lib.seekr2_oracle_set_time(o.h, integ.getAnchorState(0).time - 100 * syn.dt, 100)
    ob = restart.copy()
    x0 = hb.posq.copy()
    assert o.run(100, ob.posq, ob.corr, ob.velm, ob.force, pool, 100 * 512, x0, 500.0) == 0
    H.assert_state_equal(ob, got, 500, "continuation after setPositions")
    assert (mid.posq[:500] != restart.posq[:500]).any()


def test_ring_overflow_is_reported_not_silent(cuda_device):
    syn = S.ballistic_toy(speed=5.0, dt=0.01)          # bounces every ~4 steps
    integ = H.make_mmvt_integrator(syn, "")
    integ.ringCapacity = 8
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    ctx.setPositions(syn.positions)
    ctx.setVelocities(syn.velocities)
    with pytest.raises(s2.Seekr2Error) as err:
        integ.step(400)
    assert err.value.code == L.ERR_RING_OVERFLOW
    # counters on the device are still exact
    assert integ.getAnchorState(0).bounce_counter > 8
    integ.step(8)                                       # and the engine keeps working afterwards


def test_parameter_change_mid_run_and_statistics_file(cuda_device, tmp_path):
    """T / gamma / dt setters take effect at the next step (CudaSeekr2Kernels.cpp:188-203); the statistics
    file matches the oracle's byte for byte."""
    syn = H.toy_system()
    pool = H.gaussian_pool(17, 4000 * 32)
    stats_gpu, stats_cpu = str(tmp_path / "gpu_stats.txt"), str(tmp_path / "cpu_stats.txt")
    integ = H.make_mmvt_integrator(syn, str(tmp_path / "gpu.txt"))
    integ.setSaveStatisticsFileName(stats_gpu)
    hb = H.make_host_buffers(L.MIXED, syn.positions, syn.velocities, syn.masses)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    H.upload(ctx, [hb])
    ctx.setRandomPool(pool)
    o = H.oracle_for(H.make_mmvt_integrator(syn, ""), syn.system, L.MIXED)
    ob = hb.copy()
    integ.step(2000)
    assert o.run(2000, ob.posq, ob.corr, ob.velm, ob.force, pool, 0, None, 0.0) == 0
    integ.setTemperature(350.0); integ.setFriction(2.0); integ.setStepSize(0.001)
    o.set_params(350.0, 2.0, 0.001)
    integ.step(2000)
    assert o.run(2000, ob.posq, ob.corr, ob.velm, ob.force, pool, 2000 * 32, None, 0.0) == 0
    H.assert_state_equal(ob, H.download(ctx), 1)
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))
    assert len(o.records()) > 0
    o.write_statistics(stats_cpu, syn.milestone_groups)
    assert open(stats_gpu).read() == open(stats_cpu).read()


def test_long_run_statistics_in_kernel_rng_vs_oracle(cuda_device):
    """North star: long-run MMVT statistics agree within statistical error when the random numbers differ.
    256 replicas of the toy system with the in-kernel Philox generator vs 16 oracle replicas with the
    host-generated pool: bounce rate per milestone and mean incubation times R_i / N_ij."""
    syn = H.toy_system()
    steps = 40000
    integ = H.make_mmvt_integrator(syn, "")
    integ.setRandomNumberSeed(2024)
    integ.setUseGraph(True, 200)
    integ.ringCapacity = 4096
    A = 256
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
    ctx.setPositions(syn.positions)
    ctx.setVelocitiesToTemperature(300.0, 5)
    integ.step(steps)
    g_n = np.zeros(2); g_t = 0.0; g_r = np.zeros(2); g_nij = np.zeros(2)
    for a in range(A):
        st = integ.getAnchorState(a)
        g_n += np.array(st.N_alpha_beta[:2]); g_t += st.T_alpha
        g_r += np.array(st.Ri_alpha[:2]); g_nij += np.array([st.Nij_alpha[0 * L.MAX_MILESTONES + 1], st.Nij_alpha[1 * L.MAX_MILESTONES + 0]])
    c_n = np.zeros(2); c_t = 0.0; c_r = np.zeros(2); c_nij = np.zeros(2)
    R = 16
    rng = np.random.default_rng(1)
    for r in range(R):
        o = H.oracle_for(H.make_mmvt_integrator(syn, ""), syn.system, L.MIXED)
        vel = rng.standard_normal((1, 3)) * math.sqrt(s2.BOLTZ * 300.0 / 39.9)
        hb = H.make_host_buffers(L.MIXED, syn.positions, vel, syn.masses)
        pool = np.zeros((steps * 32, 4), dtype=np.float32)
        pool[:, :3] = rng.standard_normal((steps * 32, 3)).astype(np.float32)
        assert o.run(steps, hb.posq, hb.corr, hb.velm, hb.force, pool, 0, None, 0.0) == 0
        st = o.state()
        c_n += np.array(st.N_alpha_beta[:2]); c_t += st.T_alpha
        c_r += np.array(st.Ri_alpha[:2]); c_nij += np.array([st.Nij_alpha[0 * L.MAX_MILESTONES + 1], st.Nij_alpha[1 * L.MAX_MILESTONES + 0]])
    assert c_n.min() > 150 and g_n.min() > 2000
    for i in range(2):
        rate_g, rate_c = g_n[i] / g_t, c_n[i] / c_t
        sigma = rate_c / math.sqrt(c_n[i]) + rate_g / math.sqrt(g_n[i])      # Poisson counting error
        assert abs(rate_g - rate_c) < 5 * sigma, (i, rate_g, rate_c, sigma)
        inc_g, inc_c = g_r[i] / g_nij[i], c_r[i] / c_nij[i]                  # mean incubation time before a switch
        assert abs(inc_g - inc_c) < 0.25 * inc_c, (i, inc_g, inc_c)


def test_asynchronous_drain_pipelined_inside_one_long_step_call(cuda_device, tmp_path):
    """step(n) queues an asynchronous drain after every `drainInterval` steps and collects it while the next chunk
    runs: a run that appends many times the ring capacity loses nothing, and file + record list equal the oracle's."""
    syn = S.ballistic_toy(speed=5.0, dt=0.01)          # bounces every ~4 steps
    steps = 20000
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, ""), L.MIXED, steps)
    path = str(tmp_path / "long.txt")
    integ = H.make_mmvt_integrator(syn, path)
    integ.ringCapacity = 512
    integ.drainInterval = 1000
    integ.setUseGraph(True, 8)
    hb = H.make_host_buffers(L.MIXED, syn.positions, syn.velocities, syn.masses)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    H.upload(ctx, [hb])
    ctx.setRandomPool(np.zeros((steps * 32, 4), dtype=np.float32))
    integ.step(steps)
    assert len(o.records()) > 8 * 512
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))
    H.assert_state_equal(ob, H.download(ctx), 1)
    opath = str(tmp_path / "oracle.txt")
    o.write_log(opath, syn.milestone_groups)
    assert open(opath).read() == open(path).read()


def test_drain_begin_end_contract(cuda_device):
    """C ABI: one drain in flight; end() returns what had been appended when begin() was queued, not later records;
    a small max_records leaves the rest waiting."""
    syn = S.ballistic_toy(speed=5.0, dt=0.01)
    integ = H.make_mmvt_integrator(syn, "")
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    ctx.setPositions(syn.positions)
    ctx.setVelocities(syn.velocities)
    h, stream = integ._handle, ctx.stream_ptr
    buf = (L.Record * 4096)()
    n, waiting = C.c_int64(0), C.c_int64(0)
    assert L.lib.seekr2_b200_drain_end(h, buf, 4096, C.byref(n), None) == L.ERR_INVALID          # nothing in flight
    integ.drainInterval = 1 << 30
    integ._update_params(); ctx.prepareForces(); integ._first_step_check(); integ._before_first_step()
    integ._issue(200)
    assert L.lib.seekr2_b200_drain_begin(h, 4096, stream) == L.OK
    assert L.lib.seekr2_b200_drain_begin(h, 4096, stream) == L.ERR_INVALID                         # one at a time
    integ._issue(200)                                                                              # queued BEHIND the drain
    assert L.lib.seekr2_b200_drain_end(h, buf, 4096, C.byref(n), C.byref(waiting)) == L.OK
    first = n.value
    assert first > 10 and waiting.value == 0 and max(buf[i].step for i in range(first)) < 200
    assert L.lib.seekr2_b200_drain_begin(h, 5, stream) == L.OK                                     # small window
    assert L.lib.seekr2_b200_drain_end(h, buf, 5, C.byref(n), C.byref(waiting)) == L.OK
    assert n.value == 5 and waiting.value > 0 and min(buf[i].step for i in range(5)) >= 200
    left = waiting.value
    assert L.lib.seekr2_b200_drain(h, buf, 4096, C.byref(n), stream) == L.OK and n.value == left
    assert integ.getAnchorState(0).ring_head == first + 5 + left


def test_poll_reports_appended_records_without_a_cuda_call(cuda_device):
    """seekr2_b200_poll reads mapped host memory only: the step kernels mirror the ring head there on crossing steps."""
    syn = S.ballistic_toy(speed=5.0, dt=0.01)          # bounces every ~4 steps
    integ = H.make_mmvt_integrator(syn, "")
    integ.drainInterval = 10 ** 9                       # step() below drains once, at its end
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    ctx.setPositions(syn.positions)
    ctx.setVelocities(syn.velocities)
    pending, resets = C.c_int64(-1), C.c_int64(-1)
    L.check(L.lib.seekr2_b200_poll(integ._handle, C.byref(pending), C.byref(resets)))
    assert (pending.value, resets.value) == (0, 0)
    integ._update_params(); ctx.prepareForces(); integ._first_step_check(); integ._before_first_step()
    integ._issue(200)                                   # queued, nothing drained
    import torch
    torch.cuda.synchronize()
    L.check(L.lib.seekr2_b200_poll(integ._handle, C.byref(pending), C.byref(resets)))
    st = integ.getAnchorState(0)
    assert pending.value == st.ring_head > 20
    got = integ.drain()
    assert len(got) == st.ring_head
    L.check(L.lib.seekr2_b200_poll(integ._handle, C.byref(pending), C.byref(resets)))
    assert pending.value == 0


@pytest.mark.parametrize("decision", ["local", "handoff"])
@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
def test_resync_every_step_and_permuted_atom_order(cuda_device, precision, decision):
    """What the OpenMM adapter does: seekr2_b200_resync in front of every fused step (the slab is re-read from the caller's
    buffers), and seekr2_b200_set_atom_order when the caller's buffers hold the atoms in another order.  The same system
    with its atoms stored in reversed order must evolve identically (in-kernel normals are keyed by buffer index, so an
    injected pool, permuted alike, is used).  With a pool the fused step is launched as a programmatic dependent of the
    resync kernel and reads slab and step index behind its wait (late_step), in both fused kernels."""
    host, guest = list(range(0, 11)), list(range(11, 20))
    n = 300
    syn = H.pair_sphere_system(n, host, guest, 0.185, 0.195, H.SEED_BASE + 21, tether_k=0.0)
    H.place_pair_distance(syn, host, guest, 0.19)
    steps = 300
    npad = (n + 31) // 32 * 32
    pool = H.gaussian_pool(77, steps * npad)
    integ = H.make_mmvt_integrator(syn, "")
    rc, o, ob = S.run_oracle(syn, integ, precision, steps, pool)
    assert rc == 0 and o.state().bounce_counter > 0

    def run(order):
        integ = H.make_mmvt_integrator(syn, "")
        integ.setDecisionMode(decision)
        names = {L.SINGLE: "single", L.MIXED: "mixed", L.DOUBLE: "double"}
        hb = H.make_host_buffers(precision, syn.positions, syn.velocities, syn.masses)
        perm = np.arange(npad)
        perm[:n] = order                                # buffer index i holds atom order[i]
        hb.posq = hb.posq[perm].copy(); hb.velm = hb.velm[perm].copy()
        if hb.corr is not None:
            hb.corr = hb.corr[perm].copy()
        ppool = pool.reshape(steps, npad, 4)[:, perm].reshape(-1, 4).copy()
        ctx = s2.Context(syn.system, integ, {"CudaPrecision": names[precision]})
        H.upload(ctx, [hb])
        ctx.setRandomPool(ppool)
        index_of = np.empty(n, dtype=np.int32)
        index_of[order] = np.arange(n, dtype=np.int32)
        L.check(L.lib.seekr2_b200_set_atom_order(integ._handle, index_of.ctypes.data_as(C.POINTER(C.c_int32)), ctx.stream_ptr))
        integ._update_params(); ctx.prepareForces()
        for s in range(steps):
            bufs = ctx.buffers()
            L.check(L.lib.seekr2_b200_resync(integ._handle, C.byref(bufs), ctx.stream_ptr))
            L.check(L.lib.seekr2_b200_step(integ._handle, C.byref(bufs), s * npad, ctx.stream_ptr))
        got = H.download(ctx)
        inv = np.arange(npad); inv[perm] = np.arange(npad)
        got.posq = got.posq[inv]; got.velm = got.velm[inv]
        if got.corr is not None:
            got.corr = got.corr[inv]
        return got, integ.drain(), integ.getAnchorState(0)

    for order in (np.arange(n), np.arange(n)[::-1].copy()):
        got, recs, st = run(order)
        H.assert_state_equal(ob, got, n, "resync + fused step vs oracle")
        assert st.bounce_counter == o.state().bounce_counter and len(recs) == len(o.records())
