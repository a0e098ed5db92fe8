"""LangevinMiddle scheme on the B200: the reference's own langevinMiddle.cu kernels (oracle/_ref) vs the
CPU oracle vs the fused B200 step, bit-exact in all precisions; MMVT / Elber Middle trajectories vs the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("decision_mode")]
NAMES = {L.SINGLE: "single", L.MIXED: "mixed", L.DOUBLE: "double"}


def _ref_lib(precision):
    path = os.path.join(H.oracle_binding.REF_DIR, f"libref_langevin_middle_{NAMES[precision]}.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: the reference kernels (oracle/build_ref.sh, needs /root/reference) must be built before the GPU suite runs -- a skipped pin is not a green pin")
    lib = C.CDLL(path)
    assert lib.ref_precision() == precision
    return lib


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
@pytest.mark.parametrize("n,blocks", [(1, 0), (777, 3), (5000, 0)])
def test_reference_middle_kernels_vs_oracle_and_b200(cuda_device, precision, n, blocks):
    import torch
    ref = _ref_lib(precision)
    pos, vel, masses = H.lattice_system(n, 300 + n)
    masses[::17] = 0.0
    forces = H.gaussian_pool(n + 7, n)[:, :3].astype(np.float64) * 300.0
    hb = H.make_host_buffers(precision, pos, vel, masses, forces)
    npad = hb.posq.shape[0]
    pool = H.gaussian_pool(n + 11, npad + 64)
    ri = 37
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    integ = s2.MmvtLangevinMiddleIntegrator(300.0, 1.0, 0.002, "")
    o = H.oracle_for(integ, system, precision)
    vs, fs, ns = o.params()

    dev = cuda_device
    t = lambda a: torch.from_numpy(a.copy()).to(dev)
    posq, velm, force, rnd = t(hb.posq), t(hb.velm), t(hb.force), t(pool)
    corr = t(hb.corr) if hb.corr is not None else None
    delta, old_delta = torch.zeros((npad, 4), dtype=velm.dtype, device=dev), torch.zeros((npad, 4), dtype=velm.dtype, device=dev)
    old_posq, old_velm = torch.zeros_like(posq), torch.zeros_like(velm)
    params = torch.tensor([vs, ns], dtype=velm.dtype, device=dev)       # {VelScale, NoiseScale} (langevinMiddle.cu:1)
    dt = torch.tensor([0.0, 0.002], dtype=velm.dtype, device=dev)
    P = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
    assert ref.ref_middle_part1(n, npad, P(velm), P(force), P(dt), blocks, None) == 0
    assert ref.ref_middle_part2(n, P(velm), P(delta), P(old_delta), P(params), P(dt), P(old_velm), P(rnd), C.c_uint(ri), blocks, None) == 0
    assert ref.ref_middle_part3(n, P(posq), P(velm), P(delta), P(old_delta), P(dt), P(old_posq), P(corr), blocks, None) == 0
    torch.cuda.synchronize()
    ref_after = H.HostBuffers(precision, posq.cpu().numpy(), None if corr is None else corr.cpu().numpy(), velm.cpu().numpy(), hb.force)
    assert ref.ref_middle_bounce(n, npad, P(posq), P(velm), P(old_posq), P(old_velm), blocks, None) == 0
    torch.cuda.synchronize()
    ref_bounced = H.HostBuffers(precision, posq.cpu().numpy(), None if corr is None else corr.cpu().numpy(), velm.cpu().numpy(), hb.force)

    # CPU oracle (no boundaries => plain Middle step)
    ob = hb.copy()
    assert o.step(L.KIND_MMVT, ob.posq, ob.corr, ob.velm, ob.force, pool, ri) == 0
    H.assert_state_equal(ob, ref_after, n, "oracle vs reference Middle kernels")

    # B200 fused step
    padded_pool = np.zeros((2 * npad + 64, 4), dtype=np.float32)
    padded_pool[: pool.shape[0]] = pool
    integ3 = s2.MmvtLangevinMiddleIntegrator(300.0, 1.0, 0.002, "")
    ctx3 = s2.Context(system, integ3, {"CudaPrecision": NAMES[precision]})
    H.upload(ctx3, [hb])
    ctx3.setRandomPool(padded_pool)
    integ3._update_params()
    L.check(L.lib.seekr2_b200_step(integ3._handle, C.byref(ctx3.buffers()), ri, ctx3.stream_ptr))
    H.assert_state_equal(H.download(ctx3), ref_after, n, "B200 fused Middle step vs reference Middle kernels")

    # bounce: one boundary that is crossed for sure (plane far below every atom's x)
    system_b = s2.System()
    for m in masses:
        system_b.addParticle(float(m))
    f = s2.CustomCentroidBondForce(1, "bitcode*step(k*(x1-bound))")
    g = f.addGroup([1 % n], [1.0])          # explicit weight: the atom may be massless
    f.setForceGroup(1)
    for name in ("bitcode", "k", "bound"):
        f.addPerBondParameter(name)
    f.addBond([g], [1.0, 1.0, -1000.0])
    system_b.addForce(f)
    integ4 = s2.MmvtLangevinMiddleIntegrator(300.0, 1.0, 0.002, "")
    integ4.addMilestoneGroup(1)
    ctx4 = s2.Context(system_b, integ4, {"CudaPrecision": NAMES[precision]})
    H.upload(ctx4, [hb])
    ctx4.setRandomPool(padded_pool)
    integ4._update_params()
    L.check(L.lib.seekr2_b200_sync_groups(integ4._handle, C.byref(ctx4.buffers()), ctx4.stream_ptr))
    L.check(L.lib.seekr2_b200_step(integ4._handle, C.byref(ctx4.buffers()), ri, ctx4.stream_ptr))
    got = H.download(ctx4)
    assert integ4.getAnchorState(0).num_bounce_steps == 1
    np.testing.assert_array_equal(got.posq[:n].view(np.uint8), ref_bounced.posq[:n].view(np.uint8))
    np.testing.assert_array_equal(got.velm[:n].view(np.uint8), ref_bounced.velm[:n].view(np.uint8))
    if got.corr is not None:      # the reference's bounce leaves Part3's new correction in place
        np.testing.assert_array_equal(got.corr[:n].view(np.uint8), ref_after.corr[:n].view(np.uint8))


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
@pytest.mark.parametrize("variant", ["cuda", "reference"])
def test_toy_mmvt_middle(cuda_device, tmp_path, precision, variant):
    syn = H.toy_system()
    pool = H.gaussian_pool(H.SEED_BASE + 71, 20000 * 32)
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, "", variant, middle=True), precision, 20000, pool)
    integ = H.make_mmvt_integrator(syn, "", variant, middle=True)
    grc, ctx, gb = S.run_gpu(syn, integ, precision, 20000, pool, graph=50)
    assert rc == grc == 0 and len(o.records()) > 10
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))
    H.assert_state_equal(ob, gb, 1)
    assert H.stats_tuple(o.state(), 2) == H.stats_tuple(integ.getAnchorState(0), 2)


def test_hostguest_and_elber_middle(cuda_device):
    host, guest = list(range(0, 147)), list(range(147, 162))
    syn = H.pair_sphere_system(2520, host, guest, 0.39, 0.41, H.SEED_BASE + 2, tether_k=1000.0)
    H.place_pair_distance(syn, host, guest, 0.40)
    pool = H.gaussian_pool(H.SEED_BASE + 72, 600 * 2528)
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, "", middle=True), L.MIXED, 600, pool)
    integ = H.make_mmvt_integrator(syn, "", middle=True)
    grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, 600, pool, graph=20)
    assert rc == grc == 0 and o.state().num_bounce_steps > 0
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))
    H.assert_state_equal(ob, gb, 2520)
    for end, speed in ((False, 0.5), (True, 0.5), (False, -0.5)):
        syn_e, integ_o = S.elber_ballistic(end, speed=speed, middle=True)
        rc, o, ob = S.run_oracle(syn_e, integ_o, L.MIXED, 120)
        syn_e2, integ_e = S.elber_ballistic(end, speed=speed, middle=True)
        grc, ctx, gb = S.run_gpu(syn_e2, integ_e, L.MIXED, 120)
        assert rc == grc == 0 and len(o.records()) == 1
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ_e.getRecords(0))
        H.assert_state_equal(ob, gb, 1)
        assert H.stats_tuple(o.state(), 0) == H.stats_tuple(integ_e.getAnchorState(0), 0)


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
def test_middle_multi_pass_stage_by_stage_vs_reference_kernels(cuda_device, precision):
    """seekr2_b200_kick / middle_drift / finish == the reference's Part1 / Part2 / Part3 after every stage."""
    import torch
    ref = _ref_lib(precision)
    n = 3000
    pos, vel, masses = H.lattice_system(n, 41)
    masses[::13] = 0.0
    forces = H.gaussian_pool(43, n)[:, :3].astype(np.float64) * 300.0
    hb = H.make_host_buffers(precision, pos, vel, masses, forces)
    npad = hb.posq.shape[0]
    pool = np.zeros((2 * npad, 4), dtype=np.float32)
    pool[: npad + 64] = H.gaussian_pool(47, npad + 64)
    ri = 29
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    integ = s2.MmvtLangevinMiddleIntegrator(300.0, 1.0, 0.002, "")
    integ.setStepMode("two_pass")
    ctx = s2.Context(system, integ, {"CudaPrecision": NAMES[precision]})
    H.upload(ctx, [hb])
    ctx.setRandomPool(pool)
    integ._update_params()
    ctx.ensure_two_pass_buffers(True, True)
    vs, fs, ns = [C.c_double() for _ in range(3)], None, None
    out = (C.c_double * 3)()
    L.check(L.lib.seekr2_b200_get_params(integ._handle, out))
    dev = cuda_device
    t = lambda a: torch.from_numpy(a.copy()).to(dev)
    posq, velm, force, rnd = t(hb.posq), t(hb.velm), t(hb.force), t(pool)
    corr = t(hb.corr) if hb.corr is not None else None
    delta, old_delta = torch.zeros((npad, 4), dtype=velm.dtype, device=dev), torch.zeros((npad, 4), dtype=velm.dtype, device=dev)
    old_posq, old_velm = torch.zeros_like(posq), torch.zeros_like(velm)
    params = torch.tensor([out[0], out[2]], dtype=velm.dtype, device=dev)
    dt = torch.tensor([0.0, 0.002], dtype=velm.dtype, device=dev)
    P = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
    eq = lambda a, b, what: np.testing.assert_array_equal(a.cpu().numpy()[:n].view(np.uint8), b.cpu().numpy()[:n].view(np.uint8), err_msg=what)
    bufs = ctx.buffers()
    ref.ref_middle_part1(n, npad, P(velm), P(force), P(dt), 0, None)
    L.check(L.lib.seekr2_b200_kick(integ._handle, C.byref(bufs), ri, ctx.stream_ptr))
    ctx.stream.synchronize(); torch.cuda.synchronize()
    eq(ctx.velm[0], velm, "velm after Part1")
    ref.ref_middle_part2(n, P(velm), P(delta), P(old_delta), P(params), P(dt), P(old_velm), P(rnd), C.c_uint(ri), 0, None)
    L.check(L.lib.seekr2_b200_middle_drift(integ._handle, C.byref(bufs), ri, ctx.stream_ptr))
    ctx.stream.synchronize(); torch.cuda.synchronize()
    eq(ctx.velm[0], velm, "velm after Part2"); eq(ctx.pos_delta[0], delta, "posDelta"); eq(ctx.old_delta[0], old_delta, "oldDelta")
    eq(ctx.old_velm[0], old_velm, "oldVelm")
    # emulate a constraint solver: perturb posDelta identically on both sides
    bump = torch.zeros_like(delta); bump[:n:7, 0] = 1e-4
    delta += bump; ctx.pos_delta[0] += bump
    ref.ref_middle_part3(n, P(posq), P(velm), P(delta), P(old_delta), P(dt), P(old_posq), P(corr), 0, None)
    L.check(L.lib.seekr2_b200_finish(integ._handle, C.byref(bufs), ctx.stream_ptr))
    ctx.stream.synchronize(); torch.cuda.synchronize()
    eq(ctx.velm[0], velm, "velm after Part3"); eq(ctx.posq[0], posq, "posq after Part3")
    if corr is not None:
        eq(ctx.posq_correction[0], corr, "posqCorrection after Part3")


@pytest.mark.parametrize("variant", ["cuda", "reference"])
def test_middle_multi_pass_trajectory_vs_oracle(cuda_device, variant):
    syn = H.toy_system()
    pool = H.gaussian_pool(H.SEED_BASE + 73, 8000 * 32)
    rc, o, ob = S.run_oracle(syn, H.make_mmvt_integrator(syn, "", variant, middle=True), L.MIXED, 8000, pool)
    integ = H.make_mmvt_integrator(syn, "", variant, middle=True)
    grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, 8000, pool, mode="two_pass")
    assert rc == grc == 0 and len(o.records()) > 3
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))
    H.assert_state_equal(ob, gb, 1)
    assert H.stats_tuple(o.state(), 2) == H.stats_tuple(integ.getAnchorState(0), 2)
