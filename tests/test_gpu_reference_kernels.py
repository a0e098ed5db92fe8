"""The reference's OWN CUDA kernels (langevin.cu, compiled unmodified from /root/reference into
oracle/_ref/ by oracle/build_ref.sh) against (a) the CPU oracle's restatement and (b) the B200 two-pass
kernels, on identical inputs: bit-exact in all three CudaPrecision modes.  This is what pins the
Langevin arithmetic of the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from helpers import L, s2

pytestmark = pytest.mark.gpu

NAMES = {L.SINGLE: "single", L.MIXED: "mixed", L.DOUBLE: "double"}


def _ref_lib(precision, fastmath=False):
    path = os.path.join(H.oracle_binding.REF_DIR, f"libref_langevin_{NAMES[precision]}{'_fastmath' if fastmath else ''}.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: the reference kernels (oracle/build_ref.sh, needs /root/reference) must be built before the GPU suite runs -- a skipped pin is not a green pin")
    lib = C.CDLL(path)
    assert lib.ref_precision() == precision
    return lib


def _inputs(precision, n, seed):
    pos, vel, masses = H.lattice_system(n, seed)
    masses[::17] = 0.0                      # some immobile particles
    forces = H.gaussian_pool(seed + 7, n)[:, :3].astype(np.float64) * 300.0
    hb = H.make_host_buffers(precision, pos, vel, masses, forces)
    pool = H.gaussian_pool(seed + 11, hb.posq.shape[0] + 64)
    return hb, pool, masses


def _plain(n, masses, precision, kind=L.KIND_MMVT):
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    integ = s2.MmvtLangevinIntegrator(300.0, 1.0, 0.002, "")
    return system, integ


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
@pytest.mark.parametrize("n,blocks", [(1, 0), (777, 3), (5000, 0)])
def test_reference_kernels_vs_oracle_and_b200(cuda_device, precision, n, blocks):
    import torch
    ref = _ref_lib(precision)
    hb, pool, masses = _inputs(precision, n, 100 + n)
    npad = hb.posq.shape[0]
    system, integ = _plain(n, masses, precision)
    o = H.oracle_for(integ, system, precision)
    vs, fs, ns = o.params()
    mixed = H.mixed_dtype(precision)
    ri = 37

    # --- reference kernels on the GPU
    dev = cuda_device
    t = lambda a: torch.from_numpy(a.copy()).to(dev)
    posq, velm, force, rnd = t(hb.posq), t(hb.velm), t(hb.force), t(pool)
    corr = t(hb.corr) if hb.corr is not None else None
    delta = torch.zeros((npad, 4), dtype=velm.dtype, device=dev)
    old_posq, old_velm = torch.zeros_like(posq), torch.zeros_like(velm)
    params = torch.tensor([vs, fs, ns], dtype=velm.dtype, device=dev)
    dt = torch.tensor([0.0, 0.002], dtype=velm.dtype, device=dev)
    P = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
    assert ref.ref_part1(n, npad, P(velm), P(force), P(delta), P(params), P(dt), P(rnd), C.c_uint(ri), blocks, None) == 0
    torch.cuda.synchronize()
    ref_velm1, ref_delta = velm.cpu().numpy(), delta.cpu().numpy()
    assert ref.ref_part2_mmvt(n, P(posq), P(corr), P(delta), P(velm), P(dt), P(old_posq), P(old_velm), blocks, None) == 0
    torch.cuda.synchronize()
    ref_after = H.HostBuffers(precision, posq.cpu().numpy(), None if corr is None else corr.cpu().numpy(), velm.cpu().numpy(), hb.force)
    assert ref.ref_bounce(n, npad, P(posq), P(velm), P(old_posq), P(old_velm), blocks, None) == 0
    torch.cuda.synchronize()
    ref_bounced = H.HostBuffers(precision, posq.cpu().numpy(), None if corr is None else corr.cpu().numpy(), velm.cpu().numpy(), hb.force)

    # --- CPU oracle
    ob = hb.copy()
    odelta = np.zeros((npad, 4), dtype=mixed)
    oold_posq, oold_velm = np.zeros_like(ob.posq), np.zeros_like(ob.velm)
    o.lib.seekr2_oracle_part1(o.h, H.oracle_binding.ptr(ob.velm), H.oracle_binding.ptr(ob.force), H.oracle_binding.ptr(odelta),
                              H.oracle_binding.ptr(pool), ri)
    np.testing.assert_array_equal(ob.velm[:n].view(np.uint8), ref_velm1[:n].view(np.uint8))
    np.testing.assert_array_equal(odelta[:n].view(np.uint8), ref_delta[:n].view(np.uint8))
    o.lib.seekr2_oracle_part2(o.h, H.oracle_binding.ptr(ob.posq), H.oracle_binding.ptr(ob.corr), H.oracle_binding.ptr(odelta),
                              H.oracle_binding.ptr(ob.velm), H.oracle_binding.ptr(oold_posq), H.oracle_binding.ptr(oold_velm))
    H.assert_state_equal(ob, ref_after, n, "oracle vs reference kernels after Part2")
    o.lib.seekr2_oracle_bounce(o.h, H.oracle_binding.ptr(ob.posq), H.oracle_binding.ptr(ob.velm), H.oracle_binding.ptr(oold_posq),
                               H.oracle_binding.ptr(oold_velm))
    H.assert_state_equal(ob, ref_bounced, n, "oracle vs reference kernels after mmvtBounce")

    # --- B200 two-pass kernels (kick + finish), no boundaries => never bounces
    integ2 = s2.MmvtLangevinIntegrator(300.0, 1.0, 0.002, "")
    integ2.setStepMode("two_pass")
    ctx = s2.Context(system, integ2, {"CudaPrecision": NAMES[precision]})
    H.upload(ctx, [hb])
    padded_pool = np.zeros(((ri + npad + 31) // npad * npad + npad, 4), dtype=np.float32)
    padded_pool[: pool.shape[0]] = pool
    ctx.setRandomPool(padded_pool)
    integ2._update_params()
    ctx.ensure_two_pass_buffers(False)
    bufs = ctx.buffers()
    L.check(L.lib.seekr2_b200_kick(integ2._handle, C.byref(bufs), ri, ctx.stream_ptr))
    ctx.stream.synchronize()
    np.testing.assert_array_equal(ctx.velm[0].cpu().numpy()[:n].view(np.uint8), ref_velm1[:n].view(np.uint8))
    np.testing.assert_array_equal(ctx.pos_delta[0].cpu().numpy()[:n].view(np.uint8), ref_delta[:n].view(np.uint8))
    L.check(L.lib.seekr2_b200_finish(integ2._handle, C.byref(bufs), ctx.stream_ptr))
    H.assert_state_equal(H.download(ctx), ref_after, n, "B200 kick+finish vs reference kernels")

    # --- B200 fused single-launch step
    integ3 = s2.MmvtLangevinIntegrator(300.0, 1.0, 0.002, "")
    ctx3 = s2.Context(system, integ3, {"CudaPrecision": NAMES[precision]})
    H.upload(ctx3, [hb])
    ctx3.setRandomPool(padded_pool)
    integ3._update_params()
    L.check(L.lib.seekr2_b200_step(integ3._handle, C.byref(ctx3.buffers()), ri, ctx3.stream_ptr))
    H.assert_state_equal(H.download(ctx3), ref_after, n, "B200 fused step vs reference kernels")


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED])
def test_fast_math_build_within_tolerance(cuda_device, precision):
    """OpenMM hands NVRTC --use_fast_math; the reference kernels built that way differ from the IEEE
    build only through sqrtf.approx: positions / velocities agree to 1e-5 relative (north star)."""
    import torch
    n = 4096
    hb, pool, masses = _inputs(precision, n, 5)
    npad = hb.posq.shape[0]
    out = []
    for fast in (False, True):
        ref = _ref_lib(precision, fast)
        t = lambda a: torch.from_numpy(a.copy()).to(cuda_device)
        posq, velm, force, rnd = t(hb.posq), t(hb.velm), t(hb.force), t(pool)
        corr = t(hb.corr) if hb.corr is not None else None
        delta = torch.zeros((npad, 4), dtype=velm.dtype, device=cuda_device)
        old_posq, old_velm = torch.zeros_like(posq), torch.zeros_like(velm)
        vs, fs, ns = np.exp(-0.002), (1 - np.exp(-0.002)), np.sqrt(s2.BOLTZ * 300 * (1 - np.exp(-0.004)))
        params = torch.tensor([vs, fs, ns], dtype=velm.dtype, device=cuda_device)
        dt = torch.tensor([0.0, 0.002], dtype=velm.dtype, device=cuda_device)
        P = lambda x: C.c_void_p(x.data_ptr()) if x is not None else None
        ref.ref_part1(n, npad, P(velm), P(force), P(delta), P(params), P(dt), P(rnd), C.c_uint(0), 0, None)
        ref.ref_part2_mmvt(n, P(posq), P(corr), P(delta), P(velm), P(dt), P(old_posq), P(old_velm), 0, None)
        torch.cuda.synchronize()
        out.append((posq.cpu().numpy()[:n, :3].astype(np.float64), velm.cpu().numpy()[:n, :3].astype(np.float64)))
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-5)
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-5, atol=1e-7)
