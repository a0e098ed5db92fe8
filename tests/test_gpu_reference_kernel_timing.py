"""GPU-vs-GPU baseline: the reference's OWN integrator kernels (langevin.cu Part1 + Part2, unmodified, recompiled for
sm_100a by oracle/build_ref.sh, launched as OpenMM's executeKernel would: 128-thread blocks, grid capped at
4 x SM count, grid-stride) against this repository's fused step, on the same number of atoms in the same layouts
(mixed precision, 128 x 23 036 atoms, Gaussian pool in both).  The reference additionally needs a generic
force-group energy pass + blocking read-back (+ mmvtBounce on crossing steps) for the CV, which is not timed here --
so the ratio below UNDERSTATES the reference's integrator-side cost per step.  Result goes to
gpurun_out/reference_kernel_timing.json (summary kept in profiles/)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import helpers as H
from helpers import L, s2

pytestmark = pytest.mark.gpu


def test_fused_step_beats_reference_part1_part2(cuda_device):
    import torch
    import bench
    path = os.path.join(H.oracle_binding.REF_DIR, "libref_langevin_mixed.so")
    if not os.path.exists(path):
        pytest.fail(f"{path} not built (oracle/build_ref.sh needs /root/reference)")
    ref = C.CDLL(path)
    A, N = 128, bench.N_ATOMS
    npad = (N + 31) // 32 * 32
    total = A * npad
    dev = cuda_device
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    peak = json.load(open(os.path.join(H.ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(H.ROOT, "MEASURED_PEAKS.json")) else 6545.3

    # ---- the reference kernels on one system of A*Np atoms
    g = torch.Generator(device=dev).manual_seed(1)
    posq = torch.rand((total, 4), device=dev, generator=g, dtype=torch.float32)
    corr = torch.zeros_like(posq)
    velm = torch.randn((total, 4), device=dev, generator=g, dtype=torch.float64)
    velm[:, 3] = 0.1
    force = torch.randint(-2 ** 34, 2 ** 34, (3 * total,), device=dev, generator=g, dtype=torch.int64)
    delta = torch.zeros((total, 4), device=dev, dtype=torch.float64)
    rnd = torch.randn((total, 4), device=dev, generator=g, dtype=torch.float32)
    old_posq, old_velm = torch.zeros_like(posq), torch.zeros_like(velm)
    params = torch.tensor([0.996, 0.00399, 0.1], dtype=torch.float64, device=dev)
    dt = torch.tensor([0.004, 0.004], dtype=torch.float64, device=dev)
    P = lambda x: C.c_void_p(x.data_ptr())
    blocks = 4 * sms

    def ref_step():
        assert ref.ref_part1(total, total, P(velm), P(force), P(delta), P(params), P(dt), P(rnd), C.c_uint(0), blocks, None) == 0
        assert ref.ref_part2_mmvt(total, P(posq), P(corr), P(delta), P(velm), P(dt), P(old_posq), P(old_velm), blocks, None) == 0

    def timed(fn, iters=40, warm=5):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / iters            # us

    ref_us = timed(ref_step)
    ref_uncapped_us = None
    blocks = 0                                               # one block per 128 atoms (not what OpenMM does; upper bound for the kernels)
    ref_uncapped_us = timed(ref_step)
    del posq, corr, velm, force, delta, rnd, old_posq, old_velm

    # ---- this repository's fused MMVT step on A anchors x N atoms, with and without a Gaussian pool
    syn, positions, forces, _ = bench.build_workload(A, N, 7)
    out = {}
    for label, use_pool in (("pool", True), ("in_kernel_rng", False)):
        integ = H.make_mmvt_integrator(syn, "")
        integ.setRandomNumberSeed(3)
        integ.useRandomPool = use_pool
        integ.ringCapacity = 1 << 14
        ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A, randomPoolSteps=8)
        ctx.setPositions(positions)
        ctx.setVelocities(syn.velocities)
        f = np.zeros((3, npad), dtype=np.int64)
        f[:, :N] = np.rint(forces.T * 4294967296.0).astype(np.int64)
        ctx.force.copy_(torch.from_numpy(f).to(dev)[None].expand(A, 3, npad))
        integ.step(16)
        torch.cuda.synchronize()
        with torch.cuda.stream(ctx.stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
            steps = 8 if use_pool else 64                    # stay inside one pool fill
            bufs = ctx.buffers()
            for s in range(steps):
                L.check(L.lib.seekr2_b200_step(integ._handle, C.byref(bufs), s * npad if use_pool else 0, ctx.stream_ptr))
            e1.record(ctx.stream)
        torch.cuda.synchronize()
        out[label] = e0.elapsed_time(e1) * 1e3 / steps
        del ctx, integ

    atoms = A * N
    result = {
        "workload": f"{A} x {N} atoms, mixed precision",
        "reference_part1_part2_us": ref_us, "reference_algorithmic_bytes_per_atom": 344,
        "reference_gbs": atoms * 344 / ref_us / 1e3, "reference_frac_of_hbm_peak": atoms * 344 / ref_us / 1e3 / peak,
        "reference_part1_part2_uncapped_grid_us": ref_uncapped_us,
        "fused_step_pool_us": out["pool"], "fused_step_in_kernel_rng_us": out["in_kernel_rng"],
        "speedup_vs_reference_kernels_pool": ref_us / out["pool"],
        "speedup_vs_reference_kernels_in_kernel_rng": ref_us / out["in_kernel_rng"],
        "note": "reference side excludes its per-step CV energy pass, blocking read-back and mmvtBounce",
    }
    os.makedirs(os.path.join(H.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(H.ROOT, "gpurun_out", "reference_kernel_timing.json"), "w") as fh:
        json.dump(result, fh, indent=1)
    print(json.dumps(result))
    assert out["pool"] < ref_us and out["in_kernel_rng"] < ref_us
