"""Per-kernel bandwidth of every kernel on the step path (CUDA events, 128 anchors x 23 036 atoms, mixed
precision, working set > L2): fused step (Langevin / Middle, MMVT / Elber / plain) and the multi-pass
kernels the OpenMM adapter uses.  Prints one line per kernel: us per call, algorithmic GB/s, fraction of
the measured HBM copy bandwidth."""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import seekr2_openmm_plugin_b200 as s2
from seekr2_openmm_plugin_b200 import _lib as L, synthetic as H

A = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = bench.N_ATOMS
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
syn, positions, forces, _ = bench.build_workload(A, N, 1)
f = np.zeros((3, (N + 31) // 32 * 32), dtype=np.int64); f[:, :N] = np.rint(forces.T * 4294967296.0).astype(np.int64)


def context(cls, middle=False, mode="fused", pool=False):
    if cls == "mmvt":
        integ = H.make_mmvt_integrator(syn, "", middle=middle)
    elif cls == "elber":
        integ = (s2.ElberLangevinMiddleIntegrator if middle else s2.ElberLangevinIntegrator)(300.0, 1.0, bench.DT_PS, "")
        integ.addSrcMilestoneGroup(1); integ.addDestMilestoneGroup(1)
    else:
        integ = (s2.LangevinMiddleIntegrator if middle else s2.LangevinIntegrator)(300.0, 1.0, bench.DT_PS)
    integ.setStepMode(mode); integ.useRandomPool = pool; integ.ringCapacity = 1 << 14
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A, randomPoolSteps=2)
    ctx.setPositions(positions); ctx.setVelocities(syn.velocities)
    ctx.force.copy_(torch.from_numpy(f).cuda()[None].expand(A, -1, -1))
    integ.step(4)
    return ctx, integ


def timeit(fn, ctx, reps=40):
    for _ in range(4):
        fn()
    ctx.stream.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.stream):
        a.record(ctx.stream)
        for _ in range(reps):
            fn()
        b.record(ctx.stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


rows = []
def report(name, us, bytes_per_atom):
    gbs = A * N * bytes_per_atom / (us * 1e-6) / 1e9
    rows.append((name, us, bytes_per_atom, gbs, gbs / peak))
    print(f"{name:58s} {us:8.2f} us  {bytes_per_atom:4d} B/atom  {gbs:8.1f} GB/s  {gbs / peak:5.3f} of measured HBM peak", flush=True)

for cls in ("langevin", "mmvt", "elber"):
    for middle in (False, True):
        ctx, integ = context(cls, middle)
        bufs = ctx.buffers()
        def step():
            L.check(L.lib.seekr2_b200_step(integ._handle, C.byref(bufs), 0, ctx.stream_ptr))
        report(f"fused_step_kernel<MIXED,{cls.upper()},{'MIDDLE' if middle else 'LANGEVIN'}> (in-kernel RNG)", timeit(step, ctx), 152)
        del ctx, integ
ctx, integ = context("mmvt", False, pool=True)
bufs = ctx.buffers()
report("fused_step_kernel<MIXED,MMVT,LANGEVIN> (Gaussian pool)", timeit(lambda: L.check(L.lib.seekr2_b200_step(integ._handle, C.byref(bufs), 0, ctx.stream_ptr)), ctx), 168)
del ctx, integ
for middle in (False, True):
    ctx, integ = context("mmvt", middle, mode="two_pass", pool=True)
    ctx.ensure_two_pass_buffers(middle, middle)
    bufs = ctx.buffers()
    tag = "MIDDLE" if middle else "LANGEVIN"
    if not middle:
        report("kick_kernel (Part1; pool)", timeit(lambda: L.check(L.lib.seekr2_b200_kick(integ._handle, C.byref(bufs), 0, ctx.stream_ptr)), ctx), 136)
        report("finish_kernel, one launch: decide blocks + Part2 + bounce", timeit(lambda: L.check(L.lib.seekr2_b200_finish(integ._handle, C.byref(bufs), ctx.stream_ptr)), ctx), 160)
    else:
        report("middle_part1_kernel (force kick)", timeit(lambda: L.check(L.lib.seekr2_b200_kick(integ._handle, C.byref(bufs), 0, ctx.stream_ptr)), ctx), 88)
        report("middle_part2_kernel (O step, posDelta, oldDelta, oldVelm; pool)", timeit(lambda: L.check(L.lib.seekr2_b200_middle_drift(integ._handle, C.byref(bufs), 0, ctx.stream_ptr)), ctx), 176)
        report("middle_finish_kernel, one launch: decide blocks + Part3 + bounce", timeit(lambda: L.check(L.lib.seekr2_b200_finish(integ._handle, C.byref(bufs), ctx.stream_ptr)), ctx), 192)
    if not middle:
        # the two small launches the OpenMM adapter adds around a fused step: slab re-gather and (on crossing steps) the ring drain
        us = timeit(lambda: L.check(L.lib.seekr2_b200_resync(integ._handle, C.byref(bufs), ctx.stream_ptr)), ctx)
        print(f"{'resync_kernel (slab re-gather + step index, 128 anchors x 64 slots)':58s} {us:8.2f} us", flush=True)
        def drain():
            L.check(L.lib.seekr2_b200_drain_begin(integ._handle, 4096, ctx.stream_ptr))
            n = C.c_int64(0)
            L.check(L.lib.seekr2_b200_drain_end(integ._handle, (L.Record * 4096)(), 4096, C.byref(n), None))
        us = timeit(drain, ctx, reps=20)
        print(f"{'ring_pack_kernel + event wait (blocking drain, nothing to fetch)':58s} {us:8.2f} us", flush=True)
    del ctx, integ
json.dump([dict(kernel=r[0], us=r[1], bytes_per_atom=r[2], gbs=r[3], frac=r[4]) for r in rows], open(os.path.join(ROOT, "gpurun_out", "kernel_bench.json"), "w"), indent=1)
