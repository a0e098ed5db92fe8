"""Measures the decide -> streaming hand-off inside fused_step_kernel with the debug probe."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
from helpers import L, s2

A = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NATOMS = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_ATOMS
syn, positions, forces, _ = bench.build_workload(A, NATOMS, 1)
integ = H.make_mmvt_integrator(syn, "")
if os.environ.get("PROBE_GRAPH"): integ.setUseGraph(True, 32)
if len(sys.argv) > 3: integ.setDecisionMode(sys.argv[3])
ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
ctx.setPositions(positions); ctx.setVelocities(syn.velocities)
integ.step(8)
probe = torch.zeros(8 * A + 4 + 32, dtype=torch.int64, device="cuda")
big = torch.iinfo(torch.int64).max
for it in range(3):
    probe.zero_(); probe[6:8 * A:8] = big
    torch.cuda.synchronize()
    L.check(L.lib.seekr2_b200_debug_set_probe(integ._handle, probe.data_ptr()))
    integ.step(int(os.environ.get('PROBE_STEPS', '1')))      # > 1: the stamps of the last launch of a warm sequence
    torch.cuda.synchronize()
    p = probe.cpu().numpy()
    d = p[:8 * A].reshape(A, 8).astype(np.float64)
    t0 = d[:, 0].min()
    names = ["start", "loads returned", "advanced", "accumulated", "reduced", "published", "first streaming arrival", "gathers issued"]
    print(f"iter {it}: " + "; ".join(f"{n} {np.median(d[:, i] - t0):.0f}" for i, n in enumerate(names)) + " (median ns over anchors)")
    print(f"   streaming blocks that waited {p[8*A]}, polls {p[8*A+1]}, total wait {p[8*A+2]/1e3:.1f} us "
          f"(mean {p[8*A+2]/max(1,p[8*A]):.0f} ns per waiting block)")
    cyc = p[8 * A + 4: 8 * A + 4 + 11]
    if integ.getDecisionMode() == "local":
        stages = ["round 1 issued", "force gather issued (index back)", "normals done", "slot atom advanced", "own atom advanced",
                  "sums reduced (slot-warp barrier)", "boundary terms", "decision (streaming: barrier passed)", "own stores issued", "slab write-back issued"]
        for who, off in (("block 0, slot thread 0 (leader)", 0), ("block 1, streaming thread 128", 16)):
            c = p[8 * A + 4 + off: 8 * A + 4 + off + 11]
            print(f"   LOCAL mode, {who}, SM cycles since block start: " + "; ".join(f"{n} {int(c[i + 1] - c[0])}" for i, n in enumerate(stages) if c[i + 1] > 0))
        continue
    stages = ["step index loaded", "gathers issued", "gathers back", "cleared+sync", "advanced", "accumulated", "all warps reduced",
              "boundary terms", "published", "bookkeeping+slab+clock"]
    print("   anchor 0 decide block, SM cycles since block start: " + "; ".join(f"{n} {int(cyc[i + 1] - cyc[0])}" for i, n in enumerate(stages)))
L.check(L.lib.seekr2_b200_debug_set_probe(integ._handle, None))
