"""Measures the decide -> streaming hand-off inside the fused kernels with the debug probe.  Needs the probe build:
    make -C seekr2_openmm_plugin_b200/csrc probe
    SEEKR2_B200_LIBRARY=seekr2_openmm_plugin_b200/libseekr2_b200_probe.so python tools/handoff_probe.py 1 23036 local"""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench, helpers as H
from helpers import L, s2

A = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NATOMS = int(sys.argv[2]) if len(sys.argv) > 2 else bench.N_ATOMS
syn, positions, forces, _ = bench.build_workload(A, NATOMS, 1)
if os.environ.get("PROBE_HG"):      # the bench's host-guest leg: 2 520 atoms, 147 + 15 atom groups (aligned slot layout)
    host, guest = list(range(0, 147)), list(range(147, 162))
    syn = H.pair_sphere_system(2520, host, guest, 0.3, 0.5, 3, dt=0.002, tether_k=0.0)
    H.place_pair_distance(syn, host, guest, 0.4)
    positions = syn.positions
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
        slot = ["wait returned", "gathers issued", "slot atom advanced", "sums reduced", "terms evaluated", "published (arrived)", "slab + clock done"]
        strm = ["wait returned", "loads issued", "", "own atom advanced", "optimistic stores issued", "barrier passed", "end"]
        for who, off, stages in (("block 0, slot thread 0 (leader)", 0, slot), ("last block, streaming thread T/2", 16, strm)):
            c = p[8 * A + 4 + off: 8 * A + 4 + off + 11]
            print(f"   LOCAL mode, {who}, SM cycles since block start: " + "; ".join(f"{n} {int(c[i + 1] - c[0])}" for i, n in enumerate(stages) if c[i + 1] > 0)
                  + (f"; [loads returned {int(c[8] - c[0])}]" if c[8] > 0 else "") + (f"; [prologue done, before the wait {int(c[9] - c[0])}]" if c[9] > 0 else ""))
        continue
    stages = ["step index loaded", "gathers issued", "gathers back", "cleared+sync", "advanced", "accumulated", "all warps reduced",
              "boundary terms", "published", "bookkeeping+slab+clock"]
    print("   anchor 0 decide block, SM cycles since block start: " + "; ".join(f"{n} {int(cyc[i + 1] - cyc[0])}" for i, n in enumerate(stages)))
L.check(L.lib.seekr2_b200_debug_set_probe(integ._handle, None))
