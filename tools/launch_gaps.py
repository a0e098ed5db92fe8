"""Launch overhead of the fused step (SURVEY.md 8d "Evidence"; no nsys in this image, so CUDA events do the job):
  - us per step of back-to-back EAGER launches (seekr2_b200_step, one cudaLaunchKernelEx each) vs replays of a 32-step CUDA graph;
  - the gap between consecutive graph replays (events around every replay: total - sum of the replays);
  - the kernel's own duration from the committed ncu launch list, when given, to split a graph step into kernel and gap.
    python tools/launch_gaps.py > gpurun_out/launch_gaps.json"""
import ctypes as C, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import seekr2_openmm_plugin_b200 as s2
from seekr2_openmm_plugin_b200 import _lib as L, synthetic as H

G, GRAPHS = 32, 128
out = {"steps_per_graph": G, "graph_replays": GRAPHS, "rows": []}
for label, A, n_atoms in (("1 anchor x 23 036 atoms (LOCAL kernel)", 1, bench.N_ATOMS), ("1 anchor x 64 atoms (launch floor)", 1, 64),
                          ("128 anchors x 23 036 atoms (hand-off kernel)", 128, bench.N_ATOMS)):
    for kind in ("mmvt", "langevin"):
        syn, positions, forces, _ = bench.build_workload(A, n_atoms, 0x5EEC2004)
        integ = H.make_mmvt_integrator(syn, "") if kind == "mmvt" else s2.LangevinIntegrator(300.0, 1.0, bench.DT_PS)
        integ.setRandomNumberSeed(7)
        integ.ringCapacity = 1 << 14
        integ.setUseGraph(True, G)
        ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A, randomPoolSteps=G)
        ctx.setPositions(positions); ctx.setVelocities(syn.velocities)
        integ.step(4 * G)                                   # captures and warms the graph
        torch.cuda.synchronize()
        stream = ctx.stream
        with torch.cuda.stream(stream):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * GRAPHS + 2)]
            ev[0].record(stream)
            for g in range(GRAPHS):
                ev[1 + 2 * g].record(stream)
                L.check(L.lib.seekr2_b200_graph_launch(integ._handle, ctx.stream_ptr))
                ev[2 + 2 * g].record(stream)
            ev[-1].record(stream)
        torch.cuda.synchronize()
        total = ev[0].elapsed_time(ev[-1]) * 1e3
        inside = [ev[1 + 2 * g].elapsed_time(ev[2 + 2 * g]) * 1e3 for g in range(GRAPHS)]
        graph_us_per_step = float(np.median(inside)) / G
        gap_between_replays = (total - sum(inside)) / GRAPHS
        integ._steps_issued += G * GRAPHS
        # eager: the same number of steps as individual launches (programmatic dependent launch still applies)
        bufs = ctx.buffers()
        with torch.cuda.stream(stream):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(64):
                L.check(L.lib.seekr2_b200_step(integ._handle, C.byref(bufs), 0, ctx.stream_ptr))
            a.record(stream)
            for _ in range(G * GRAPHS // 4):
                L.check(L.lib.seekr2_b200_step(integ._handle, C.byref(bufs), 0, ctx.stream_ptr))
            b.record(stream)
        torch.cuda.synchronize()
        eager_us_per_step = a.elapsed_time(b) * 1e3 / (G * GRAPHS // 4)
        out["rows"].append({"workload": label, "kind": kind, "decision_mode": integ.getDecisionMode() if kind == "mmvt" else None,
                            "graph_us_per_step": round(graph_us_per_step, 3), "gap_between_graph_replays_us": round(gap_between_replays, 3),
                            "eager_us_per_step": round(eager_us_per_step, 3)})
        integ.cleanup()
print(json.dumps(out, indent=1))
