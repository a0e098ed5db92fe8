"""Time per step of the fused MMVT step in both decision modes (LOCAL: every block decides; HANDOFF: one decide block per
anchor publishes a flag) and of the plain Langevin step, over anchors per GPU -- picks the auto-mode threshold.
    python tools/decide_mode_sweep.py [atoms] > gpurun_out/decide_mode_sweep.json"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import seekr2_openmm_plugin_b200 as s2
from seekr2_openmm_plugin_b200 import synthetic as H

NATOMS = int(sys.argv[1]) if len(sys.argv) > 1 else bench.N_ATOMS
STEPS, G = int(os.environ.get("SWEEP_STEPS", "4096")), bench.STEPS_PER_GRAPH


def us_per_step(A, kind, mode):
    syn, positions, forces, _ = bench.build_workload(A, NATOMS, 0x5EEC2004)
    if kind == "mmvt":
        integ = H.make_mmvt_integrator(syn, "")
    elif kind == "elber":                                    # slot role in the anchor's first block only, no undo log, never bounces
        integ = s2.ElberLangevinIntegrator(300.0, 1.0, bench.DT_PS, "")
        integ.addSrcMilestoneGroup(1)
        integ.setEndOnSrcMilestone(False)
    else:
        integ = s2.LangevinIntegrator(300.0, 1.0, bench.DT_PS)
    integ.setRandomNumberSeed(7)
    integ.setDecisionMode(mode)
    integ.ringCapacity = 1 << 14
    integ.setUseGraph(True, G)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A, randomPoolSteps=G)
    ctx.setPositions(positions); ctx.setVelocities(syn.velocities)
    f = np.zeros((3, ctx.padded_num_atoms), dtype=np.int64)
    f[:, :NATOMS] = np.rint(forces.T * 4294967296.0).astype(np.int64)
    ctx.force.copy_(torch.from_numpy(f).cuda()[None].expand(A, -1, -1))
    integ.step(4 * G)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(ctx.stream):
            e0.record(ctx.stream)
            integ.step(STEPS)
            e1.record(ctx.stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / STEPS)
    assert mode == "auto" or integ.getDecisionMode() == mode
    return best, integ.getDecisionMode()


KINDS = os.environ.get("SWEEP_KINDS", "mmvt,langevin").split(",")
out = {"atoms": NATOMS, "steps": STEPS, "steps_per_graph": G, "rows": []}
for A in [int(a) for a in os.environ.get("SWEEP_ANCHORS", "1,2,4,6,8,12,16,24,32,48,64,96,128").split(",")]:
    row = {"anchors": A}
    for kind in KINDS:
        for mode in os.environ.get("SWEEP_MODES", "auto,local,handoff").split(","):
            t, chosen = us_per_step(A, kind, mode)
            row[f"{kind}_{mode}_us"] = round(t, 3)
            if mode == "auto" and kind == "mmvt":
                row["auto_mode"] = chosen
    modes = os.environ.get("SWEEP_MODES", "auto,local,handoff").split(",")
    if "langevin" not in KINDS or "mmvt" not in KINDS:
        out["rows"].append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
        continue
    plain = min(row[f"langevin_{m}_us"] for m in modes)      # the fastest plain Langevin step, whatever launch shape
    for m in modes:
        row[f"mmvt_over_plain_{m}"] = round(row[f"mmvt_{m}_us"] / plain, 3)
    out["rows"].append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
print(json.dumps(out))
