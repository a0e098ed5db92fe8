"""Latency-regime sweep of the fused step: us per step of MMVT / Elber / plain Langevin over anchors per GPU, decision mode and
the LOCAL kernel's knobs (streaming threads per block, early normals, resident blocks), on the trypsin-size system (23 036 atoms),
the host-guest size system (2 520 atoms, 147 + 15 atom groups) and a 64-atom system (the launch floor).  Every engine reads its
knobs from the environment at create time, so one process sweeps them all.

    python tools/latency_sweep.py > gpurun_out/latency_sweep.jsonl        (SWEEP=quick|full, SWEEP_ANCHORS=1,2,4,...)
"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import seekr2_openmm_plugin_b200 as s2
from seekr2_openmm_plugin_b200 import synthetic as H

STEPS, G = int(os.environ.get("SWEEP_STEPS", "4096")), bench.STEPS_PER_GRAPH
KNOBS = ("SEEKR2_B200_LOCAL_THREADS", "SEEKR2_B200_TUNE", "SEEKR2_B200_PACKED", "SEEKR2_B200_LOCAL_REGS", "SEEKR2_B200_LOCAL_RESIDENT", "SEEKR2_B200_PDL")
_cache = {}


def workload(system, A):
    key = (system, A)
    if key not in _cache:
        if system == "tryp":
            syn, positions, forces, _ = bench.build_workload(A, bench.N_ATOMS, 0x5EEC2004)
            dt = bench.DT_PS
        else:
            n = 2520 if system == "hg" else 64
            host, guest = (list(range(0, 147)), list(range(147, 162))) if system == "hg" else (list(range(0, 11)), list(range(11, 20)))
            rin, rout = (0.3, 0.5) if system == "hg" else (1.1, 1.3)
            syn = H.pair_sphere_system(n, host, guest, rin, rout, 0x5EEC2002, dt=0.002, tether_k=0.0)
            H.place_pair_distance(syn, host, guest, 0.5 * (rin + rout))
            positions, forces, dt = np.broadcast_to(syn.positions, (A,) + syn.positions.shape).copy(), None, 0.002
        _cache[key] = (syn, positions, forces, dt)
    return _cache[key]


def us_per_step(system, A, kind, mode, knobs):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update({k: str(v) for k, v in knobs.items()})
    syn, positions, forces, dt = workload(system, A)
    if kind == "mmvt":
        integ = H.make_mmvt_integrator(syn, "")
    elif kind == "elber":
        integ = s2.ElberLangevinIntegrator(300.0, 1.0, dt, "")
        integ.addSrcMilestoneGroup(1)
        integ.setEndOnSrcMilestone(False)
    else:
        integ = s2.LangevinIntegrator(300.0, 1.0, dt)
    integ.setRandomNumberSeed(7)
    integ.setDecisionMode(mode)
    integ.ringCapacity = 1 << 14
    integ.setUseGraph(True, G)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A, randomPoolSteps=G)
    ctx.setPositions(positions); ctx.setVelocities(syn.velocities)
    if forces is not None:
        f = np.zeros((3, ctx.padded_num_atoms), dtype=np.int64)
        f[:, :forces.shape[0]] = np.rint(forces.T * 4294967296.0).astype(np.int64)
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
    assert integ.getDecisionMode() == mode
    bounces = sum(integ.getAnchorState(a).num_bounce_steps for a in range(A)) if kind == "mmvt" else 0
    integ.cleanup()
    return round(best, 3), bounces


full = os.environ.get("SWEEP", "quick") == "full"
tunes = [int(t) for t in os.environ.get("SWEEP_TUNES", "0").split(",")]
extra = [dict(kv.split("=") for kv in v.split("+") if kv) for v in os.environ.get("SWEEP_EXTRA", "").split(",")]   # e.g. SEEKR2_B200_SEGMENTED=0,SEEKR2_B200_CHAIN=0
threads = [int(t) for t in os.environ.get("SWEEP_THREADS", "64,128,192,256" if full else "128,256").split(",")]
variants = [("local", {"SEEKR2_B200_LOCAL_THREADS": T, "SEEKR2_B200_TUNE": t, **x}) for T in threads for t in tunes for x in extra]
if os.environ.get("SWEEP_HANDOFF", "1") == "1":
    variants += [("handoff", {"SEEKR2_B200_TUNE": 0})]
if full:
    variants += [("local", {"SEEKR2_B200_LOCAL_THREADS": 128, "SEEKR2_B200_PACKED": 0})]
anchors = [int(a) for a in os.environ.get("SWEEP_ANCHORS", "1,2,4,8,16,32").split(",")]
for system, kinds, alist in (("tryp", ("mmvt", "langevin"), anchors), ("hg", ("mmvt", "elber", "langevin"), [1]), ("tiny", ("mmvt", "langevin"), [1])):
    if system not in os.environ.get("SWEEP_SYSTEMS", "tryp,hg,tiny").split(","):
        continue
    for A in alist:
        rows = []
        for mode, knobs in variants:
            row = {"system": system, "anchors": A, "mode": mode, **{k.replace("SEEKR2_B200_", "").lower(): v for k, v in knobs.items()}}
            for kind in kinds:
                row[f"{kind}_us"], b = us_per_step(system, A, kind, mode, knobs)
                if kind == "mmvt":
                    row["bounce_steps"] = b
            rows.append(row)
            print(json.dumps(row), flush=True)
        plain = min(r["langevin_us"] for r in rows if "langevin_us" in r)
        best = min(rows, key=lambda r: r.get("mmvt_us", 1e9))
        print(json.dumps({"system": system, "anchors": A, "best_plain_us": plain, "best_mmvt_us": best["mmvt_us"], "ratio": round(best["mmvt_us"] / plain, 3),
                          "best_variant": {k: v for k, v in best.items() if k not in ("system", "anchors")}}), flush=True)
