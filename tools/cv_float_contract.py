#!/usr/bin/env python
"""How often would OpenMM's SINGLE-precision evaluation of a boundary term decide differently from this repository's
double / fixed-point one?  ("Boundary-value contract".)

The reference reads the boundary force group's energy from OpenMM (CudaSeekr2Kernels.cpp:248); OpenMM's CUDA platform
evaluates CustomCentroidBondForce in `real` = float in single and mixed precision: float group centres (normalised float
weights, float accumulation), float `distance(g1,g2)` = sqrtf(dx*dx+dy*dy+dz*dz), float `distance^2 - radius^2`.  OpenMM is
absent here, so that evaluation is RESTATED below in numpy float32 (two accumulation orders, since the order of OpenMM's
block reduction is not something to rely on) and compared, step by step, with the oracle's evaluation of the very same
configuration along a thermal MMVT run that keeps bouncing between two surfaces 4 pm apart.  The configuration tested is the
advanced one (before a bounce restores the old positions): a plain-Langevin twin of the oracle produces it each step.

Runs on the CPU (oracle only).  `python tools/cv_float_contract.py [steps]` writes profiles/r02_cv_float_contract.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def float_decisions(pos32, groups, weights32, radii32, ks):
    """bit string of step(k * (distance(g1,g2)^2 - radius^2)) >= 0 terms, evaluated in float32; two accumulation orders"""
    out = []
    for order in ("sequential", "pairwise"):
        cen = []
        for atoms, w in zip(groups, weights32):
            terms = pos32[atoms] * w[:, None]                                   # float32 products
            cen.append(np.cumsum(terms, axis=0, dtype=np.float32)[-1] if order == "sequential" else terms.sum(axis=0, dtype=np.float32))
        d = cen[1] - cen[0]
        r2 = np.float32(np.float32(d[0] * d[0] + d[1] * d[1]) + d[2] * d[2])
        dist = np.sqrt(r2, dtype=np.float32)
        bits = 0
        us = []
        for b, (r, k) in enumerate(zip(radii32, ks)):
            u = np.float32(dist * dist) - np.float32(r * r)
            us.append(float(u))
            if np.float32(k) * u >= 0:
                bits |= 1 << b
        out.append((bits, us))
    return out


def run(name, n_atoms, host, guest, radius, steps, seed):
    import helpers as H
    from helpers import L, s2
    half = 0.002                                                                # shell 4 pm wide: a bounce every few tens of steps
    syn = H.pair_sphere_system(n_atoms, host, guest, radius - half, radius + half, seed, dt=0.002, tether_k=0.0)
    H.place_pair_distance(syn, host, guest, radius)
    npad = (n_atoms + 31) // 32 * 32
    chunk = 256
    pool = H.gaussian_pool(seed + 1, chunk * npad)
    mm = H.oracle_for(H.make_mmvt_integrator(syn, ""), syn.system, L.MIXED)
    twin = H.oracle_for(s2.LangevinIntegrator(300.0, 1.0, 0.002), syn.system, L.MIXED)
    hb = H.make_host_buffers(L.MIXED, syn.positions, syn.velocities, syn.masses)
    groups = [np.array(host), np.array(guest)]
    weights32 = [(syn.masses[g] / syn.masses[g].sum()).astype(np.float32) for g in groups]
    radii32 = [np.float32(radius - half), np.float32(radius + half)]
    ks = [-1.0, 1.0]
    mismatches = {"sequential": 0, "pairwise": 0}
    worst = {"sequential": 0.0, "pairwise": 0.0}
    near = 0
    for s in range(steps):
        ri = (s % chunk) * npad
        tb = hb.copy()
        twin.step(L.KIND_LANGEVIN, tb.posq, tb.corr, tb.velm, tb.force, pool, ri)      # the advanced configuration
        value = np.float32(mm.group_value(tb.posq, tb.corr, 1))
        bits64 = int(value) if value > 0 else 0
        for order, (bits32, us) in zip(("sequential", "pairwise"), float_decisions(tb.posq[:, :3], groups, weights32, radii32, ks)):
            if (bits32 > 0) != (bits64 > 0) or bits32 != bits64:
                mismatches[order] += 1
                worst[order] = max(worst[order], min(abs(u) for u in us))
        rc = mm.step(L.KIND_MMVT, hb.posq, hb.corr, hb.velm, hb.force, pool, ri)
        assert rc == 0
    st = mm.state()
    # float resolution of u = d^2 - r^2 at this radius: a few ulps of r^2
    ulp = float(np.spacing(np.float32(radius * radius)))
    return {"system": name, "atoms": n_atoms, "group_sizes": [len(host), len(guest)], "radius_nm": radius, "steps": steps,
            "bounce_steps": int(st.num_bounce_steps), "decisions_differing": mismatches,
            "rate_per_step": {k: v / steps for k, v in mismatches.items()},
            "rate_per_bounce": {k: v / max(1, st.num_bounce_steps) for k, v in mismatches.items()},
            "largest_abs_u_float_at_a_difference_nm2": worst, "ulp_of_r2_float_nm2": ulp}


def main(steps=100000):
    res = {"what": "steps on which a float32 restatement of OpenMM's CustomCentroidBondForce evaluation (float centres, sqrtf, "
                   "float distance^2 - radius^2) gives another MMVT bit string than the oracle's double / fixed-point evaluation of "
                   "the same advanced configuration; thermal run at 300 K, dt 2 fs, shell 4 pm wide around the radius",
           "runs": [run("C2-like host-guest groups (147 + 15 atoms)", 192, list(range(0, 147)), list(range(147, 162)), 0.40, steps, 0x5EEC2000 + 71),
                    run("C4-like trypsin-benzamidine groups (11 + 9 atoms)", 32, list(range(0, 11)), list(range(11, 20)), 1.20, steps, 0x5EEC2000 + 72)]}
    return res


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    out = main(n)
    path = os.path.join(ROOT, "profiles", "r02_cv_float_contract.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))
