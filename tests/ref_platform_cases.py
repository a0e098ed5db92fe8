"""Cases whose outputs were produced by the REFERENCE'S OWN CUDA-platform host code and kernels (CudaSeekr2Kernels.cpp +
langevin.cu / langevinMiddle.cu, unmodified, under tests/openmm_shim) on a B200 and committed under
tests/golden/ref_cuda_platform/ by tests/golden/make_ref_platform_golden.py.  All are force-free, so that the CPU
oracle can be run on exactly the same inputs without the stand-in's force evaluation entering."""
import numpy as np

import helpers as H
import scenarios as S
from helpers import L, s2


def _hostguest():
    host, guest = list(range(0, 147)), list(range(147, 162))
    syn = H.pair_sphere_system(600, host, guest, 0.395, 0.405, H.SEED_BASE + 8, tether_k=0.0)
    H.place_pair_distance(syn, host, guest, 0.40)
    return syn


def _weighted():
    """explicit group weights, axis planes, the demo4 pair-sum and the demo5 angle boundary, all in force group 1"""
    host, guest = list(range(0, 11)), list(range(11, 20))
    syn = H.pair_sphere_system(300, host, guest, 0.185, 0.195, H.SEED_BASE + 9, tether_k=0.0)
    H.place_pair_distance(syn, host, guest, 0.19)
    x, m = syn.positions, syn.masses
    w = [1.0 + 0.1 * i for i in range(20)]
    pf = s2.CustomCentroidBondForce(1, "bitcode*step(k*(x1-bound))")
    g = pf.addGroup(list(range(40, 60)), w)
    pf.setForceGroup(1)
    for n in ("bitcode", "k", "bound"):
        pf.addPerBondParameter(n)
    cx = float(np.average(x[40:60, 0], weights=w))
    pf.addBond([g], [4.0, 1.0, cx + 0.004])
    pf.addBond([g], [8.0, -1.0, cx - 0.004])
    syn.system.addForce(pf)
    af = s2.CustomCentroidBondForce(3, "bitcode*step(k*(angle(g1, g2, g3) - ref_angle))")
    ga, gb, gc = af.addGroup([100, 101]), af.addGroup([110, 111, 112]), af.addGroup([120])
    af.setForceGroup(1)
    for n in ("bitcode", "k", "ref_angle"):
        af.addPerBondParameter(n)
    ca = np.average(x[[100, 101]], axis=0, weights=m[[100, 101]])
    cb = np.average(x[[110, 111, 112]], axis=0, weights=m[[110, 111, 112]])
    u, v = ca - cb, x[120] - cb
    ang = float(np.arccos(np.dot(u, v) / np.sqrt(np.dot(u, u) * np.dot(v, v))))
    af.addBond([ga, gb, gc], [16.0, 1.0, ang + 0.01])
    syn.system.addForce(af)
    sf = s2.CustomCentroidBondForce(4, "bitcode*step(k*(distance(g1,g2)^2 + distance(g3,g4)^2 - radius^2))")
    gs = [sf.addGroup([200 + 3 * i, 201 + 3 * i]) for i in range(4)]
    sf.setForceGroup(1)
    for n in ("bitcode", "k", "radius"):
        sf.addPerBondParameter(n)
    cen = [np.average(x[[200 + 3 * i, 201 + 3 * i]], axis=0, weights=m[[200 + 3 * i, 201 + 3 * i]]) for i in range(4)]
    r2 = float(np.sum((cen[0] - cen[1]) ** 2) + np.sum((cen[2] - cen[3]) ** 2))
    sf.addBond(gs, [32.0, 1.0, float(np.sqrt(r2)) + 0.004])
    syn.system.addForce(sf)
    syn.milestone_groups = [1, 2, 3, 4, 5, 6]
    return syn


def cases():
    """name -> dict(make, kind, precision, steps, pool_seed, + integrator options).  `make` returns a Synthetic."""
    out = {}
    for prec in ("single", "mixed", "double"):
        for kind in ("mmvt", "mmvt_middle"):
            out[f"toy_{kind}_{prec}"] = dict(make=H.toy_system, kind=kind, precision=prec, steps=[3000], pool_seed=H.SEED_BASE + 41,
                                             counter=0, stats=True)
    out["corner_mixed"] = dict(make=lambda: S.ballistic_toy(extra_plane=True), kind="mmvt", precision="mixed", steps=[200, 200],
                               pool_seed=None, counter=17, milestones=[1, 2, 3], stats=True)
    out["rawstyle_double"] = dict(make=lambda: S.ballistic_toy(raw_style=True), kind="mmvt", precision="double", steps=[200],
                                  pool_seed=None, counter=0, milestones=[1])
    for kind in ("elber", "elber_middle"):
        for tag, end, speed in (("end0", False, 0.5), ("end1", True, 0.5), ("rev", False, -0.5)):
            out[f"{kind}_{tag}_mixed"] = dict(make=lambda end=end, speed=speed, kind=kind: S.elber_ballistic(end, speed=speed, middle=kind.endswith("middle"))[0],
                                              kind=kind, precision="mixed", steps=[60, 60], pool_seed=None, counter=5,
                                              srcs=[1], dests=[2, 3], end_on_src=end)
    out["hostguest_mmvt_mixed"] = dict(make=_hostguest, kind="mmvt", precision="mixed", steps=[400], pool_seed=H.SEED_BASE + 42, counter=0, stats=True)
    out["hostguest_mmvt_middle_mixed"] = dict(make=_hostguest, kind="mmvt_middle", precision="mixed", steps=[400], pool_seed=H.SEED_BASE + 42, counter=0, stats=True)
    out["weighted_planes_angle_pairsum_mixed"] = dict(make=_weighted, kind="mmvt", precision="mixed", steps=[500], pool_seed=H.SEED_BASE + 43, counter=0, stats=True)
    return out


def pool_for(case, syn):
    if case["pool_seed"] is None:
        return None
    npad = (syn.positions.shape[0] + 31) // 32 * 32
    return H.gaussian_pool(case["pool_seed"], sum(case["steps"]) * npad)


def make_integrator(case, syn):
    kind = case["kind"]
    if kind.startswith("mmvt"):
        integ = H.make_mmvt_integrator(syn, "", middle=kind.endswith("middle"))
        if "milestones" in case:
            integ.milestoneGroups = list(case["milestones"])
        integ.setBounceCounter(case["counter"])
        return integ
    integ = (s2.ElberLangevinMiddleIntegrator if kind.endswith("middle") else s2.ElberLangevinIntegrator)(syn.temperature, syn.friction, syn.dt, "")
    for g in case["srcs"]:
        integ.addSrcMilestoneGroup(g)
    for g in case["dests"]:
        integ.addDestMilestoneGroup(g)
    integ.setEndOnSrcMilestone(case["end_on_src"])
    integ.setCrossingCounter(case["counter"])
    return integ


def labels_for(case, syn):
    if case["kind"].startswith("mmvt"):
        return list(case.get("milestones", syn.milestone_groups))
    return list(case["srcs"]) + list(case["dests"])
