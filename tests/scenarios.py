"""Deterministic scenarios shared by the CPU (oracle) tests, the golden-vector generator and the GPU
parity tests.  Ballistic cases use T = 0, friction = 0 so every crossing step is known in advance."""
from __future__ import annotations

import numpy as np

import helpers as H
from helpers import L, s2


def ballistic_toy(extra_plane: bool = False, raw_style: bool = False, speed: float = 0.5, dt: float = 0.01):
    """One particle moving along +x from r = 0.9 between spheres r = 0.8 (bit 0) and r = 1.0 (bit 1)
    centred at (1.2, 1.2, 1.2); optionally a plane x >= 2.2 (bit 2) that is reached in the same step
    as the outer sphere (corner bounce, CudaSeekr2Kernels.cpp:264-283)."""
    syn = H.toy_system()
    if raw_style:
        # the older one-expression-per-milestone style: bounces, but (int)(1e-9 * ...) == 0 logs nothing
        syn.system._forces = []
        f = s2.CustomCentroidBondForce(1, "k*((x1-centerx)^2 + (y1-centery)^2 + (z1-centerz)^2 - radius^2)")
        g = f.addGroup([0])
        f.setForceGroup(1)
        for name in ("k", "centerx", "centery", "centerz", "radius"):
            f.addPerBondParameter(name)
        f.addBond([g], [1.0e-9, 1.2, 1.2, 1.2, 1.0])
        syn.system.addForce(f)
        syn.boundary_force = f
        syn.milestone_groups = [1]
    if extra_plane:
        pf = s2.CustomCentroidBondForce(1, "bitcode*step(k*(x1-bound))")
        g = pf.addGroup([0])
        pf.setForceGroup(1)
        for name in ("bitcode", "k", "bound"):
            pf.addPerBondParameter(name)
        pf.addBond([g], [4.0, 1.0, 2.2])
        syn.system.addForce(pf)
        syn.milestone_groups = [1, 2, 3]
    syn.positions = np.array([[1.2 + 0.9, 1.2, 1.2]])
    syn.velocities = np.array([[speed, 0.0, 0.0]])
    syn.temperature, syn.friction, syn.dt = 0.0, 0.0, dt
    return syn


def elber_ballistic(end_on_src: bool, speed: float = 0.5, start: float = 0.95, dt: float = 0.01, middle: bool = False):
    """One particle moving along +x; source milestone sphere r = 1.0 (force group 1), destination
    milestones r = 0.9 (group 2) and r = 1.1 (group 3), all centred at (1.2, 1.2, 1.2)."""
    system = s2.System()
    system.addParticle(39.9)
    for group, radius in ((1, 1.0), (2, 0.9), (3, 1.1)):
        f = s2.CustomCentroidBondForce(1, "bitcode*step(k*((x1-centerx)^2 + (y1-centery)^2 + (z1-centerz)^2 - radius^2))")
        g = f.addGroup([0])
        f.setForceGroup(group)
        for name in ("bitcode", "k", "centerx", "centery", "centerz", "radius"):
            f.addPerBondParameter(name)
        f.addBond([g], [1.0, 1.0, 1.2, 1.2, 1.2, radius])
        system.addForce(f)
    syn = H.Synthetic(system, np.array([[1.2 + start, 1.2, 1.2]]), np.array([[speed, 0.0, 0.0]]), np.array([39.9]),
                      None, [], 0.0, dt, temperature=0.0, friction=0.0)
    integ = (s2.ElberLangevinMiddleIntegrator if middle else s2.ElberLangevinIntegrator)(0.0, 0.0, dt, "")
    integ.addSrcMilestoneGroup(1)
    integ.addDestMilestoneGroup(2)
    integ.addDestMilestoneGroup(3)
    integ.setEndOnSrcMilestone(end_on_src)
    return syn, integ


def run_oracle(syn, integ, precision, steps, pool=None):
    hb = H.make_host_buffers(precision, syn.positions, syn.velocities, syn.masses)
    if pool is None:
        pool = np.zeros((steps * hb.posq.shape[0], 4), dtype=np.float32)
    o = H.oracle_for(integ, syn.system, precision)
    x0 = hb.posq.copy() if syn.tether_k > 0 else None
    rc = o.run(steps, hb.posq, hb.corr, hb.velm, hb.force, pool, 0, x0, syn.tether_k)
    return rc, o, hb


def run_gpu(syn, integ, precision, steps, pool=None, mode="fused", graph=0):
    run_gpu.last_error = ""
    names = {L.SINGLE: "single", L.MIXED: "mixed", L.DOUBLE: "double"}
    hb = H.make_host_buffers(precision, syn.positions, syn.velocities, syn.masses)
    if pool is None:
        pool = np.zeros((steps * hb.posq.shape[0], 4), dtype=np.float32)
    integ.setStepMode(mode)
    if graph:
        integ.setUseGraph(True, graph)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": names[precision]})
    H.upload(ctx, [hb])
    ctx.setRandomPool(pool)
    try:
        integ.step(steps)
        rc = 0
    except s2.Seekr2Error as e:
        rc = e.code
        run_gpu.last_error = str(e)                          # shown by the callers' assertion messages
    return rc, ctx, H.download(ctx)
