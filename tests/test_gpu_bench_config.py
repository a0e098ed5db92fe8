"""Parity ON the benchmarked configurations (bench.py's headline and its C5 leg), not on scaled-down relatives of them:
A = 128 anchors x 23 036 atoms (C4) and A = 32 x 100 000 atoms with spheres + planes (C5), decide-block hand-off kernel,
in-kernel Philox normals, 32-step CUDA graphs -- so that the blocks beyond the first wave, the head start, the L2 prefetch of
the successor tile and programmatic dependent launch are all live, as they are in the timed region.  The oracle gets the
very same normals through a pool filled by the harness kernel that shares the generator with the step kernels
(seekr2_b200_harness_fill_pool), one anchor at a time, and must reproduce state, crossing records and statistics of the
first, a middle and the last anchor bit for bit."""
import ctypes as C

import numpy as np
import pytest

import helpers as H
from helpers import L, s2

pytestmark = pytest.mark.gpu

REC = [2478, 2489, 2499, 2535, 2718, 2745, 2769, 2787, 2794, 2867, 2926]
LIG = [3221, 3222, 3223, 3224, 3225, 3226, 3227, 3228, 3229]
STEPS, GRAPH = 64, 32


@pytest.mark.parametrize("n_atoms,anchors,planes", [(23036, 128, False), (100000, 32, True)])
def test_benchmarked_configuration_equals_the_oracle(cuda_device, n_atoms, anchors, planes):
    import torch
    seed = 0x5EEC2000 + 4
    A = anchors
    checked = (0, A // 2 - 1, A - 1)
    dt = 0.004 if not planes else 0.002
    syn = H.pair_sphere_system(n_atoms, REC, LIG, 1.1, 1.3, seed, dt=dt, tether_k=0.0, extra_planes=planes,
                               plane_group=list(range(5000, 5064)))
    # bench.py's shells (0.2 nm wide, from 0.9 nm); the checked anchors sit in shells 0.4 pm wide so that they bounce
    # within the 64 steps -- the bounce path of the streaming blocks is what the timed region rarely sees
    shells = [(0.9 + 0.2 * (a % 8), 1.1 + 0.2 * (a % 8)) for a in range(A)]
    for a in checked:
        shells[a] = (1.10 + 0.01 * a / A, 1.1004 + 0.01 * a / A)
    for a, (rin, rout) in enumerate(shells):
        syn.boundary_force.setAnchorBondParameters(a, 0, [1.0, -1.0, rin])
        syn.boundary_force.setAnchorBondParameters(a, 1, [2.0, 1.0, rout])
    base = syn.positions.copy()
    positions = np.empty((A, n_atoms, 3))
    for a, (rin, rout) in enumerate(shells):
        syn.positions = base.copy()
        H.place_pair_distance(syn, REC, LIG, 0.5 * (rin + rout))
        positions[a] = syn.positions
    syn.positions = base
    forces = H.gaussian_pool(seed + 99, n_atoms)[:, :3].astype(np.float64)        # static force buffer, as in bench.py

    integ = H.make_mmvt_integrator(syn, "")
    integ.setRandomNumberSeed(seed)
    integ.rngAnchorBase, integ.rngAnchorStride = 5, 8                                # rank 5 of 8, anchors dealt round-robin (bench.py --gpus 8)
    integ.setDecisionMode("handoff")
    integ.setUseGraph(True, GRAPH)
    integ.ringCapacity = 1 << 10
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A, randomPoolSteps=GRAPH)
    ctx.setPositions(positions)
    ctx.setVelocities(syn.velocities)
    npad = ctx.padded_num_atoms
    f = np.zeros((3, npad), dtype=np.int64)
    f[:, :n_atoms] = np.rint(forces.T * 4294967296.0).astype(np.int64)
    ctx.force.copy_(torch.from_numpy(f).to(ctx.device)[None].expand(A, -1, -1))
    torch.cuda.synchronize()
    start = {a: H.download(ctx, a) for a in checked}
    assert ctx.random is None                                                        # in-kernel generator, no pool
    integ.step(STEPS)
    assert integ.getDecisionMode() == "handoff" and integ._graph_key is not None     # graph replay really ran

    pool_dev = torch.empty((1, STEPS * npad, 4), dtype=torch.float32, device=ctx.device)
    total_bounces = 0
    for a in checked:
        # the normals the kernel generated for this anchor: (seed; atom, anchor_base + a, step)
        L.check(L.lib.seekr2_b200_harness_fill_pool(pool_dev.data_ptr(), 1, npad, STEPS * npad, STEPS, seed, 5 + 8 * a, 1, 0, ctx.stream_ptr))
        ctx.stream.synchronize()
        pool = pool_dev[0].cpu().numpy()
        o = H.oracle_for(H.make_mmvt_integrator(syn, ""), syn.system, L.MIXED, anchor=a, num_anchors=A)
        ob = start[a].copy()
        assert o.run(STEPS, ob.posq, ob.corr, ob.velm, ob.force, pool, 0, None, 0.0) == 0
        H.assert_state_equal(ob, H.download(ctx, a), n_atoms, f"anchor {a}")
        m = len(syn.milestone_groups)
        assert H.stats_tuple(o.state(), m) == H.stats_tuple(integ.getAnchorState(a), m), f"statistics of anchor {a}"
        assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(a)), f"records of anchor {a}"
        total_bounces += o.state().num_bounce_steps
    assert total_bounces >= 3, "the checked anchors must bounce"
