"""North star: "long-run MMVT statistics (N_ij, R_i, and the derived k_on/k_off ...) must agree within statistical
error".  Four concentric-shell anchors of the toy system, many replicas each: the B200 path (in-kernel Philox
numbers, 256 trajectories batched in one launch, CUDA-graph replay, statistics taken from the crossings FILES it
writes) against the CPU oracle (host-generated normals), compared on the MFPT from the innermost to the outermost
milestone obtained with the MMVT equations (tests/mmvt_analysis.py)."""
import math

import numpy as np
import pytest

import helpers as H
import mmvt_analysis as M
from helpers import L, s2

pytestmark = pytest.mark.gpu

CENTER = 1.2
RADII = [0.5, 0.6, 0.7, 0.8, 0.9]        # milestone m at RADII[m]; anchor a between milestones a and a + 1
K = len(RADII) - 1


def _system():
    syn = H.toy_system()
    return syn


def _aggregate(stats_per_replica):
    total = M.AnchorStats()
    for st in stats_per_replica:
        total.add(st)
    return total


def test_shell_mfpt_b200_vs_oracle(cuda_device, tmp_path):
    syn = _system()
    steps, RG, RC = 40000, 64, 12
    # ---- B200: K anchors x RG replicas in one engine, per-anchor radii and labels
    integ = H.make_mmvt_integrator(syn, str(tmp_path / "cross.txt"))
    integ.setRandomNumberSeed(777)
    integ.setUseGraph(True, 200)
    integ.ringCapacity = 4096
    A = K * RG
    positions = np.zeros((A, 1, 3))
    for a in range(A):
        alpha = a // RG
        syn.boundary_force.setAnchorBondParameters(a, 0, [1.0, -1.0, CENTER, CENTER, CENTER, RADII[alpha]])
        syn.boundary_force.setAnchorBondParameters(a, 1, [2.0, 1.0, CENTER, CENTER, CENTER, RADII[alpha + 1]])
        integ.setAnchorMilestoneGroups(a, [alpha, alpha + 1])
        positions[a, 0] = [CENTER + 0.5 * (RADII[alpha] + RADII[alpha + 1]), CENTER, CENTER]
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
    ctx.setPositions(positions)
    ctx.setVelocitiesToTemperature(300.0, 11)
    integ.step(steps)
    names = integ.outputFileNames()
    gpu = [[M.parse_crossings(names[alpha * RG + r], syn.dt) for r in range(RG)] for alpha in range(K)]
    # the files carry what the device accumulated
    for a in (0, A // 2, A - 1):
        dev = M.stats_from_state(integ.getAnchorState(a), integ._labels(a), L.MAX_MILESTONES)
        fil = gpu[a // RG][a % RG]
        assert dev.N_alpha_beta == fil.N_alpha_beta and dev.N_ij == fil.N_ij and abs(dev.T - fil.T) < 2e-3

    # ---- oracle: the same anchors, RC replicas each, host normals
    rng = np.random.default_rng(5)
    cpu = [[] for _ in range(K)]
    for alpha in range(K):
        for r in range(RC):
            syn_o = _system()
            syn_o.boundary_force.bonds[0] = ([0], [1.0, -1.0, CENTER, CENTER, CENTER, RADII[alpha]])
            syn_o.boundary_force.bonds[1] = ([0], [2.0, 1.0, CENTER, CENTER, CENTER, RADII[alpha + 1]])
            integ_o = H.make_mmvt_integrator(syn_o, "")
            o = H.oracle_for(integ_o, syn_o.system, L.MIXED)
            vel = rng.standard_normal((1, 3)) * math.sqrt(s2.BOLTZ * 300.0 / 39.9)
            pos = np.array([[CENTER + 0.5 * (RADII[alpha] + RADII[alpha + 1]), CENTER, CENTER]])
            hb = H.make_host_buffers(L.MIXED, pos, vel, syn_o.masses)
            pool = np.zeros((steps * 32, 4), dtype=np.float32)
            pool[:, :3] = rng.standard_normal((steps * 32, 3)).astype(np.float32)
            assert o.run(steps, hb.posq, hb.corr, hb.velm, hb.force, pool, 0, None, 0.0) == 0
            path = str(tmp_path / f"oracle_{alpha}_{r}.txt")
            o.write_log(path, [alpha, alpha + 1])
            cpu[alpha].append(M.parse_crossings(path, syn.dt))

    g = M.shell_mfpt([_aggregate(x) for x in gpu], 1, K)
    c = M.shell_mfpt([_aggregate(x) for x in cpu], 1, K)
    # bootstrap over replicas for the statistical error of each estimate
    def boot(samples, n=200, seed=1):
        r = np.random.default_rng(seed)
        vals = []
        for _ in range(n):
            pick = [[reps[i] for i in r.integers(0, len(reps), len(reps))] for reps in samples]
            vals.append(M.shell_mfpt([_aggregate(x) for x in pick], 1, K)["mfpt"])
        return float(np.std(vals))
    sg, sc = boot(gpu), boot(cpu)
    print(f"MFPT milestone 1 -> {K}: B200 {g['mfpt']:.3f} +- {sg:.3f} ps ({RG} replicas/anchor), oracle {c['mfpt']:.3f} +- {sc:.3f} ps ({RC} replicas/anchor); "
          f"pi B200 {np.round(g['pi'], 4)}, oracle {np.round(c['pi'], 4)}")
    assert g["mfpt"] > 0 and c["mfpt"] > 0
    assert abs(g["mfpt"] - c["mfpt"]) < 4.0 * math.hypot(sg, sc), (g["mfpt"], sg, c["mfpt"], sc)
    assert abs(g["mfpt"] - c["mfpt"]) < 0.2 * c["mfpt"]
    np.testing.assert_allclose(g["pi"], c["pi"], atol=0.03)
    # free particle in shells: the stationary weights follow the shell volumes
    vol = np.array([RADII[a + 1] ** 3 - RADII[a] ** 3 for a in range(K)])
    np.testing.assert_allclose(g["pi"], vol / vol.sum(), atol=0.03)
