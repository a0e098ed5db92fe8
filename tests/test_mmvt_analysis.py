"""The MMVT analysis restatement (tests/mmvt_analysis.py) on hand-computable inputs, and the oracle's device-style
statistics against the same statistics re-derived from its own crossings file (what downstream analysis reads)."""
import numpy as np

import helpers as H
import mmvt_analysis as M
from helpers import L


def test_stationary_distribution_and_mfpt_closed_forms():
    x, y = 3.0, 5.0
    pi = M.stationary_distribution(np.array([[0.0, x], [y, 0.0]]))
    np.testing.assert_allclose(pi, [y / (x + y), x / (x + y)])
    a, b, c = 2.0, 0.5, 4.0                                     # 0 <-> 1 -> 2 (absorbing)
    q = np.array([[-a, a, 0.0], [b, -(b + c), c], [0.0, 0.0, 0.0]])
    np.testing.assert_allclose(M.mfpt(q, 0, 2), 1 / a + (a + b) / (a * c))
    np.testing.assert_allclose(M.mfpt(q, 1, 2), (a + b) / (a * c))


def test_crossings_file_carries_the_statistics(tmp_path):
    """N_alpha_beta, N_ij, R_i and T_alpha recomputed from the crossings file equal the ones the step kernel's host
    block accumulates (to the file's %.3f time resolution)."""
    syn = H.toy_system()
    steps = 60000
    integ = H.make_mmvt_integrator(syn, "")
    hb = H.make_host_buffers(L.MIXED, syn.positions, syn.velocities, syn.masses)
    pool = H.gaussian_pool(H.SEED_BASE + 71, steps * 32)
    o = H.oracle_for(integ, syn.system, L.MIXED)
    assert o.run(steps, hb.posq, hb.corr, hb.velm, hb.force, pool, 0, None, 0.0) == 0
    path = str(tmp_path / "crossings.txt")
    o.write_log(path, syn.milestone_groups)
    from_file = M.parse_crossings(path, syn.dt)
    from_state = M.stats_from_state(o.state(), syn.milestone_groups, L.MAX_MILESTONES)
    assert sum(from_state.N_alpha_beta.values()) > 50
    assert from_file.N_alpha_beta == from_state.N_alpha_beta and from_file.N_ij == from_state.N_ij
    assert abs(from_file.T - from_state.T) < 2e-3
    for k, v in from_state.R_i.items():
        assert abs(from_file.R_i[k] - v) < 1e-3 * (1 + from_state.N_ij.get((k, 3 - k), 0))
