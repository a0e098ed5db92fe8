"""SEEKR2_B200_FLAG_CV_FLOAT: the boundary value in single precision, restating OpenMM's CUDA CustomCentroidBondForce
evaluation (float centres summed in group order, sqrtf, float d^2 - r^2; the reference reads OpenMM's value,
CudaSeekr2Kernels.cpp:248).  GPU (fused hand-off kernel and the two-pass kernels) == oracle bit for bit, on systems
whose groups span one and several 256-slot rounds, with spheres (pair, fixed centre) and planes."""
import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2

pytestmark = pytest.mark.gpu


def _both(syn, steps, pool, precision, mode, graph=0):
    integ_o = H.make_mmvt_integrator(syn, "")
    integ_o.setBoundaryValueInFloat(True)
    rc, o, ob = S.run_oracle(syn, integ_o, precision, steps, pool)
    integ = H.make_mmvt_integrator(syn, "")
    integ.setBoundaryValueInFloat(True)
    grc, ctx, gb = S.run_gpu(syn, integ, precision, steps, pool, mode=mode, graph=graph)
    assert rc == grc == 0
    n = syn.positions.shape[0]
    H.assert_state_equal(ob, gb, n, f"FLAG_CV_FLOAT {mode}")
    assert H.records_as_tuples(o.records()) == H.records_as_tuples(integ.getRecords(0))
    m = len(syn.milestone_groups)
    assert H.stats_tuple(o.state(), m) == H.stats_tuple(integ.getAnchorState(0), m)
    return o, integ


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED])
@pytest.mark.parametrize("mode,graph", [("fused", 0), ("fused", 8), ("two_pass", 0)])
def test_float_boundary_value_equals_the_oracle(cuda_device, precision, mode, graph):
    # 11 + 9 atom groups (one round) with extra planes; and 300 + 40 atom groups (the first group spans two 256-slot rounds)
    for host, guest, n, planes in ((list(range(0, 11)), list(range(11, 20)), 400, True), (list(range(0, 300)), list(range(300, 340)), 700, False)):
        syn = H.pair_sphere_system(n, host, guest, 0.198, 0.202, H.SEED_BASE + 81, tether_k=200.0, extra_planes=planes,
                                   plane_group=list(range(40, 60)) if planes else None)
        H.place_pair_distance(syn, host, guest, 0.20)
        steps = 400
        npad = (n + 31) // 32 * 32
        pool = H.gaussian_pool(H.SEED_BASE + 82, steps * npad)
        o, integ = _both(syn, steps, pool, precision, mode, graph)
        assert integ.getDecisionMode() == "handoff"                 # the float evaluation lives in the decide blocks
        assert o.state().num_bounce_steps >= 3, "the run must bounce"


def test_float_and_double_contracts_differ_only_next_to_a_surface(cuda_device):
    """Same run under both contracts: whenever the two engines part, the step on which they do sits within a few float ulps
    of r^2 of a surface (tools/cv_float_contract.py measures how rarely that happens); here: a short run where they agree."""
    host, guest = list(range(0, 11)), list(range(11, 20))
    syn = H.pair_sphere_system(300, host, guest, 0.198, 0.202, H.SEED_BASE + 83, tether_k=200.0)
    H.place_pair_distance(syn, host, guest, 0.20)
    pool = H.gaussian_pool(H.SEED_BASE + 84, 300 * 320)
    out = []
    for on in (False, True):
        integ = H.make_mmvt_integrator(syn, "")
        integ.setBoundaryValueInFloat(on)
        integ.setDecisionMode("handoff")
        rc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, 300, pool)
        out.append((gb, H.records_as_tuples(integ.getRecords(0))))
    H.assert_state_equal(out[0][0], out[1][0], 300, "double / fixed point vs float boundary value")
    assert out[0][1] == out[1][1] and len(out[0][1]) >= 2


def test_float_flag_is_refused_where_it_has_no_meaning(cuda_device):
    syn = H.toy_system()
    integ = H.make_mmvt_integrator(syn, "")
    integ.setBoundaryValueInFloat(True)
    with pytest.raises(s2.Seekr2Error):
        s2.Context(syn.system, integ, {"CudaPrecision": "double"})
