"""Host-side mirror of the reference API: expression recogniser, boundary lowering, integrator
setters/getters (MmvtLangevinIntegrator.h / ElberLangevinIntegrator.h), file headers and formats."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers as H
from helpers import L, s2


def test_recognises_reference_example_expressions():
    f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")            # demo3:37
    assert (f._typed.cv_type, f._typed.style) == (L.CV_SPHERE_PAIR, L.STYLE_STEP)
    f = s2.CustomCentroidBondForce(1, "k*((x1-centerx)^2 + (y1-centery)^2 + (z1-centerz)^2 - radius^2)")   # demo1:59
    assert (f._typed.cv_type, f._typed.style) == (L.CV_SPHERE_FIXED, L.STYLE_RAW)
    f = s2.CustomCentroidBondForce(1, "-k*(y1-bound)")                                             # demo2:82
    assert (f._typed.cv_type, f._typed.style, f._typed.axis, f._typed.k_sign) == (L.CV_PLANE, L.STYLE_RAW, 1, -1.0)
    f = s2.CustomCentroidBondForce(1, "bitcode*step(k*(z1-bound))")
    assert (f._typed.cv_type, f._typed.style, f._typed.axis) == (L.CV_PLANE, L.STYLE_STEP, 2)
    f = s2.CustomCentroidBondForce(3, "-k*(angle(g1, g2, g3) - ref_angle)")                        # demo5:24
    assert (f._typed.cv_type, f._typed.style, f._typed.k_sign, f._typed.threshold_name) == (L.CV_ANGLE, L.STYLE_RAW, -1.0, "ref_angle")
    f = s2.CustomCentroidBondForce(6, "-k*(distance(g1,g2)^2 + distance(g3,g4)^2 + distance(g5,g6)^2 - radius^2)")   # demo4:27
    assert (f._typed.cv_type, f._typed.axis, f._typed.k_sign) == (L.CV_PAIR_SUM, 3, -1.0)
    with pytest.raises(s2.Seekr2Error):
        s2.CustomCentroidBondForce(2, "k*(distance(g1,g2)-radius)")                                # not a form the examples use
    with pytest.raises(s2.Seekr2Error):
        s2.CustomCentroidBondForce(4, "k*(dihedral(g1,g2,g3,g4)-phi0)")


def test_lowering_matches_demo3_definition():
    system = s2.System()
    for i in range(40):
        system.addParticle(1.0 + i)
    f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
    g1, g2 = f.addGroup([3, 5, 7]), f.addGroup([10, 11])
    f.setForceGroup(1)
    for n in ("bitcode", "k", "radius"):
        f.addPerBondParameter(n)
    f.addBond([g1, g2], [1.0, -1.0, 1.1])
    f.addBond([g1, g2], [2.0, 1.0, 1.3])
    f.setAnchorBondParameters(1, 1, [2.0, 1.0, 1.5])
    system.addForce(f)
    low = s2.system.lower_boundaries(system, 2)
    assert list(low.group_offsets) == [0, 3, 5] and list(low.group_atoms) == [3, 5, 7, 10, 11]
    np.testing.assert_array_equal(low.group_weights, [4.0, 6.0, 8.0, 11.0, 12.0])      # masses: OpenMM's default weights
    b = low.boundaries[0][0]
    assert (b.type, b.groups[0], b.groups[1], b.force_group, b.bitcode, b.k, b.threshold) == (L.CV_SPHERE_PAIR, 0, 1, 1, 1.0, -1.0, 1.1)
    assert low.boundaries[0][1].threshold == 1.3 and low.boundaries[1][1].threshold == 1.5


def test_mmvt_integrator_api_surface():
    integ = s2.MmvtLangevinIntegrator(300.0, 1.0, 0.002, "out.txt")
    assert (integ.getTemperature(), integ.getFriction(), integ.getStepSize()) == (300.0, 1.0, 0.002)
    assert integ.getRandomNumberSeed() == 0 and integ.getConstraintTolerance() == 1e-5 and integ.getBounceCounter() == 0
    assert integ.getOutputFileName() == "out.txt" and integ.getSaveStateFileName() == "" and integ.getSaveStatisticsFileName() == ""
    assert integ.addMilestoneGroup(1) == 0 and integ.addMilestoneGroup(2) == 1
    assert integ.getNumMilestoneGroups() == 2 and integ.getMilestoneGroup(1) == 2
    with pytest.raises(s2.Seekr2Error):
        integ.getMilestoneGroup(2)
    integ.setBounceCounter(12)
    integ.setTemperature(310.0); integ.setFriction(2.0); integ.setStepSize(0.004); integ.setRandomNumberSeed(5)
    assert (integ.getBounceCounter(), integ.getTemperature(), integ.getFriction(), integ.getStepSize(), integ.getRandomNumberSeed()) == (12, 310.0, 2.0, 0.004, 5)
    assert integ.getKernelNames() == ["IntegrateMmvtLangevinStep"]
    with pytest.raises(s2.Seekr2Error, match="not bound to a context"):
        integ.step(1)
    syn = H.toy_system()
    cfg = integ.makeConfig(syn.system, L.MIXED, num_anchors=3)
    assert cfg.num_milestone_groups == 2 and [cfg.bounce_counter0[i] for i in range(3)] == [12, 12, 12]


def test_elber_integrator_api_surface():
    integ = s2.ElberLangevinIntegrator(300.0, 1.0, 0.002, "out.txt")
    assert integ.addSrcMilestoneGroup(1) == 0 and integ.addDestMilestoneGroup(2) == 0 and integ.addDestMilestoneGroup(3) == 1
    assert (integ.getNumSrcMilestoneGroups(), integ.getNumDestMilestoneGroups()) == (1, 2)
    assert integ.getSrcMilestoneGroup(0) == 1 and integ.getDestMilestoneGroup(1) == 3
    integ.setCrossingCounter(4); integ.setEndOnSrcMilestone(True)
    assert integ.getCrossingCounter() == 4 and integ.getEndOnSrcMilestone() is True
    assert integ.getKernelNames() == ["IntegrateElberLangevinStep"]
    with pytest.raises(s2.Seekr2Error):
        integ.getDestMilestoneGroup(5)


def test_headers_written_only_when_file_absent(tmp_path):
    """CudaSeekr2Kernels.cpp:158-171 and :447-462 (product C++ vs oracle C, byte for byte)."""
    for kind, first in ((L.KIND_MMVT, '#"Bounced boundary ID","bounce index","total time (ps)"\n'),
                        (L.KIND_ELBER, '#"Crossed boundary ID","crossing counter","total time (ps)"\n')):
        p, q = str(tmp_path / f"p{kind}.txt"), str(tmp_path / f"q{kind}.txt")
        L.check(L.lib.seekr2_b200_write_header(kind, p.encode()))
        olib = H.oracle_binding.load(L)
        assert olib.seekr2_oracle_write_header(kind, q.encode()) == 0
        assert open(p).read() == open(q).read()
        assert open(p).readline() == first
        with open(p, "a") as f:
            f.write("1,0,0.002\n")
        L.check(L.lib.seekr2_b200_write_header(kind, p.encode()))      # existing file: untouched
        assert open(p).read().endswith("1,0,0.002\n") and open(p).read().count("#") == (2 if kind == L.KIND_ELBER else 1)


def test_record_and_statistics_formatting(tmp_path):
    """MMVT lines use fixed precision(3) (:259-260,281), Elber lines default ostream formatting (:540,565-567);
    statistics file per :325-343."""
    recs = (L.Record * 4)()
    vals = [(0, L.REC_MMVT, 1, 7, 1, 12, 0.0240000001), (0, L.REC_ELBER_SRC, 0, 3, 1, 5, 1.25e-5),
            (0, L.REC_ELBER_DEST_STAR, 1, 3, 1, 9, 123456.789), (1, L.REC_MMVT, 0, 0, 1, 1, 9.9996)]
    for r, (a, k, i, c, n, s, t) in zip(recs, vals):
        r.anchor, r.kind, r.index, r.counter, r.num_surfaces, r.step, r.time = a, k, i, c, n, s, t
    labels = (C.c_int32 * 3)(11, 22, 33)
    p = str(tmp_path / "log.txt")
    L.check(L.lib.seekr2_b200_append_records(L.KIND_MMVT, p.encode(), recs, 4, 0, labels, 3, 1))
    assert open(p).read() == "22,7,0.024\n11,3,1.25e-05\n33*,3,123457\n"
    L.check(L.lib.seekr2_b200_append_records(L.KIND_MMVT, p.encode(), recs, 4, 1, labels, 3, 1))
    assert open(p).read().endswith("11,0,10.000\n")
    st = L.AnchorState()
    st.N_alpha_beta[0], st.N_alpha_beta[1] = 4, 5
    st.Nij_alpha[0 * L.MAX_MILESTONES + 1], st.Nij_alpha[1 * L.MAX_MILESTONES + 0] = 2, 3
    st.Ri_alpha[0], st.Ri_alpha[1], st.T_alpha = 1.23456, 0.5, 17.00049
    q = str(tmp_path / "stats.txt")
    L.check(L.lib.seekr2_b200_write_statistics(q.encode(), C.byref(st), (C.c_int32 * 2)(1, 2), 2))
    assert open(q).read() == ("N_alpha_1: 4\nN_alpha_2: 5\nN_1_1_alpha: 0\nN_1_2_alpha: 2\nN_2_1_alpha: 3\nN_2_2_alpha: 0\n"
                              "R_1_alpha: 1.235\nR_2_alpha: 0.500\nT_alpha: 17.000\n")
