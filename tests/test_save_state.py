"""Save-state-on-crossing: the configuration at the moment a single surface is
crossed, before the bounce -- `getState(Positions | Velocities)` at CudaSeekr2Kernels.cpp:282-293 (MMVT)
and :575-587 (Elber) -- its file name, and the State XML document.  CPU part: oracle semantics, derived by
hand from those lines, and the host-side writer."""
import ctypes as C
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2


def test_oracle_mmvt_crossing_state_is_the_pre_bounce_configuration(tmp_path):
    syn = S.ballistic_toy()
    integ = H.make_mmvt_integrator(syn, "")
    integ.setSaveStateFileName(str(tmp_path / "state"))
    rc, o, ob = S.run_oracle(syn, integ, L.DOUBLE, 60)
    recs = o.records()
    assert rc == 0 and len(recs) >= 1 and recs[0].snapshot == 1 and recs[0].num_surfaces == 1
    x, v = o.snapshot(1)
    # ballistic: x(t) = 2.1 + 0.5 t; the outer sphere (r = 1.0 around 1.2) is first passed at step 20
    # (x = 2.1 + 21 * 0.005 after the step that starts at time 0.2): the saved state is that crossed position
    crossed_step = recs[0].step
    np.testing.assert_allclose(x[0], [2.1 + 0.5 * 0.01 * (crossed_step + 1), 1.2, 1.2], rtol=0, atol=1e-12)
    assert x[0, 0] - 1.2 >= 1.0 - 1e-6             # outside the sphere: NOT the restored position
    np.testing.assert_allclose(v[0], [0.5, 0, 0], atol=1e-12)      # NOT the reversed velocity
    assert o.state().num_snapshots == len([r for r in recs if r.snapshot])


def test_oracle_corner_bounce_saves_nothing_and_off_by_default(tmp_path):
    syn = S.ballistic_toy(extra_plane=True)
    integ = H.make_mmvt_integrator(syn, "")
    integ.setSaveStateFileName(str(tmp_path / "state"))
    rc, o, ob = S.run_oracle(syn, integ, L.MIXED, 400)
    recs = o.records()
    corner = [r for r in recs if r.num_surfaces == 2]
    single = [r for r in recs if r.num_surfaces == 1]
    assert corner and single
    assert all(r.snapshot == 0 for r in corner)                    # `num_bounced_surfaces == 1` (:282)
    assert [r.snapshot for r in single] == list(range(1, len(single) + 1))
    integ2 = H.make_mmvt_integrator(syn, "")
    rc, o2, _ = S.run_oracle(syn, integ2, L.MIXED, 400)
    assert all(r.snapshot == 0 for r in o2.records()) and o2.state().num_snapshots == 0


@pytest.mark.parametrize("end_on_src", [False, True])
def test_oracle_elber_state_at_the_ending_crossing(tmp_path, end_on_src):
    syn, integ = S.elber_ballistic(end_on_src)
    integ.setSaveStateFileName(str(tmp_path / "state"))
    rc, o, ob = S.run_oracle(syn, integ, L.MIXED, 120)
    recs = o.records()
    assert len(recs) == 1 and recs[0].snapshot == 1
    x, v = o.snapshot(1)
    radius = 1.0 if end_on_src else 1.1
    # the first configuration on or beyond the surface (the boundary test reads the float posq: 1e-6 slack)
    assert radius - 1e-6 <= x[0, 0] - 1.2 < radius + 0.5 * 0.01 + 1e-6
    assert integ.stateFileName(recs[0]).endswith(f"state_0_{1 if end_on_src else 3}")  # _<crossingCounter>_<endMilestoneGroup>


def test_state_file_names_follow_the_reference(tmp_path):
    integ = s2.MmvtLangevinIntegrator(300.0, 1.0, 0.002, "")
    integ.addMilestoneGroup(7)
    integ.addMilestoneGroup(9)
    integ.setSaveStateFileName("/x/state")
    r = L.Record()
    r.kind, r.index, r.counter, r.anchor = L.REC_MMVT, 1, 41, 0
    assert integ.stateFileName(r) == "/x/state_41_9"               # "_" << bounceCounter << "_" << milestoneGroups[i]


def test_state_xml_document(tmp_path):
    """The writer is host code of the C-ABI library (no GPU needed).  Layout of OpenMM's StateProxy:
    <State time= type="State" version="1"> with PeriodicBoxVectors, Positions, Velocities."""
    path = str(tmp_path / "state_3_2")
    x = np.array([[0.1, -0.2, 0.30000000000000004], [1e-12, 2.5, -3.0]])
    v = np.array([[1.5, 0.0, -0.25], [0.125, 1e5, 3.0]])
    box = np.diag([3.0, 4.0, 5.0]).reshape(9)
    dp = C.POINTER(C.c_double)
    rc = L.lib.seekr2_b200_write_state_xml(path.encode(), 2, x.ctypes.data_as(dp), v.ctypes.data_as(dp), 12.5, box.ctypes.data_as(dp))
    assert rc == 0
    root = ET.parse(path).getroot()
    assert root.tag == "State" and root.attrib["type"] == "State" and root.attrib["version"] == "1"
    assert float(root.attrib["time"]) == 12.5
    pos = np.array([[float(p.attrib[k]) for k in "xyz"] for p in root.find("Positions")])
    vel = np.array([[float(p.attrib[k]) for k in "xyz"] for p in root.find("Velocities")])
    np.testing.assert_array_equal(pos, x)            # 17 significant digits round-trip exactly
    np.testing.assert_array_equal(vel, v)
    bx = root.find("PeriodicBoxVectors")
    assert [float(bx.find(t).attrib[k]) for t, k in zip("ABC", "xyz")] == [3.0, 4.0, 5.0]
    assert L.lib.seekr2_b200_write_state_xml(str(tmp_path / "no" / "dir").encode(), 2, x.ctypes.data_as(dp), None, 0.0, None) == L.ERR_IO
