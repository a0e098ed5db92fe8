"""Save-state-on-crossing on the B200 path: the snapshot the fused / two-pass kernels keep on the device is
bit-equal to the oracle's pre-bounce configuration, files appear under the reference's names
(CudaSeekr2Kernels.cpp:282-293, :575-587), and an overwritten snapshot is reported, never silently wrong."""
import ctypes as C
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2

pytestmark = pytest.mark.gpu


def _read_state(path):
    root = ET.parse(path).getroot()
    pos = np.array([[float(p.attrib[k]) for k in "xyz"] for p in root.find("Positions")])
    vel = np.array([[float(p.attrib[k]) for k in "xyz"] for p in root.find("Velocities")])
    return float(root.attrib["time"]), pos, vel


def _snap_tuples(records):
    return [(r.kind, r.index, r.counter, r.num_surfaces, r.step, float(r.time), r.snapshot) for r in records]


@pytest.mark.parametrize("precision", [L.SINGLE, L.MIXED, L.DOUBLE])
@pytest.mark.parametrize("mode", ["fused", "two_pass"])
@pytest.mark.parametrize("middle", [False, True])
def test_toy_mmvt_crossing_states_match_oracle(cuda_device, tmp_path, precision, mode, middle):
    syn = H.toy_system()
    steps = 3000
    hb = H.make_host_buffers(precision, syn.positions, syn.velocities, syn.masses)
    pool = H.gaussian_pool(H.SEED_BASE + 3, steps * hb.posq.shape[0])
    integ_o = H.make_mmvt_integrator(syn, "", middle=middle) if middle else H.make_mmvt_integrator(syn, "")
    integ_o.setSaveStateFileName(str(tmp_path / "o"))
    rc, o, ob = S.run_oracle(syn, integ_o, precision, steps, pool)
    integ = H.make_mmvt_integrator(syn, "", middle=middle) if middle else H.make_mmvt_integrator(syn, "")
    prefix = str(tmp_path / "state")
    integ.setSaveStateFileName(prefix)
    grc, ctx, gb = S.run_gpu(syn, integ, precision, steps, pool, mode=mode)
    assert rc == grc == 0
    recs = integ.getRecords(0)
    assert len(recs) >= 2 and _snap_tuples(o.records()) == _snap_tuples(recs)
    H.assert_state_equal(ob, gb, 1)
    labels = integ._labels(0)
    for r in recs:
        assert r.snapshot > 0                       # one surface at a time in the toy system
        ox, ov = o.snapshot(r.snapshot)
        t, x, v = _read_state(f"{prefix}_{r.counter}_{labels[r.index]}")
        np.testing.assert_array_equal(x, ox)
        np.testing.assert_array_equal(v, ov)
        assert t == float(r.time)


def test_hostguest_states_multi_anchor_graph(cuda_device, tmp_path):
    """147 + 15-atom centroid groups, 2 anchors in one launch, CUDA graph replay: every atom of the system
    (group atoms and plain ones, massless ones included) is in the saved state."""
    host, guest = list(range(0, 147)), list(range(147, 162))
    syn = H.pair_sphere_system(1000, host, guest, 0.395, 0.405, H.SEED_BASE + 7, tether_k=0.0)
    syn.system.setParticleMass(500, 0.0)
    syn.system.setParticleMass(150, 0.0)       # a massless atom inside the guest group
    syn.masses = syn.system.masses
    H.place_pair_distance(syn, host, guest, 0.40)
    steps, A = 640, 2
    precision = L.MIXED
    hbs, pools = [], []
    for a in range(A):
        hb = H.make_host_buffers(precision, syn.positions, syn.velocities * (1.0 + 0.5 * a), syn.masses)
        hbs.append(hb)
        pools.append(H.gaussian_pool(H.SEED_BASE + 10 + a, steps * hb.posq.shape[0]))
    integ = H.make_mmvt_integrator(syn, "")
    prefix = str(tmp_path / "hg")
    integ.setSaveStateFileName(prefix)
    integ.setUseGraph(True, 8)
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"}, numAnchors=A)
    H.upload(ctx, hbs)
    ctx.setRandomPool(np.stack(pools))
    integ.step(steps)
    total = 0
    for a in range(A):
        integ_o = H.make_mmvt_integrator(syn, "")
        integ_o.setSaveStateFileName(prefix)
        o = H.oracle_for(integ_o, syn.system, precision, a, A)
        ob = hbs[a].copy()
        assert o.run(steps, ob.posq, ob.corr, ob.velm, ob.force, pools[a], 0, None, 0.0) == 0
        recs = integ.getRecords(a)
        assert _snap_tuples(o.records()) == _snap_tuples(recs)
        H.assert_state_equal(ob, H.download(ctx, a), syn.positions.shape[0])
        labels = integ._labels(a)
        for r in recs:
            if r.snapshot == 0:
                continue
            ox, ov = o.snapshot(r.snapshot)
            t, x, v = _read_state(f"{prefix}.{a}_{r.counter}_{labels[r.index]}")
            np.testing.assert_array_equal(x, ox)
            np.testing.assert_array_equal(v, ov)
            total += 1
    assert total >= 2


@pytest.mark.parametrize("mode", ["fused", "two_pass"])
@pytest.mark.parametrize("end_on_src", [False, True])
def test_elber_end_state(cuda_device, tmp_path, mode, end_on_src):
    syn, integ_o = S.elber_ballistic(end_on_src)
    integ_o.setSaveStateFileName(str(tmp_path / "o"))
    rc, o, ob = S.run_oracle(syn, integ_o, L.MIXED, 120)
    syn2, integ = S.elber_ballistic(end_on_src)
    prefix = str(tmp_path / "elber")
    integ.setSaveStateFileName(prefix)
    grc, ctx, gb = S.run_gpu(syn2, integ, L.MIXED, 120, mode=mode)
    assert rc == grc == 0
    recs = integ.getRecords(0)
    assert _snap_tuples(o.records()) == _snap_tuples(recs) and len(recs) == 1 and recs[0].snapshot == 1
    H.assert_state_equal(ob, gb, 1)
    ox, ov = o.snapshot(1)
    t, x, v = _read_state(f"{prefix}_0_{1 if end_on_src else 3}")
    np.testing.assert_array_equal(x, ox)
    np.testing.assert_array_equal(v, ov)


def test_corner_bounce_saves_no_state_and_lost_snapshot_is_reported(cuda_device, tmp_path):
    syn = S.ballistic_toy(extra_plane=True)
    integ = H.make_mmvt_integrator(syn, "")
    prefix = str(tmp_path / "c")
    integ.setSaveStateFileName(prefix)
    grc, ctx, gb = S.run_gpu(syn, integ, L.MIXED, 400)
    recs = integ.getRecords(0)
    corner = [r for r in recs if r.num_surfaces == 2]
    assert corner and all(r.snapshot == 0 for r in corner)
    labels = integ._labels(0)
    for r in recs:
        assert os.path.exists(f"{prefix}_{r.counter}_{labels[r.index]}") == (r.snapshot > 0)
    # straight through the C ABI: 2 slots, many crossings without a drain in between -> the first is gone
    syn = H.toy_system()
    integ = H.make_mmvt_integrator(syn, "")
    integ.saveStateSlots = 2
    integ.setSaveStateFileName(prefix + "lost")
    ctx = s2.Context(syn.system, integ, {"CudaPrecision": "mixed"})
    ctx.setPositions(syn.positions)
    ctx.setVelocities(syn.velocities)
    integ.saveStateFileName = ""        # host side: plain step(n), no chunked draining (the engine still saves)
    integ.step(4000)
    recs = [r for r in integ.getRecords(0) if r.snapshot > 0]
    assert len(recs) > 2
    with pytest.raises(s2.Seekr2Error) as err:
        integ.fetchCrossingState(recs[0])
    assert err.value.code == L.ERR_SNAPSHOT_LOST
    integ.fetchCrossingState(recs[-1])
    integ.fetchCrossingState(recs[-2])
