"""Drop-in boundary, end to end: the reference's OWN integrator classes and the reference's OWN
platform tests drive this repository's platforms/b200 adapter through the plugin's kernel interface, and the
reference's OWN CUDA-platform host code + kernels run next to it as the comparison -- both against
tests/openmm_shim, a minimal stand-in for the OpenMM API (NOT OpenMM, which is absent here; see
tests/openmm_shim/include/openmm/ShimCore.h).  Binaries: oracle/build_shim.sh -> oracle/_ref/shim/.

What this pins that nothing else does: the crossing bookkeeping, file formats, statistics, save-state naming,
Elber end / reset logic and the first-step check of CudaSeekr2Kernels.cpp, executed by the reference's real
host code, against the B200 engine -- byte for byte.  What it cannot pin: OpenMM's own evaluation of the
CustomCentroidBondForce expressions (the stand-in evaluates them in double precision)."""
import filecmp
import os
import subprocess
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import helpers as H
import scenarios as S
from helpers import L, s2

pytestmark = pytest.mark.gpu

SHIM = os.path.join(H.ROOT, "oracle", "_ref", "shim")


def _need(binary):
    path = os.path.join(SHIM, binary)
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run oracle/build_shim.sh where /root/reference is present")
    return path


# ---- the reference's own platform tests --------------------------------------------------------------------

REF_TESTS = ["TestCudaMmvtLangevinIntegrator", "TestCudaElberLangevinIntegrator",
             "TestCudaMmvtLangevinMiddleIntegrator", "TestCudaElberLangevinMiddleIntegrator"]


@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
@pytest.mark.parametrize("test", REF_TESTS)
def test_reference_platform_tests_pass_on_the_b200_adapter(cuda_device, test, precision):
    """platforms/cuda/tests/Test*.cpp, unmodified: testSingleBond, testTemperature, testConstraints,
    testConstrainedMasslessParticles, testRandomSeed -- with the B200 kernels behind the kernel factory."""
    proc = subprocess.run([_need("b200_" + test), precision], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and proc.stdout.strip().endswith("Done"), proc.stdout[-2000:] + proc.stderr[-2000:]


@pytest.mark.parametrize("test", REF_TESTS)
def test_reference_platform_tests_pass_on_the_reference_platform_under_the_shim(cuda_device, test):
    """Control: the same tests on the reference's own CUDA platform code.  If these fail the stand-in is wrong."""
    proc = subprocess.run([_need("ref_" + test), "mixed"], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0 and proc.stdout.strip().endswith("Done"), proc.stdout[-2000:] + proc.stderr[-2000:]


from shim_io import PRECISIONS, read_dump, write_scenario  # noqa: E402


# how the adapter steps and drains (platforms/b200/src/B200Seekr2Kernels.cpp): by default one fused launch per step when the
# System has no constraints, and an asynchronous drain; the knobs force the two-pass kernels / the blocking per-step drain
ADAPTER_MODES = {"default": {}, "two_pass": {"SEEKR2_B200_ADAPTER_FUSED": "0"}, "blocking_drain": {"SEEKR2_B200_ASYNC_DRAIN": "0"}}


def run_both(tmp_path, syn, kind, precision, steps_list, pool=None, want_rc=0, env=None, sides=("ref", "b200"), **kw):
    """Same scenario through ref_scenario (reference CUDA platform) and b200_scenario (this repository)."""
    out = {}
    for side in sides:
        d = tmp_path / side
        d.mkdir(parents=True, exist_ok=True)
        names = {k: str(d / k) for k in ("crossings.txt", "stats.txt", "state", "dump.bin")}
        extra = dict(kw)
        use_state = extra.pop("savestate", False)
        use_stats = extra.pop("savestats", False)
        scen = str(d / "scenario.txt")
        write_scenario(scen, syn, kind, precision, steps_list, pool, output=names["crossings.txt"], dump=names["dump.bin"],
                       savestate=names["state"] if use_state else "", savestats=names["stats.txt"] if use_stats else "", **extra)
        proc = subprocess.run([_need(side + "_scenario"), scen], capture_output=True, text=True, timeout=600,
                              env={**os.environ, **(env or {})})
        assert proc.returncode == want_rc, f"{side}: rc {proc.returncode}\n{proc.stdout[-3000:]}\n{proc.stderr[-3000:]}"
        out[side] = (d, proc.stdout)
    return out


def assert_same_outputs(out, precision, n_atoms, states=False, stats=False, min_lines=2):
    (rd, _), (bd, _) = out["ref"], out["b200"]
    ref_lines = open(rd / "crossings.txt").read().splitlines()
    assert len(ref_lines) >= min_lines, ref_lines
    assert filecmp.cmp(rd / "crossings.txt", bd / "crossings.txt", shallow=False), \
        (ref_lines[:12], open(bd / "crossings.txt").read().splitlines()[:12])
    rb, rt, rs = read_dump(rd / "dump.bin", precision, n_atoms)
    bb, bt, bs = read_dump(bd / "dump.bin", precision, n_atoms)
    H.assert_state_equal(rb, bb, n_atoms, "reference CUDA platform vs B200 adapter")
    assert (rt, rs) == (bt, bs)
    if stats:
        assert filecmp.cmp(rd / "stats.txt", bd / "stats.txt", shallow=False)
    if states:
        ref_states = sorted(f for f in os.listdir(rd) if f.startswith("state_"))
        assert ref_states and ref_states == sorted(f for f in os.listdir(bd) if f.startswith("state_"))
        for f in ref_states:
            a, b = ET.parse(rd / f).getroot(), ET.parse(bd / f).getroot()
            assert float(a.attrib["time"]) == float(b.attrib["time"])
            for tag in ("Positions", "Velocities"):
                xa = np.array([[float(p.attrib[k]) for k in "xyz"] for p in a.find(tag)])
                xb = np.array([[float(p.attrib[k]) for k in "xyz"] for p in b.find(tag)])
                np.testing.assert_array_equal(xa, xb, err_msg=f"{f} {tag}")
    return ref_lines, rb


# ---- MMVT ----------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("mode", list(ADAPTER_MODES))
@pytest.mark.parametrize("precision", ["single", "mixed", "double"])
@pytest.mark.parametrize("kind", ["mmvt", "mmvt_middle"])
def test_toy_mmvt_reference_platform_vs_b200_vs_oracle(cuda_device, tmp_path, kind, precision, mode):
    """C1 at 300 K with injected normals: crossings file, statistics file, saved states and final buffers of the
    reference's CUDA platform == B200 adapter == CPU oracle.  (Crossing states are kept here, which makes every mode
    drain inside the step; the asynchronous drain has its own test below.)"""
    syn = H.toy_system()
    steps = 3000
    pool = H.gaussian_pool(H.SEED_BASE + 41, steps * 32)
    out = run_both(tmp_path, syn, kind, precision, [steps], pool, milestones=syn.milestone_groups, savestate=True, savestats=True,
                   env=ADAPTER_MODES[mode])
    lines, rb = assert_same_outputs(out, precision, 1, states=True, stats=True, min_lines=4)
    integ = H.make_mmvt_integrator(syn, "", middle=(kind == "mmvt_middle"))
    rc, o, ob = S.run_oracle(syn, integ, PRECISIONS[precision], steps, pool)
    assert rc == 0
    H.assert_state_equal(ob, rb, 1, "oracle vs reference CUDA platform")
    opath = str(tmp_path / "oracle.txt")
    o.write_log(opath, syn.milestone_groups)
    assert filecmp.cmp(opath, out["ref"][0] / "crossings.txt", shallow=False)
    spath = str(tmp_path / "oracle_stats.txt")
    o.write_statistics(spath, syn.milestone_groups)
    assert filecmp.cmp(spath, out["ref"][0] / "stats.txt", shallow=False)
    # the reference evaluates the boundary group on the host every step; the adapter never does
    assert "energyEvaluations 0 " in out["b200"][1] and f"energyEvaluations {steps + 1} " in out["ref"][1]


@pytest.mark.parametrize("mode", list(ADAPTER_MODES))
@pytest.mark.parametrize("kind", ["mmvt", "mmvt_middle"])
def test_adapter_asynchronous_drain_writes_the_same_files(cuda_device, tmp_path, kind, mode):
    """No crossing states kept: the adapter's execute() never waits for the device (the step kernels report appended
    records through mapped host memory; only then is a drain queued behind a step, and collected by a later step once it
    has landed; flushed at computeKineticEnergy / destruction).  Crossings file, statistics file and final buffers still
    equal the reference CUDA platform's, which reads back and writes every step."""
    syn = H.toy_system()
    pool = H.gaussian_pool(H.SEED_BASE + 43, 3000 * 32)
    out = run_both(tmp_path, syn, kind, "mixed", [1000, 1500, 500], pool, milestones=syn.milestone_groups, savestats=True,
                   env=ADAPTER_MODES[mode])
    lines, _ = assert_same_outputs(out, "mixed", 1, stats=True, min_lines=4)
    assert "energyEvaluations 0 " in out["b200"][1]


@pytest.mark.parametrize("mode", ["default", "two_pass"])
@pytest.mark.parametrize("kind", ["mmvt", "mmvt_middle", "elber"])
def test_adapter_follows_atom_reordering(cuda_device, tmp_path, kind, mode):
    """OpenMM's CUDA platform reorders atoms (cu.reorderAtoms(), CudaSeekr2Kernels.cpp:365); the stand-in does so on request
    (a fixed permutation every 5th call).  The reference is immune -- its kernels treat all atoms alike and the boundary
    Forces are OpenMM's -- while the adapter has to remap its boundary-group atoms (seekr2_b200_set_atom_order)."""
    host, guest = list(range(0, 147)), list(range(147, 162))
    env = {**ADAPTER_MODES[mode], "OPENMM_SHIM_REORDER_EVERY": "5"}
    steps = 300
    pool = H.gaussian_pool(H.SEED_BASE + 44, steps * 608)
    if kind == "elber":
        syn = H.pair_sphere_system(600, host, guest, 0.30, 0.50, H.SEED_BASE + 8, tether_k=300.0)
        syn.system._forces = [f for f in syn.system._forces if not isinstance(f, s2.CustomCentroidBondForce)]
        for group, radius in ((1, 0.40), (2, 0.397), (3, 0.403)):
            f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
            g1, g2 = f.addGroup(host), f.addGroup(guest)
            f.setForceGroup(group)
            for name in ("bitcode", "k", "radius"):
                f.addPerBondParameter(name)
            f.addBond([g1, g2], [1.0, 1.0, radius])
            syn.system.addForce(f)
        H.place_pair_distance(syn, host, guest, 0.3995)
        out = run_both(tmp_path, syn, kind, "mixed", [steps], pool, srcs=[1], dests=[2, 3], end_on_src=False, env=env, settle=True)
        assert_same_outputs(out, "mixed", 600, min_lines=2)
    else:
        syn = H.pair_sphere_system(600, host, guest, 0.395, 0.405, H.SEED_BASE + 8, tether_k=300.0)
        H.place_pair_distance(syn, host, guest, 0.40)
        out = run_both(tmp_path, syn, kind, "mixed", [steps], pool, milestones=syn.milestone_groups, savestats=True, env=env)
        assert_same_outputs(out, "mixed", 600, stats=True, min_lines=3)
    assert f"shimReorders {steps // 5} " in out["b200"][1] and f"shimReorders {steps // 5} " in out["ref"][1]


def test_adapter_refuses_what_it_cannot_type(cuda_device, tmp_path):
    """The reference evaluates whatever sits in a milestone's force group; the adapter evaluates typed terms only and must
    say so instead of silently using defaults: a name that is not a per-bond parameter, a non-centroid Force in the
    MMVT force group."""
    syn = S.ballistic_toy()
    f = next(f for f in syn.system._forces if isinstance(f, s2.CustomCentroidBondForce))
    f.param_names[f.param_names.index("radius")] = "rad"            # the expression still says radius
    out = run_both(tmp_path / "a", syn, "mmvt", "mixed", [3], None, want_rc=1, milestones=[1, 2], sides=("b200",))
    assert "is not a per-bond parameter" in out["b200"][1]
    syn = S.ballistic_toy()
    out = run_both(tmp_path / "b", syn, "mmvt", "mixed", [3], None, want_rc=1, milestones=[1, 2], sides=("b200",),
                   extra_lines=["harmonic 0 0 0.1 1.0 1"])
    assert "not a CustomCentroidBondForce" in out["b200"][1]


def test_corner_bounce_rawstyle_restart_counter(cuda_device, tmp_path):
    for name, syn, steps, ms in (("corner", S.ballistic_toy(extra_plane=True), 400, [1, 2, 3]), ("raw", S.ballistic_toy(raw_style=True), 200, [1])):
        d = tmp_path / name
        d.mkdir()
        out = run_both(d, syn, "mmvt", "mixed", [steps // 2, steps - steps // 2], None, milestones=ms, counter=17, savestate=True)
        raw = name == "raw"
        lines, rb = assert_same_outputs(out, "mixed", 1, states=not raw, min_lines=1 if raw else 3)
        if not raw:
            assert lines[1].split(",")[1] == "17"            # bounceCounter seeds the log (CudaSeekr2Kernels.cpp:151)
            times = [l.split(",")[2] for l in lines[1:]]
            assert len(times) != len(set(times))              # a corner bounce: two lines at the same time
        integ = H.make_mmvt_integrator(syn, "")
        integ.setBounceCounter(17)
        rc, o, ob = S.run_oracle(syn, integ, L.MIXED, steps)
        H.assert_state_equal(ob, rb, 1, f"oracle vs reference CUDA platform ({name})")


def test_first_step_outside_raises_the_reference_message(cuda_device, tmp_path):
    syn = S.ballistic_toy()
    syn.positions[0, 0] = 1.2 + 1.01
    out = run_both(tmp_path, syn, "mmvt", "mixed", [3], None, want_rc=3, milestones=[1, 2])
    for side in ("ref", "b200"):
        assert "MMVT simulation bouncing on first step" in out[side][1]


@pytest.mark.parametrize("kind", ["mmvt", "mmvt_middle"])
def test_host_guest_centroid_pair_spheres(cuda_device, tmp_path, kind):
    """147- and 15-atom mass-weighted centroid groups in a 600-atom system (demo3 form), tether forces."""
    host, guest = list(range(0, 147)), list(range(147, 162))
    syn = H.pair_sphere_system(600, host, guest, 0.395, 0.405, H.SEED_BASE + 8, tether_k=300.0)
    H.place_pair_distance(syn, host, guest, 0.40)
    steps = 400
    pool = H.gaussian_pool(H.SEED_BASE + 42, steps * 608)
    out = run_both(tmp_path, syn, kind, "mixed", [steps], pool, milestones=syn.milestone_groups, savestate=True, savestats=True)
    assert_same_outputs(out, "mixed", 600, states=True, stats=True, min_lines=3)


def test_weighted_groups_planes_pair_sum_angle(cuda_device, tmp_path):
    """Explicit group weights, axis planes, the demo4 pair-sum and the demo5 angle boundary, all in force group 1."""
    host, guest = list(range(0, 11)), list(range(11, 20))
    syn = H.pair_sphere_system(300, host, guest, 0.185, 0.195, H.SEED_BASE + 9, tether_k=0.0)
    H.place_pair_distance(syn, host, guest, 0.19)
    x = syn.positions
    pf = s2.CustomCentroidBondForce(1, "bitcode*step(k*(x1-bound))")
    g = pf.addGroup(list(range(40, 60)), [1.0 + 0.1 * i for i in range(20)])
    pf.setForceGroup(1)
    for n in ("bitcode", "k", "bound"):
        pf.addPerBondParameter(n)
    cx = float(np.average(x[40:60, 0], weights=[1.0 + 0.1 * i for i in range(20)]))
    pf.addBond([g], [4.0, 1.0, cx + 0.004])
    pf.addBond([g], [8.0, -1.0, cx - 0.004])
    syn.system.addForce(pf)
    af = s2.CustomCentroidBondForce(3, "bitcode*step(k*(angle(g1, g2, g3) - ref_angle))")
    ga, gb, gc = af.addGroup([100, 101]), af.addGroup([110, 111, 112]), af.addGroup([120])
    af.setForceGroup(1)
    for n in ("bitcode", "k", "ref_angle"):
        af.addPerBondParameter(n)
    m = syn.masses
    ca = np.average(x[[100, 101]], axis=0, weights=m[[100, 101]])
    cb = np.average(x[[110, 111, 112]], axis=0, weights=m[[110, 111, 112]])
    cc = x[120]
    u, w = ca - cb, cc - cb
    ang = float(np.arccos(np.dot(u, w) / np.sqrt(np.dot(u, u) * np.dot(w, w))))
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
    steps = 500
    pool = H.gaussian_pool(H.SEED_BASE + 43, steps * 320)
    out = run_both(tmp_path, syn, "mmvt", "mixed", [steps], pool, milestones=[1, 2, 3, 4, 5, 6], savestats=True)
    lines, _ = assert_same_outputs(out, "mixed", 300, stats=True, min_lines=4)
    assert len({l.split(",")[0] for l in lines[1:]}) >= 3, lines        # several different surfaces were hit


# ---- Elber ---------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("kind", ["elber", "elber_middle"])
@pytest.mark.parametrize("end_on_src,speed", [(False, 0.5), (True, 0.5), (False, -0.5)])
def test_elber_end_reset_and_asterisk(cuda_device, tmp_path, kind, end_on_src, speed):
    syn, integ = S.elber_ballistic(end_on_src, speed=speed, middle=(kind == "elber_middle"))
    out = run_both(tmp_path, syn, kind, "mixed", [60, 60], None, srcs=[1], dests=[2, 3], end_on_src=end_on_src, counter=5, savestate=True)
    lines, rb = assert_same_outputs(out, "mixed", 1, states=True, min_lines=3)
    assert len(lines) == 3 and lines[2].split(",")[1] == "5"
    integ.setCrossingCounter(5)
    rc, o, ob = S.run_oracle(syn, integ, L.MIXED, 120)
    H.assert_state_equal(ob, rb, 1, "oracle vs reference CUDA platform")
    opath = str(tmp_path / "oracle.txt")
    o.write_log(opath, [1, 2, 3])
    assert filecmp.cmp(opath, out["ref"][0] / "crossings.txt", shallow=False)


@pytest.mark.parametrize("mode", list(ADAPTER_MODES))
def test_elber_thermal_host_guest(cuda_device, tmp_path, mode):
    """No crossing states kept: asynchronous drain; Context time follows the device clock (which Elber resets, :545) through
    the clock every drain carries -- exact again at the latest at calcKineticEnergy() (`settle`)."""
    host, guest = list(range(0, 11)), list(range(11, 20))
    syn = H.pair_sphere_system(300, host, guest, 0.30, 0.50, H.SEED_BASE + 5, tether_k=0.0)
    syn.system._forces = [f for f in syn.system._forces if not isinstance(f, s2.CustomCentroidBondForce)]
    for group, radius in ((1, 0.40), (2, 0.37), (3, 0.43)):
        f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
        g1, g2 = f.addGroup(host), f.addGroup(guest)
        f.setForceGroup(group)
        for name in ("bitcode", "k", "radius"):
            f.addPerBondParameter(name)
        f.addBond([g1, g2], [1.0, 1.0, radius])
        syn.system.addForce(f)
    H.place_pair_distance(syn, host, guest, 0.395)
    steps = 1500
    pool = H.gaussian_pool(H.SEED_BASE + 55, steps * 320)
    out = run_both(tmp_path, syn, "elber", "mixed", [steps // 3, steps - steps // 3], pool, srcs=[1], dests=[2, 3], end_on_src=False,
                   env=ADAPTER_MODES[mode], settle=True)
    assert_same_outputs(out, "mixed", 300, min_lines=3)


@pytest.mark.parametrize("kind", ["elber", "elber_middle"])
def test_elber_reset_reaches_context_time_without_crossing_states(cuda_device, tmp_path, kind):
    """Source crossing resets the clock on the device; no state files, so the drain is asynchronous.  After `settle` the
    Context time equals the reference's; a context.setTime() between two step() calls is followed."""
    syn, integ = S.elber_ballistic(False, speed=0.5, middle=(kind == "elber_middle"))
    out = run_both(tmp_path, syn, kind, "mixed", [40, 80], None, srcs=[1], dests=[2, 3], end_on_src=False, counter=5,
                   settime_before=(1, 0.25), settle=True)
    lines, rb = assert_same_outputs(out, "mixed", 1, min_lines=3)


def test_elber_missing_force_group_message(cuda_device, tmp_path):
    syn, integ = S.elber_ballistic(False)
    out = run_both(tmp_path, syn, "elber", "mixed", [2], None, want_rc=1, srcs=[1], dests=[2, 7], end_on_src=False)
    for side in ("ref", "b200"):
        assert "System contains no force groups used to detect Elber boundary crossings" in out[side][1]


@pytest.mark.parametrize("mode", ["default", "blocking_drain"])
def test_elber_many_clock_resets(cuda_device, tmp_path, mode):
    """A run that sits on its source milestone (tethered atoms, destinations far away): the source is crossed again and again,
    each crossing resets the clock on the device (:545), the adapter learns of it through the reset counter in mapped
    memory and queues drain after drain -- a collected drain is often followed by the next one within the same execute().
    Context time and the final buffers must still equal the reference's, which reads the values back every step."""
    host, guest = list(range(0, 11)), list(range(11, 20))
    syn = H.pair_sphere_system(300, host, guest, 0.30, 0.50, H.SEED_BASE + 5, tether_k=500.0)     # tethered: it keeps coming back
    syn.friction = 5.0
    syn.system._forces = [f for f in syn.system._forces if not isinstance(f, s2.CustomCentroidBondForce)]
    for group, radius in ((1, 0.40), (2, 0.30), (3, 0.50)):
        f = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
        g1, g2 = f.addGroup(host), f.addGroup(guest)
        f.setForceGroup(group)
        for name in ("bitcode", "k", "radius"):
            f.addPerBondParameter(name)
        f.addBond([g1, g2], [1.0, 1.0, radius])
        syn.system.addForce(f)
    H.place_pair_distance(syn, host, guest, 0.40001)
    steps = 2000
    pool = H.gaussian_pool(H.SEED_BASE + 56, steps * 320)
    out = run_both(tmp_path, syn, "elber", "mixed", [steps // 2, steps // 2], pool, srcs=[1], dests=[2, 3], end_on_src=False,
                   env=ADAPTER_MODES[mode], settle=True)
    assert_same_outputs(out, "mixed", 300, min_lines=2)
    (rd, _), (bd, _) = out["ref"], out["b200"]
    _, t_ref, _ = read_dump(rd / "dump.bin", "mixed", 300)
    assert 0.0 < t_ref < 0.5 * steps * syn.dt, t_ref          # the clock was last reset in the second half of the run
