"""The CPU oracle against outputs of the REFERENCE ITSELF: tests/golden/ref_cuda_platform/ holds what the reference's own
CUDA-platform host code (CudaSeekr2Kernels.cpp) and kernels (langevin.cu, langevinMiddle.cu) -- unmodified, compiled
against tests/openmm_shim, run on a B200 by tests/golden/make_ref_platform_golden.py -- wrote for the cases of
tests/ref_platform_cases.py: crossings files, statistics files, and the digest of the final device buffers.  The oracle
must reproduce them byte for byte / bit for bit from the same seeded inputs.  This is what pins the oracle's crossing
bookkeeping, file formats, Elber logic and mixed-precision arithmetic without a GPU in the loop.  (What it cannot pin is
OpenMM's own evaluation of the boundary expressions: the stand-in evaluates them in double precision.)"""
import importlib.util
import json
import os

import numpy as np
import pytest

import helpers as H
import ref_platform_cases as RC
from helpers import L
from shim_io import PRECISIONS

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_cuda_platform")
spec = importlib.util.spec_from_file_location("make_ref_platform_golden", os.path.join(os.path.dirname(GOLDEN), "make_ref_platform_golden.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))


def test_every_case_has_a_golden_entry():
    assert sorted(MANIFEST) == sorted(RC.cases())
    assert sum(e["crossing_lines"] for e in MANIFEST.values()) > 100


@pytest.mark.parametrize("name", sorted(RC.cases()))
def test_oracle_reproduces_the_reference_cuda_platform(name, tmp_path):
    case = RC.cases()[name]
    want = MANIFEST[name]
    syn = case["make"]()
    n = syn.positions.shape[0]
    precision = PRECISIONS[case["precision"]]
    integ = RC.make_integrator(case, syn)
    steps = sum(case["steps"])
    hb = H.make_host_buffers(precision, syn.positions, syn.velocities, syn.masses)
    pool = RC.pool_for(case, syn)
    if pool is None:
        pool = np.zeros((steps * hb.posq.shape[0], 4), dtype=np.float32)
    o = H.oracle_for(integ, syn.system, precision)
    assert o.run(steps, hb.posq, hb.corr, hb.velm, hb.force, pool, 0, None, 0.0) == 0
    labels = RC.labels_for(case, syn)
    path = str(tmp_path / "crossings.csv")
    o.write_log(path, labels)
    assert open(path, "rb").read() == open(os.path.join(GOLDEN, name + ".csv"), "rb").read()
    assert len(o.records()) == want["crossing_lines"]
    assert G.state_digest(hb, n) == want["state_sha256"]
    st = o.state()
    assert (st.time, st.step_count) == (want["time"], want["step_count"])
    if want.get("has_stats"):
        spath = str(tmp_path / "stats.txt")
        o.write_statistics(spath, labels)
        assert open(spath, "rb").read() == open(os.path.join(GOLDEN, name + ".stats"), "rb").read()
