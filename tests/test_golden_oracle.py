"""The CPU oracle still reproduces the committed golden vectors (tests/golden/, made by make_golden.py)."""
import filecmp
import importlib.util
import json
import os

import pytest

import helpers as H

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
G = importlib.util.module_from_spec(spec)
spec.loader.exec_module(G)
MANIFEST = json.load(open(os.path.join(GOLDEN, "manifest.json")))
FAST = [n for n in G.cases() if not n.startswith("toy_mmvt_") or n in ("toy_mmvt_mixed",)]


@pytest.mark.parametrize("name", sorted(G.cases()))
def test_oracle_reproduces_golden(name, tmp_path):
    rc, o, hb, syn, labels = G.run_case(name, G.cases()[name])
    want = MANIFEST[name]
    assert rc == want["status"] and len(o.records()) == want["records"]
    assert G.digest(hb, syn.system.getNumParticles()) == want["state_sha256"]
    st = o.state()
    assert (st.time, st.step_count, st.bounce_counter, st.crossing_counter, st.T_alpha) == \
           (want["time"], want["step_count"], want["bounce_counter"], want["crossing_counter"], want["T_alpha"])
    assert list(st.N_alpha_beta[:4]) == want["N_alpha_beta"] and list(st.Ri_alpha[:4]) == want["Ri_alpha"]
    path = str(tmp_path / "log.csv")
    o.write_log(path, labels)
    assert filecmp.cmp(path, os.path.join(GOLDEN, name + ".csv"), shallow=False)
