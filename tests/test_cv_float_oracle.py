"""The oracle's FLAG_CV_FLOAT evaluation == the numpy float32 restatement of OpenMM's evaluation in tools/cv_float_contract.py
("sequential" order), bit for bit, on random configurations next to the surfaces."""
import os
import sys

import numpy as np

import helpers as H
from helpers import L

sys.path.insert(0, os.path.join(H.ROOT, "tools"))
import cv_float_contract as F  # noqa: E402


def test_oracle_float_value_equals_the_numpy_restatement():
    host, guest = list(range(0, 37)), list(range(37, 50))
    rin, rout = 0.40 - 1e-6, 0.40 + 1e-6
    syn = H.pair_sphere_system(64, host, guest, rin, rout, H.SEED_BASE + 91, tether_k=0.0)
    integ = H.make_mmvt_integrator(syn, "")
    integ.setBoundaryValueInFloat(True)
    o = H.oracle_for(integ, syn.system, L.MIXED)
    groups = [np.array(host), np.array(guest)]
    weights32 = [(syn.masses[g] / syn.masses[g].sum()).astype(np.float32) for g in groups]
    rng = np.random.default_rng(5)
    seen = set()
    for trial in range(400):
        H.place_pair_distance(syn, host, guest, 0.40 + rng.normal() * 2e-6)
        hb = H.make_host_buffers(L.MIXED, syn.positions + rng.normal(size=syn.positions.shape) * 1e-7, syn.velocities, syn.masses)
        value = o.group_value(hb.posq, hb.corr, 1)
        (bits32, _), _ = F.float_decisions(hb.posq[:, :3], groups, weights32, [np.float32(rin), np.float32(rout)], [-1.0, 1.0])
        assert int(value) == bits32, (trial, value, bits32)
        seen.add(bits32)
    assert seen == {0, 1, 2}, seen            # inside, beyond the inner and beyond the outer surface all occurred
