"""Generates the golden vectors under tests/golden/ from the CPU oracle.

The reference cannot run in this environment (OpenMM absent) and ships no crossing fixtures, so the
golden files pin the ORACLE's output at the time of writing: the CPU suite checks the oracle still
reproduces them (regression pin) and the GPU suite checks the B200 kernels reproduce them byte for
byte.  Inputs are rebuilt from seeds by tests/helpers.py (libm-free generators), so only outputs are
stored.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import helpers as H  # noqa: E402
import scenarios as S  # noqa: E402
from helpers import L, s2  # noqa: E402

NAMES = {L.SINGLE: "single", L.MIXED: "mixed", L.DOUBLE: "double"}


def digest(hb, n):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(hb.posq[:n]).tobytes())
    if hb.corr is not None:
        h.update(np.ascontiguousarray(hb.corr[:n]).tobytes())
    h.update(np.ascontiguousarray(hb.velm[:n]).tobytes())
    return h.hexdigest()


def cases():
    """name -> (syn, integrator factory, precision, steps, pool seed or None, labels)"""
    out = {}
    for prec in (L.SINGLE, L.MIXED, L.DOUBLE):
        out[f"toy_mmvt_{NAMES[prec]}"] = (H.toy_system, lambda syn: H.make_mmvt_integrator(syn, ""), prec, 100000, H.SEED_BASE + 1)
    out["toy_mmvt_mixed_reference_variant"] = (H.toy_system, lambda syn: H.make_mmvt_integrator(syn, "", "reference"), L.MIXED, 100000, H.SEED_BASE + 1)
    out["toy_mmvt_middle_mixed"] = (H.toy_system, lambda syn: H.make_mmvt_integrator(syn, "", middle=True), L.MIXED, 100000, H.SEED_BASE + 1)
    out["ballistic_corner_mixed"] = (lambda: S.ballistic_toy(extra_plane=True), lambda syn: H.make_mmvt_integrator(syn, ""), L.MIXED, 400, None)
    out["ballistic_rawstyle_double"] = (lambda: S.ballistic_toy(raw_style=True), lambda syn: H.make_mmvt_integrator(syn, ""), L.DOUBLE, 200, None)

    def hostguest():
        host, guest = list(range(0, 147)), list(range(147, 162))
        syn = H.pair_sphere_system(2520, host, guest, 0.39, 0.41, H.SEED_BASE + 2, tether_k=1000.0)
        H.place_pair_distance(syn, host, guest, 0.40)
        return syn
    out["hostguest_mmvt_mixed"] = (hostguest, lambda syn: H.make_mmvt_integrator(syn, ""), L.MIXED, 1000, H.SEED_BASE + 22)
    for end in (False, True):
        def elber(end=end):
            return S.elber_ballistic(end)[0]

        def einteg(syn, end=end):
            integ = S.elber_ballistic(end)[1]
            integ.setCrossingCounter(5)
            return integ
        out[f"elber_ballistic_end{int(end)}_mixed"] = (elber, einteg, L.MIXED, 120, None)

        def elber_m(end=end):
            return S.elber_ballistic(end, middle=True)[0]

        def einteg_m(syn, end=end):
            integ = S.elber_ballistic(end, middle=True)[1]
            integ.setCrossingCounter(5)
            return integ
        out[f"elber_middle_ballistic_end{int(end)}_mixed"] = (elber_m, einteg_m, L.MIXED, 120, None)
    return out


def run_case(name, spec):
    make_syn, make_integ, prec, steps, seed = spec
    syn = make_syn()
    integ = make_integ(syn)
    npad = (syn.system.getNumParticles() + 31) // 32 * 32
    pool = H.gaussian_pool(seed, steps * npad) if seed is not None else None
    rc, o, hb = S.run_oracle(syn, integ, prec, steps, pool)
    labels = syn.milestone_groups if integ.KIND == L.KIND_MMVT else integ._labels(0)
    return rc, o, hb, syn, labels


def main():
    manifest = {}
    for name, spec in cases().items():
        rc, o, hb, syn, labels = run_case(name, spec)
        path = os.path.join(HERE, name + ".csv")
        if os.path.exists(path):
            os.remove(path)
        o.write_log(path, labels)
        st = o.state()
        manifest[name] = {"status": rc, "records": len(o.records()), "state_sha256": digest(hb, syn.system.getNumParticles()),
                          "time": st.time, "step_count": st.step_count, "bounce_counter": st.bounce_counter,
                          "crossing_counter": st.crossing_counter, "T_alpha": st.T_alpha,
                          "N_alpha_beta": list(st.N_alpha_beta[:4]), "Ri_alpha": list(st.Ri_alpha[:4])}
        print(name, manifest[name]["records"], "records")
    json.dump(manifest, open(os.path.join(HERE, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
