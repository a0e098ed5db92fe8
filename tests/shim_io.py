"""Scenario files of tests/openmm_shim/src/scenario_driver.cpp and the dumps it writes (shared by the GPU tests that run
the reference's CUDA platform next to the B200 adapter and by the generator of tests/golden/ref_cuda_platform/)."""
import numpy as np

import helpers as H
from helpers import L, s2

PRECISIONS = {"single": L.SINGLE, "mixed": L.MIXED, "double": L.DOUBLE}


def _g(x):
    return "%.17g" % float(x)


def write_scenario(path, syn, kind, precision, steps_list, pool=None, output="", savestate="", savestats="",
                   counter=0, end_on_src=None, milestones=(), srcs=(), dests=(), dump="", settime_before=None, settle=False,
                   extra_lines=()):
    lines = [f"precision {precision}", f"integrator {kind} {_g(syn.temperature)} {_g(syn.friction)} {_g(syn.dt)}"]
    if output:
        lines.append(f"output {output}")
    if savestate:
        lines.append(f"savestate {savestate}")
    if savestats:
        lines.append(f"savestats {savestats}")
    lines.append(f"counter {counter}")
    if end_on_src is not None:
        lines.append(f"endonsrc {1 if end_on_src else 0}")
    lines += [f"milestone {g}" for g in milestones] + [f"src {g}" for g in srcs] + [f"dest {g}" for g in dests]
    for m, x, v in zip(syn.system.masses, syn.positions, syn.velocities):
        lines.append("particle " + " ".join(_g(t) for t in (m, *x, *v)))
    for f in syn.system._forces:
        if isinstance(f, s2.CustomCentroidBondForce):
            lines.append(f"centroid {f.numGroupsPerBond} {f.getForceGroup()} {f.expression}")
            lines += [f"param {n}" for n in f.param_names]
            for atoms, weights in f.groups:
                if weights is None:
                    lines.append(f"group {len(atoms)} " + " ".join(str(a) for a in atoms))
                else:
                    lines.append(f"wgroup {len(atoms)} " + " ".join(str(a) for a in atoms) + " " + " ".join(_g(w) for w in weights))
            for groups, params in f.bonds:
                lines.append("bond " + " ".join(str(g) for g in groups) + " : " + " ".join(_g(p) for p in params))
            lines.append("end")
        elif isinstance(f, s2.TetherForce):
            lines.append(f"tether {_g(f.k)}")
        elif isinstance(f, s2.HarmonicBondForce):
            lines += [f"harmonic {a} {b} {_g(r0)} {_g(k)}" for a, b, r0, k in f.bonds]
    if pool is not None:
        pool_path = path + ".pool"
        np.ascontiguousarray(pool, dtype=np.float32).tofile(pool_path)
        lines.append(f"pool {pool_path} {pool.shape[0]}")
    for i, n in enumerate(steps_list):
        if settime_before is not None and settime_before[0] == i:
            lines.append(f"settime {_g(settime_before[1])}")
        lines.append(f"steps {n}")
    if settle:
        lines.append("settle")        # ContextImpl::calcKineticEnergy(): an asynchronous platform brings files and clock up to date
    lines += list(extra_lines)
    if dump:
        lines.append(f"dump {dump}")
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")



def read_dump(path, precision, n_atoms):
    npad = (n_atoms + 31) // 32 * 32
    raw = open(path, "rb").read()
    real = np.float64 if precision == "double" else np.float32
    mixed = np.float32 if precision == "single" else np.float64
    off = 0
    posq = np.frombuffer(raw, dtype=real, count=4 * npad, offset=off).reshape(npad, 4); off += posq.nbytes
    corr = None
    if precision == "mixed":
        corr = np.frombuffer(raw, dtype=np.float32, count=4 * npad, offset=off).reshape(npad, 4); off += corr.nbytes
    velm = np.frombuffer(raw, dtype=mixed, count=4 * npad, offset=off).reshape(npad, 4); off += velm.nbytes
    time = np.frombuffer(raw, dtype=np.float64, count=1, offset=off)[0]
    step_count = np.frombuffer(raw, dtype=np.int64, count=1, offset=off + 8)[0]
    return H.HostBuffers(PRECISIONS[precision], posq, corr, velm, None), float(time), int(step_count)


