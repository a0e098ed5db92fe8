"""ctypes binding of the CPU oracle (oracle/libseekr2_oracle.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs,
never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libseekr2_oracle.so")
REF_DIR = os.path.join(_HERE, "_ref")


def build() -> None:
    proc = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("building the oracle failed:\n" + proc.stdout[-2000:] + proc.stderr[-2000:])


def load(structs):
    """structs: the product's _lib module (for the shared struct definitions only)."""
    if not os.path.exists(LIB_PATH):
        build()
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.seekr2_oracle_create.restype = P
    lib.seekr2_oracle_create.argtypes = [C.POINTER(structs.Config), C.c_int]
    lib.seekr2_oracle_destroy.argtypes = [P]
    lib.seekr2_oracle_set_params.argtypes = [P, C.c_double, C.c_double, C.c_double]
    lib.seekr2_oracle_get_params.argtypes = [P, C.POINTER(C.c_double)]
    lib.seekr2_oracle_part1.argtypes = [P, P, P, P, P, C.c_uint]
    lib.seekr2_oracle_part2.argtypes = [P, P, P, P, P, P, P]
    lib.seekr2_oracle_bounce.argtypes = [P, P, P, P, P]
    lib.seekr2_oracle_group_value.restype = C.c_double
    lib.seekr2_oracle_group_value.argtypes = [P, P, P, C.c_int]
    lib.seekr2_oracle_centroids.argtypes = [P, P, P, P]
    for name in ("seekr2_oracle_mmvt_step", "seekr2_oracle_elber_step", "seekr2_oracle_langevin_step"):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [P, P, P, P, P, P, C.c_uint]
    lib.seekr2_oracle_run.restype = C.c_int
    lib.seekr2_oracle_run.argtypes = [P, C.c_int, P, P, P, P, P, C.c_uint, P, C.c_double]
    lib.seekr2_oracle_num_records.restype = C.c_int64
    lib.seekr2_oracle_num_records.argtypes = [P]
    lib.seekr2_oracle_get_records.argtypes = [P, C.POINTER(structs.Record)]
    lib.seekr2_oracle_get_state.argtypes = [P, C.POINTER(structs.AnchorState)]
    lib.seekr2_oracle_get_snapshot.restype = C.c_int
    lib.seekr2_oracle_get_snapshot.argtypes = [P, C.c_int, P, P]
    lib.seekr2_oracle_set_time.argtypes = [P, C.c_double, C.c_int64]
    lib.seekr2_oracle_write_header.restype = C.c_int
    lib.seekr2_oracle_write_header.argtypes = [C.c_int, C.c_char_p]
    lib.seekr2_oracle_append_log.restype = C.c_int
    lib.seekr2_oracle_append_log.argtypes = [P, C.c_char_p, C.POINTER(C.c_int)]
    lib.seekr2_oracle_write_statistics.restype = C.c_int
    lib.seekr2_oracle_write_statistics.argtypes = [P, C.c_char_p, C.POINTER(C.c_int)]
    return lib


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Oracle:
    """One anchor of a configuration, stepped on host numpy buffers in the CUDA-platform layouts."""

    def __init__(self, structs, cfg, anchor: int = 0):
        self.S = structs
        self.lib = load(structs)
        self.h = self.lib.seekr2_oracle_create(C.byref(cfg), anchor)
        if not self.h:
            raise RuntimeError("seekr2_oracle_create failed")
        self.cfg = cfg

    def __del__(self):
        try:
            self.lib.seekr2_oracle_destroy(self.h)
        except Exception:
            pass

    def set_params(self, T, gamma, dt):
        self.lib.seekr2_oracle_set_params(self.h, T, gamma, dt)

    def params(self):
        out = (C.c_double * 3)()
        self.lib.seekr2_oracle_get_params(self.h, out)
        return list(out)

    def step(self, kind, posq, corr, velm, force, random, random_index):
        fn = {self.S.KIND_MMVT: self.lib.seekr2_oracle_mmvt_step, self.S.KIND_ELBER: self.lib.seekr2_oracle_elber_step,
              self.S.KIND_LANGEVIN: self.lib.seekr2_oracle_langevin_step}[kind]
        return fn(self.h, ptr(posq), ptr(corr), ptr(velm), ptr(force), ptr(random), random_index)

    def run(self, steps, posq, corr, velm, force, random, random_index0=0, x0=None, tether_k=0.0):
        return self.lib.seekr2_oracle_run(self.h, steps, ptr(posq), ptr(corr), ptr(velm), ptr(force), ptr(random),
                                          random_index0, ptr(x0), tether_k)

    def group_value(self, posq, corr, force_group):
        return self.lib.seekr2_oracle_group_value(self.h, ptr(posq), ptr(corr), force_group)

    def records(self):
        n = self.lib.seekr2_oracle_num_records(self.h)
        buf = (self.S.Record * max(1, n))()
        self.lib.seekr2_oracle_get_records(self.h, buf)
        return [buf[i] for i in range(n)]

    def state(self):
        st = self.S.AnchorState()
        self.lib.seekr2_oracle_get_state(self.h, C.byref(st))
        return st

    def snapshot(self, seq):
        """(positions[N,3], velocities[N,3]) of crossing state ``seq`` (Record.snapshot)."""
        n = self.cfg.num_atoms
        x, v = np.zeros((n, 3)), np.zeros((n, 3))
        if self.lib.seekr2_oracle_get_snapshot(self.h, seq, ptr(x), ptr(v)) != 0:
            raise IndexError(f"no crossing state {seq}")
        return x, v

    def write_log(self, path, labels):
        arr = (C.c_int * max(1, len(labels)))(*labels)
        self.lib.seekr2_oracle_write_header(self.cfg.kind, path.encode())
        return self.lib.seekr2_oracle_append_log(self.h, path.encode(), arr)

    def write_statistics(self, path, labels):
        arr = (C.c_int * max(1, len(labels)))(*labels)
        return self.lib.seekr2_oracle_write_statistics(self.h, path.encode(), arr)
