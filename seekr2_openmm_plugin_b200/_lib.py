"""ctypes binding of ``libseekr2_b200.so`` (the C ABI declared in ``include/seekr2_b200.h``).

There is no fallback of any kind: if the CUDA library is missing or does not export a symbol the
header declares, importing this module raises.  The structures below mirror the header field by
field; ``tests/test_abi.py`` checks sizes and symbols against the header text.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SEEKR2_B200_LIBRARY") or os.path.join(_HERE, "libseekr2_b200.so")   # override: A/B builds of the kernels
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "seekr2_b200.h")

ABI_VERSION = 9
MAX_MILESTONES = 32
MAX_BOUNDARIES = 32
MAX_GROUPS = 16
MAX_ELBER = 16

OK, ERR_INVALID, ERR_CUDA, ERR_FIRST_STEP, ERR_RING_OVERFLOW, ERR_IO, ERR_NO_FORCE_GROUP, ERR_SNAPSHOT_LOST = range(8)
SINGLE, MIXED, DOUBLE = 0, 1, 2
VARIANT_CUDA, VARIANT_REFERENCE = 0, 1
FLAG_RESTORE_CORRECTION = 1
FLAG_DECIDE_LOCAL = 2
FLAG_DECIDE_HANDOFF = 4
FLAG_CV_FLOAT = 8
KIND_LANGEVIN, KIND_MMVT, KIND_ELBER = 0, 1, 2
SCHEME_LANGEVIN, SCHEME_MIDDLE = 0, 1
CV_SPHERE_PAIR, CV_SPHERE_FIXED, CV_PLANE, CV_PAIR_SUM, CV_ANGLE = 0, 1, 2, 3, 4
STYLE_STEP, STYLE_RAW = 0, 1
REC_MMVT, REC_ELBER_SRC, REC_ELBER_DEST, REC_ELBER_DEST_STAR = 0, 1, 2, 3

PRECISION_NAMES = {"single": SINGLE, "mixed": MIXED, "double": DOUBLE}


class Seekr2Error(RuntimeError):
    """What the OpenMM adapter turns into ``OpenMM::OpenMMException`` (seekr2plugin.i:38-48)."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class Boundary(C.Structure):
    _fields_ = [
        ("type", C.c_int32), ("style", C.c_int32), ("groups", C.c_int32 * 6),
        ("axis", C.c_int32), ("force_group", C.c_int32),
        ("bitcode", C.c_double), ("k", C.c_double), ("threshold", C.c_double), ("center", C.c_double * 3),
    ]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("kind", C.c_int32), ("precision", C.c_int32), ("variant", C.c_int32),
        ("flags", C.c_int32), ("scheme", C.c_int32), ("device", C.c_int32), ("num_anchors", C.c_int32), ("num_atoms", C.c_int32),
        ("padded_num_atoms", C.c_int32),
        ("num_groups", C.c_int32), ("group_offsets", C.POINTER(C.c_int32)), ("group_atoms", C.POINTER(C.c_int32)),
        ("group_weights", C.POINTER(C.c_double)),
        ("num_boundaries", C.c_int32), ("boundaries", C.POINTER(Boundary)),
        ("num_milestone_groups", C.c_int32), ("bounce_counter0", C.POINTER(C.c_int32)),
        ("num_src", C.c_int32), ("src_groups", C.POINTER(C.c_int32)),
        ("num_dest", C.c_int32), ("dest_groups", C.POINTER(C.c_int32)),
        ("end_on_src_milestone", C.c_int32), ("crossing_counter0", C.POINTER(C.c_int32)),
        ("ring_capacity", C.c_int32), ("save_state_slots", C.c_int32),
    ]


class Buffers(C.Structure):
    _fields_ = [
        ("d_posq", C.c_void_p), ("d_posq_correction", C.c_void_p), ("d_velm", C.c_void_p),
        ("d_force", C.c_void_p), ("d_pos_delta", C.c_void_p), ("d_random", C.c_void_p),
        ("random_anchor_stride", C.c_int64), ("d_old_velm", C.c_void_p), ("d_old_delta", C.c_void_p),
    ]


class Record(C.Structure):
    _fields_ = [
        ("anchor", C.c_int32), ("kind", C.c_int32), ("index", C.c_int32), ("counter", C.c_int32),
        ("num_surfaces", C.c_int32), ("snapshot", C.c_int32), ("step", C.c_int64), ("time", C.c_double),
    ]


class AnchorState(C.Structure):
    _fields_ = [
        ("time", C.c_double), ("incubation_time", C.c_double), ("first_crossing_time", C.c_double),
        ("T_alpha", C.c_double), ("Ri_alpha", C.c_double * MAX_MILESTONES),
        ("step_count", C.c_int64), ("ring_head", C.c_int64),
        ("N_alpha_beta", C.c_int32 * MAX_MILESTONES), ("Nij_alpha", C.c_int32 * (MAX_MILESTONES * MAX_MILESTONES)),
        ("bounce_counter", C.c_int32), ("previous_milestone", C.c_int32), ("num_bounce_steps", C.c_int32),
        ("last_value_positive", C.c_int32),
        ("crossing_counter", C.c_int32), ("crossed_src", C.c_int32), ("end_simulation", C.c_int32),
        ("elber_sampled", C.c_int32), ("elber_values", C.c_float * MAX_ELBER),
        ("num_snapshots", C.c_int32), ("clock_resets", C.c_int32),
    ]


# every symbol the header declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "seekr2_b200_create": (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    "seekr2_b200_destroy": (C.c_int, [_P]),
    "seekr2_b200_last_error": (C.c_char_p, []),
    "seekr2_b200_abi_version": (C.c_int, []),
    "seekr2_b200_decision_mode": (C.c_int, [C.c_void_p]),
    "seekr2_b200_set_params": (C.c_int, [_P, C.c_double, C.c_double, C.c_double]),
    "seekr2_b200_get_params": (C.c_int, [_P, C.POINTER(C.c_double)]),
    "seekr2_b200_sync_groups": (C.c_int, [_P, C.POINTER(Buffers), _P]),
    "seekr2_b200_resync": (C.c_int, [_P, C.POINTER(Buffers), _P]),
    "seekr2_b200_set_atom_order": (C.c_int, [_P, C.POINTER(C.c_int32), _P]),
    "seekr2_b200_poll": (C.c_int, [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "seekr2_b200_check_first_step": (C.c_int, [_P, C.POINTER(Buffers), C.POINTER(C.c_int32), _P]),
    "seekr2_b200_step": (C.c_int, [_P, C.POINTER(Buffers), C.c_uint32, _P]),
    "seekr2_b200_kick": (C.c_int, [_P, C.POINTER(Buffers), C.c_uint32, _P]),
    "seekr2_b200_finish": (C.c_int, [_P, C.POINTER(Buffers), _P]),
    "seekr2_b200_middle_drift": (C.c_int, [_P, C.POINTER(Buffers), C.c_uint32, _P]),
    "seekr2_b200_graph_begin": (C.c_int, [_P, _P]),
    "seekr2_b200_graph_end": (C.c_int, [_P, _P]),
    "seekr2_b200_graph_launch": (C.c_int, [_P, _P]),
    "seekr2_b200_drain": (C.c_int, [_P, C.POINTER(Record), C.c_int64, C.POINTER(C.c_int64), _P]),
    "seekr2_b200_drain_begin": (C.c_int, [_P, C.c_int64, _P]),
    "seekr2_b200_drain_ready": (C.c_int, [_P]),
    "seekr2_b200_drain_clock": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "seekr2_b200_drain_end": (C.c_int, [_P, C.POINTER(Record), C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "seekr2_b200_get_state": (C.c_int, [_P, C.c_int32, C.POINTER(AnchorState), _P]),
    "seekr2_b200_set_time": (C.c_int, [_P, C.c_int32, C.c_double, C.c_int64, _P]),
    "seekr2_b200_fetch_snapshot": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), _P]),
    "seekr2_b200_write_state_xml": (C.c_int, [C.c_char_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_double, C.POINTER(C.c_double)]),
    "seekr2_b200_write_header": (C.c_int, [C.c_int32, C.c_char_p]),
    "seekr2_b200_append_records": (C.c_int, [C.c_int32, C.c_char_p, C.POINTER(Record), C.c_int64, C.c_int32,
                                             C.POINTER(C.c_int32), C.c_int32, C.c_int32]),
    "seekr2_b200_write_statistics": (C.c_int, [C.c_char_p, C.POINTER(AnchorState), C.POINTER(C.c_int32), C.c_int32]),
    "seekr2_b200_harness_tether_force": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, C.c_double, _P, _P]),
    "seekr2_b200_harness_bond_force": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int32, _P, _P, _P, _P]),
    "seekr2_b200_harness_fill_pool": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, _P]),
    "seekr2_b200_harness_check_sqrt": (C.c_int, [_P, _P]),
    "seekr2_b200_set_rng": (C.c_int, [_P, C.c_uint64, C.c_uint32, C.c_uint32]),
    "seekr2_b200_debug_set_probe": (C.c_int, [_P, _P]),
}


def build(verbose: bool = False) -> str:
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    proc = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc")], capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout[-4000:])
        print(proc.stderr[-4000:])
    if proc.returncode != 0:
        raise RuntimeError("building libseekr2_b200.so failed")
    return LIB_PATH


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU fallback for the SEEKR2 step kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)            # AttributeError if the library lacks a declared symbol
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.seekr2_b200_abi_version() != ABI_VERSION:
        raise ImportError("libseekr2_b200.so ABI version does not match the Python binding")
    return lib


class _StructsOnly:
    """SEEKR2_B200_STRUCTS_ONLY=1 (bench.py --impl reference, the CPU arm): the process wants the struct definitions of the
    ABI and must NOT load the product library, so that the driver's record of loaded .so files shows a clean CPU arm.
    Any attempt to call into the library fails loudly."""

    def __getattr__(self, name):
        raise ImportError(f"libseekr2_b200.so was deliberately not loaded (SEEKR2_B200_STRUCTS_ONLY=1): {name} is unavailable")


lib = _StructsOnly() if os.environ.get("SEEKR2_B200_STRUCTS_ONLY") == "1" else _load()


def check(rc: int) -> None:
    if rc != OK:
        raise Seekr2Error(rc, (lib.seekr2_b200_last_error() or b"").decode())
