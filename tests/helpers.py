"""Shared test infrastructure: deterministic inputs, the synthetic systems of SURVEY.md 8d, host
buffers in the OpenMM CUDA-platform layouts, and runners for the CPU oracle and the B200 path."""
from __future__ import annotations

import ctypes as C
import os
import sys
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import seekr2_openmm_plugin_b200 as s2  # noqa: E402
from seekr2_openmm_plugin_b200 import _lib as L  # noqa: E402
from seekr2_openmm_plugin_b200.synthetic import (  # noqa: E402,F401
    SEED_BASE, Synthetic, gaussian_pool, lattice_system, make_mmvt_integrator, pair_sphere_system, place_pair_distance,
    toy_system,
)


class _LazyOracle:
    """oracle/ is loaded only when a test (or a CPU-baseline leg) actually asks for it."""

    def __getattr__(self, name):
        from oracle import binding
        return getattr(binding, name)


oracle_binding = _LazyOracle()


@dataclass
class HostBuffers:
    """One anchor's buffers, laid out as OpenMM's CudaContext does."""
    precision: int
    posq: np.ndarray
    corr: Optional[np.ndarray]
    velm: np.ndarray
    force: np.ndarray   # int64 [3, Np]

    def copy(self) -> "HostBuffers":
        return HostBuffers(self.precision, self.posq.copy(), None if self.corr is None else self.corr.copy(),
                           self.velm.copy(), self.force.copy())


def real_dtype(precision):
    return np.float64 if precision == L.DOUBLE else np.float32


def mixed_dtype(precision):
    return np.float32 if precision == L.SINGLE else np.float64


def make_host_buffers(precision: int, positions: np.ndarray, velocities: np.ndarray, masses: np.ndarray,
                      forces: Optional[np.ndarray] = None) -> HostBuffers:
    n = positions.shape[0]
    npad = (n + 31) // 32 * 32
    real, mixed = real_dtype(precision), mixed_dtype(precision)
    posq = np.zeros((npad, 4), dtype=real)
    hi = positions.astype(real)
    posq[:n, :3] = hi
    corr = None
    if precision == L.MIXED:
        corr = np.zeros((npad, 4), dtype=np.float32)
        corr[:n, :3] = (positions - hi.astype(np.float64)).astype(np.float32)
    velm = np.zeros((npad, 4), dtype=mixed)
    velm[:n, :3] = velocities.astype(mixed)
    velm[:n, 3] = np.where(masses > 0, 1.0 / np.where(masses > 0, masses, 1.0), 0.0).astype(mixed)
    force = np.zeros((3, npad), dtype=np.int64)
    if forces is not None:
        force[:, :n] = np.rint(forces.T * 4294967296.0).astype(np.int64)
    return HostBuffers(precision, posq, corr, velm, force)


# ---- running both implementations ---------------------------------------------------------------------

def oracle_for(integ, system, precision: int, anchor: int = 0, num_anchors: int = 1):
    cfg = integ.makeConfig(system, precision, num_anchors)
    o = oracle_binding.Oracle(L, cfg, anchor)
    o.set_params(integ.getTemperature(), integ.getFriction(), integ.getStepSize())
    o._cfg_owner = integ    # keeps the arrays the config points at alive
    return o


def tether_x0(hb: HostBuffers) -> np.ndarray:
    return hb.posq.copy()


def upload(ctx, hb_list: List[HostBuffers]) -> None:
    """Put raw host buffers (one per anchor) into the Context's device tensors bit for bit."""
    import torch
    ctx.stream.synchronize()
    for a, hb in enumerate(hb_list):
        ctx.posq[a].copy_(torch.from_numpy(hb.posq))
        if hb.corr is not None:
            ctx.posq_correction[a].copy_(torch.from_numpy(hb.corr))
        ctx.velm[a].copy_(torch.from_numpy(hb.velm))
        ctx.force[a].copy_(torch.from_numpy(hb.force))
    torch.cuda.synchronize(ctx.device)
    ctx.groups_dirty = True


def download(ctx, anchor: int = 0) -> HostBuffers:
    ctx.stream.synchronize()
    return HostBuffers(ctx.precision, ctx.posq[anchor].cpu().numpy(),
                       None if ctx.posq_correction is None else ctx.posq_correction[anchor].cpu().numpy(),
                       ctx.velm[anchor].cpu().numpy(), ctx.force[anchor].cpu().numpy())


def records_as_tuples(records) -> list:
    return [(r.kind, r.index, r.counter, r.num_surfaces, r.step, float(r.time)) for r in records]


def assert_state_equal(a: HostBuffers, b: HostBuffers, n_atoms: int, what: str = "") -> None:
    """Bit-exact comparison (the arithmetic contract makes this achievable in every precision)."""
    np.testing.assert_array_equal(a.posq[:n_atoms].view(np.uint8), b.posq[:n_atoms].view(np.uint8), err_msg=f"posq {what}")
    if a.corr is not None:
        np.testing.assert_array_equal(a.corr[:n_atoms].view(np.uint8), b.corr[:n_atoms].view(np.uint8), err_msg=f"posqCorrection {what}")
    np.testing.assert_array_equal(a.velm[:n_atoms].view(np.uint8), b.velm[:n_atoms].view(np.uint8), err_msg=f"velm {what}")


def stats_tuple(st, m: int):
    nij = [st.Nij_alpha[i * L.MAX_MILESTONES + j] for i in range(m) for j in range(m)]
    return (st.time, st.incubation_time, st.first_crossing_time, st.T_alpha, list(st.Ri_alpha[:m]), st.step_count,
            list(st.N_alpha_beta[:m]), nij, st.bounce_counter, st.previous_milestone, st.num_bounce_steps,
            st.crossing_counter, st.crossed_src, st.end_simulation)
