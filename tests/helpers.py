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
from oracle import binding as oracle_binding  # noqa: E402

SEED_BASE = 0x5EEC2000


def gaussian_pool(seed: int, count: int) -> np.ndarray:
    """float32 [count, 4] unit-variance numbers (.w = 0) that are bit-reproducible everywhere:
    built from the raw PCG64 stream (stable by NumPy policy) with integer arithmetic only
    (Irwin-Hall sum of twelve 16-bit uniforms), so no libm call is involved."""
    raw = np.random.PCG64(seed).random_raw(count * 3 * 3).reshape(count, 3, 3)
    total = np.zeros((count, 3), dtype=np.int64)
    for word in range(3):
        w = raw[:, :, word]
        for shift in (0, 16, 32, 48):
            total += ((w >> np.uint64(shift)) & np.uint64(0xFFFF)).astype(np.int64)
    normal = (total.astype(np.float64) - 12 * 32767.5) / 65536.0
    out = np.zeros((count, 4), dtype=np.float32)
    out[:, :3] = normal.astype(np.float32)
    return out


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


# ---- synthetic systems ---------------------------------------------------------------

@dataclass
class Synthetic:
    system: "s2.System"
    positions: np.ndarray
    velocities: np.ndarray
    masses: np.ndarray
    boundary_force: Optional["s2.CustomCentroidBondForce"]
    milestone_groups: List[int]
    tether_k: float
    dt: float
    temperature: float = 300.0
    friction: float = 1.0


def lattice_system(n_atoms: int, seed: int, n_solute: int = 0) -> tuple:
    """Jittered simple-cubic lattice at 100 atoms/nm^3, water-like masses, Maxwell-Boltzmann 300 K."""
    raw = np.random.PCG64(seed).random_raw(n_atoms * 3).reshape(n_atoms, 3)
    uniform = (raw >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)   # exact, libm-free
    side = int(np.ceil(n_atoms ** (1.0 / 3.0)))
    spacing = (1.0 / 100.0) ** (1.0 / 3.0)
    idx = np.arange(side ** 3)[:n_atoms]
    grid = np.stack([idx // (side * side), (idx // side) % side, idx % side], axis=1).astype(np.float64)
    jitter = (uniform - 0.5) * 0.2 * spacing
    positions = (grid + 0.5) * spacing + jitter
    masses = np.tile(np.array([15.999, 1.008, 1.008]), n_atoms // 3 + 1)[:n_atoms].copy()
    masses[:n_solute] = 12.011
    sigma = np.sqrt(s2.BOLTZ * 300.0 / masses)
    velocities = gaussian_pool(seed + 1, n_atoms)[:, :3].astype(np.float64) * sigma[:, None]
    return positions, velocities, masses


def toy_system() -> Synthetic:
    """C1: one particle (m = 39.9, demo1:27) between two fixed-centre spheres."""
    system = s2.System()
    system.addParticle(39.9)
    force = s2.CustomCentroidBondForce(
        1, "bitcode*step(k*((x1-centerx)^2 + (y1-centery)^2 + (z1-centerz)^2 - radius^2))")
    g = force.addGroup([0])
    force.setForceGroup(1)
    for name in ("bitcode", "k", "centerx", "centery", "centerz", "radius"):
        force.addPerBondParameter(name)
    force.addBond([g], [1.0, -1.0, 1.2, 1.2, 1.2, 0.8])
    force.addBond([g], [2.0, 1.0, 1.2, 1.2, 1.2, 1.0])
    system.addForce(force)
    positions = np.array([[1.2 + 0.9, 1.2, 1.2]])
    velocities = np.array([[0.05, -0.11, 0.2]])
    return Synthetic(system, positions, velocities, np.array([39.9]), force, [1, 2], 0.0, 0.002)


def pair_sphere_system(n_atoms: int, host: List[int], guest: List[int], r_in: float, r_out: float, seed: int,
                       dt: float = 0.002, tether_k: float = 1000.0, extra_planes: bool = False,
                       plane_group: Optional[List[int]] = None) -> Synthetic:
    """C2 / C4 / C5: solvated lattice with two centroid groups and centroid-pair sphere milestones
    (demo3:37-45); optionally four axis planes on a third group (C5)."""
    positions, velocities, masses = lattice_system(n_atoms, seed, n_solute=len(host) + len(guest))
    system = s2.System()
    for m in masses:
        system.addParticle(float(m))
    if tether_k > 0:
        system.addForce(s2.TetherForce(tether_k, positions.copy()))
    force = s2.CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
    g1 = force.addGroup(host)
    g2 = force.addGroup(guest)
    force.setForceGroup(1)
    for name in ("bitcode", "k", "radius"):
        force.addPerBondParameter(name)
    force.addBond([g1, g2], [1.0, -1.0, r_in])
    force.addBond([g1, g2], [2.0, 1.0, r_out])
    system.addForce(force)
    groups = [1, 2]
    if extra_planes:
        pg = plane_group if plane_group is not None else list(range(64))
        c = positions[pg].mean(axis=0)
        bit = 4.0
        for axis, name in enumerate("xy"):
            for sign, off in ((1.0, 0.02), (-1.0, -0.02)):
                pf = s2.CustomCentroidBondForce(1, f"bitcode*step(k*({name}1-bound))")
                gi = pf.addGroup(pg)
                pf.setForceGroup(1)
                for pname in ("bitcode", "k", "bound"):
                    pf.addPerBondParameter(pname)
                pf.addBond([gi], [bit, sign, float(c[axis] + off)])
                system.addForce(pf)
                groups.append(len(groups) + 1)
                bit *= 2
    return Synthetic(system, positions, velocities, masses, force, groups, tether_k, dt)


def place_pair_distance(syn: Synthetic, host: List[int], guest: List[int], distance: float) -> None:
    """Rigidly shift the guest atoms so the mass-weighted centroid distance equals ``distance``."""
    m = syn.masses
    ch = (syn.positions[host] * m[host, None]).sum(0) / m[host].sum()
    cg = (syn.positions[guest] * m[guest, None]).sum(0) / m[guest].sum()
    d = cg - ch
    target = ch + d / np.linalg.norm(d) * distance
    syn.positions[guest] += target - cg
    for f in syn.system._forces:
        if isinstance(f, s2.TetherForce):
            f.reference_positions = syn.positions.copy()


# ---- running both implementations ---------------------------------------------------------------------

def make_mmvt_integrator(syn: Synthetic, path: str = "", variant: str = "cuda", middle: bool = False) -> "s2.MmvtLangevinIntegrator":
    cls = s2.MmvtLangevinMiddleIntegrator if middle else s2.MmvtLangevinIntegrator
    integ = cls(syn.temperature, syn.friction, syn.dt, path)
    for g in syn.milestone_groups:
        integ.addMilestoneGroup(g)
    integ.setVariant(variant)
    return integ


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
