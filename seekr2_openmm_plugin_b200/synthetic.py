"""Synthetic systems of SURVEY.md 8d (C1 toy, C2/C4/C5 solvated lattices with centroid-pair spheres and
axis planes) and libm-free deterministic input generators.  Harness-side only: used by the tests, the
benchmark and the smoke test to build inputs; nothing here touches the CPU oracle."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import _lib as L
from .context import BOLTZ
from .integrators import MmvtLangevinIntegrator, MmvtLangevinMiddleIntegrator
from .system import CustomCentroidBondForce, System, TetherForce

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
class Synthetic:
    system: System
    positions: np.ndarray
    velocities: np.ndarray
    masses: np.ndarray
    boundary_force: Optional[CustomCentroidBondForce]
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
    sigma = np.sqrt(BOLTZ * 300.0 / masses)
    velocities = gaussian_pool(seed + 1, n_atoms)[:, :3].astype(np.float64) * sigma[:, None]
    return positions, velocities, masses


def toy_system() -> Synthetic:
    """C1: one particle (m = 39.9, demo1:27) between two fixed-centre spheres."""
    system = System()
    system.addParticle(39.9)
    force = CustomCentroidBondForce(
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
    system = System()
    for m in masses:
        system.addParticle(float(m))
    if tether_k > 0:
        system.addForce(TetherForce(tether_k, positions.copy()))
    force = CustomCentroidBondForce(2, "bitcode*step(k*(distance(g1,g2)^2-radius^2))")
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
                pf = CustomCentroidBondForce(1, f"bitcode*step(k*({name}1-bound))")
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
        if isinstance(f, TetherForce):
            f.reference_positions = syn.positions.copy()


def make_mmvt_integrator(syn: Synthetic, path: str = "", variant: str = "cuda", middle: bool = False) -> MmvtLangevinIntegrator:
    cls = MmvtLangevinMiddleIntegrator if middle else MmvtLangevinIntegrator
    integ = cls(syn.temperature, syn.friction, syn.dt, path)
    for g in syn.milestone_groups:
        integ.addMilestoneGroup(g)
    integ.setVariant(variant)
    return integ


