"""Mini ``Context``: owns device buffers laid out exactly like OpenMM's ``CudaContext``
 so the same C-ABI entry points serve this harness and the OpenMM adapter.

PyTorch is used only as the device allocator / stream provider; every kernel that runs is one of the
library's own (``libseekr2_b200.so``).  ``num_anchors`` replicas of the system are batched in one
set of buffers ([A][Np] slabs) -- the B200-native way to run many MMVT anchors per GPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib as L
from .system import HarmonicBondForce, System, TetherForce, lower_boundaries

BOLTZ = (1.380649e-23 * 6.02214076e23) / 1000.0   # kJ/mol/K, OpenMM SimTKOpenMMRealType.h


class State:
    def __init__(self, time, positions=None, velocities=None, kinetic_energy=None):
        self._time, self._pos, self._vel, self._ke = time, positions, velocities, kinetic_energy

    def getTime(self):
        return self._time

    def getPositions(self, asNumpy: bool = True):
        return self._pos

    def getVelocities(self, asNumpy: bool = True):
        return self._vel

    def getKineticEnergy(self):
        return self._ke


class Context:
    """``Context(system, integrator, properties)`` as in the reference's examples
    (mmvt_demonstration3_tryp_benz_spherical.py:66-70), plus ``numAnchors``."""

    def __init__(self, system: System, integrator, properties: Optional[dict] = None, numAnchors: int = 1,
                 randomPoolSteps: int = 32):
        if not torch.cuda.is_available():
            raise L.Seekr2Error(L.ERR_CUDA, "the SEEKR2 B200 kernels need a CUDA device; there is no CPU fallback")
        properties = properties or {}
        self.system = system
        self.integrator = integrator
        self.precision = L.PRECISION_NAMES[properties.get("CudaPrecision", properties.get("Precision", "mixed"))]
        self.device_index = int(properties.get("CudaDeviceIndex", properties.get("DeviceIndex", torch.cuda.current_device())))
        self.device = torch.device("cuda", self.device_index)
        self.num_anchors = int(numAnchors)
        self.num_atoms = system.getNumParticles()
        self.padded_num_atoms = (self.num_atoms + 31) // 32 * 32
        self.random_pool_steps = int(randomPoolSteps)
        A, Np = self.num_anchors, self.padded_num_atoms
        real = torch.float64 if self.precision == L.DOUBLE else torch.float32
        mixed = torch.float32 if self.precision == L.SINGLE else torch.float64
        self.real_dtype, self.mixed_dtype = real, mixed
        dev = self.device
        self.posq = torch.zeros((A, Np, 4), dtype=real, device=dev)
        self.posq_correction = torch.zeros((A, Np, 4), dtype=real, device=dev) if self.precision == L.MIXED else None
        self.velm = torch.zeros((A, Np, 4), dtype=mixed, device=dev)
        self.force = torch.zeros((A, 3, Np), dtype=torch.int64, device=dev)
        self.pos_delta = None          # allocated on first two-pass step
        self.old_velm = None
        self.old_delta = None
        self.random = None             # allocated by the integrator (pool) or injected
        self.random_injected = False
        self.random_cursor = 0         # step slot inside the pool
        inv_mass = np.zeros(Np, dtype=np.float64)
        m = system.masses
        inv_mass[: self.num_atoms] = np.where(m > 0, 1.0 / np.where(m > 0, m, 1.0), 0.0)
        self.velm[:, :, 3] = torch.as_tensor(inv_mass, dtype=mixed, device=dev)[None, :]
        self.stream = torch.cuda.Stream(device=dev)
        self.lowered = lower_boundaries(system, A)
        # force providers
        self._tether = [f for f in system._forces if isinstance(f, TetherForce)]
        self._bonds = [f for f in system._forces if isinstance(f, HarmonicBondForce)]
        self._x0 = None
        self._bond_atoms = self._bond_params = None
        if self._bonds:
            if A != 1:
                raise L.Seekr2Error(L.ERR_INVALID, "HarmonicBondForce harness supports one anchor")
            atoms = [(b[0], b[1]) for f in self._bonds for b in f.bonds]
            params = [(b[2], b[3]) for f in self._bonds for b in f.bonds]
            self._bond_atoms = torch.tensor(atoms, dtype=torch.int32, device=dev).contiguous()
            self._bond_params = torch.tensor(params, dtype=torch.float64, device=dev).contiguous()
        self.groups_dirty = True       # slab must be (re)gathered before the next fused step
        integrator._initialize(self)

    # -- buffers ---------------------------------------------------------------------------------
    def buffers(self) -> L.Buffers:
        b = L.Buffers()
        b.d_posq = self.posq.data_ptr()
        b.d_posq_correction = self.posq_correction.data_ptr() if self.posq_correction is not None else None
        b.d_velm = self.velm.data_ptr()
        b.d_force = self.force.data_ptr()
        b.d_pos_delta = self.pos_delta.data_ptr() if self.pos_delta is not None else None
        b.d_random = self.random.data_ptr() if self.random is not None else None
        b.random_anchor_stride = int(self.random.shape[1]) if self.random is not None else 0
        b.d_old_velm = self.old_velm.data_ptr() if self.old_velm is not None else None
        b.d_old_delta = self.old_delta.data_ptr() if self.old_delta is not None else None
        return b

    @property
    def stream_ptr(self) -> int:
        return self.stream.cuda_stream

    def ensure_two_pass_buffers(self, need_old_velm: bool, need_old_delta: bool = False) -> None:
        A, Np = self.num_anchors, self.padded_num_atoms
        if need_old_delta and self.old_delta is None:
            self.old_delta = torch.zeros((A, Np, 4), dtype=self.mixed_dtype, device=self.device)
        if self.pos_delta is None:
            self.pos_delta = torch.zeros((A, Np, 4), dtype=self.mixed_dtype, device=self.device)
        if need_old_velm and self.old_velm is None:
            self.old_velm = torch.zeros((A, Np, 4), dtype=self.mixed_dtype, device=self.device)

    # -- state in / out ----------------------------------------------------------------------------
    def _broadcast(self, arr) -> np.ndarray:
        arr = np.asarray(arr, dtype=np.float64)
        if arr.ndim == 2:
            arr = np.broadcast_to(arr[None], (self.num_anchors,) + arr.shape)
        if arr.shape != (self.num_anchors, self.num_atoms, 3):
            raise L.Seekr2Error(L.ERR_INVALID, f"expected an array of shape [A,]N,3; got {arr.shape}")
        return arr

    def _to_device_f64(self, arr) -> torch.Tensor:
        """[A, N, 3] float64 on the device; a single configuration is uploaded once and replicated there."""
        arr = np.asarray(arr, dtype=np.float64)
        if arr.ndim == 3 and arr.shape[0] == 1 and self.num_anchors > 1:
            arr = arr[0]
        if arr.ndim == 2:
            if arr.shape != (self.num_atoms, 3):
                raise L.Seekr2Error(L.ERR_INVALID, f"expected an array of shape [A,]N,3; got {arr.shape}")
            return torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)[None].expand(self.num_anchors, -1, -1)
        if arr.shape != (self.num_anchors, self.num_atoms, 3):
            raise L.Seekr2Error(L.ERR_INVALID, f"expected an array of shape [A,]N,3; got {arr.shape}")
        return torch.from_numpy(np.ascontiguousarray(arr)).to(self.device)

    def setPositions(self, positions) -> None:
        """posq = (real) x; posqCorrection = x - (real) x in mixed precision (OpenMM CudaContext).  One upload of the
        doubles; the split is done on the device (same IEEE roundings as on the host)."""
        self.stream.synchronize()
        x = self._to_device_f64(positions)
        hi = x.to(self.real_dtype)
        self.posq[:, : self.num_atoms, :3] = hi
        if self.posq_correction is not None:
            self.posq_correction[:, : self.num_atoms, :3] = (x - hi.to(torch.float64)).to(torch.float32)
        torch.cuda.synchronize(self.device)
        self.groups_dirty = True

    def setVelocities(self, velocities) -> None:
        self.stream.synchronize()
        self.velm[:, : self.num_atoms, :3] = self._to_device_f64(velocities).to(self.mixed_dtype)
        torch.cuda.synchronize(self.device)
        self.groups_dirty = True

    def setVelocitiesToTemperature(self, temperature: float, randomSeed: int = 0) -> None:
        rng = np.random.default_rng(randomSeed)
        m = self.system.masses
        sigma = np.where(m > 0, np.sqrt(BOLTZ * temperature / np.where(m > 0, m, 1.0)), 0.0)
        v = rng.standard_normal((self.num_anchors, self.num_atoms, 3)) * sigma[None, :, None]
        self.setVelocities(v)

    def setTime(self, time: float) -> None:
        for a in range(self.num_anchors):
            L.check(L.lib.seekr2_b200_set_time(self.integrator._handle, a, float(time), 0 if time == 0 else self.getStepCount(a), self.stream_ptr))

    def getStepCount(self, anchor: int = 0) -> int:
        return int(self.integrator.getAnchorState(anchor).step_count)

    def getTime(self, anchor: int = 0) -> float:
        return float(self.integrator.getAnchorState(anchor).time)

    def positions_host(self) -> np.ndarray:
        self.stream.synchronize()
        x = self.posq[:, : self.num_atoms, :3].to(torch.float64)
        if self.posq_correction is not None:
            x = x + self.posq_correction[:, : self.num_atoms, :3].to(torch.float64)
        return x.cpu().numpy()

    def velocities_host(self) -> np.ndarray:
        self.stream.synchronize()
        return self.velm[:, : self.num_atoms, :3].to(torch.float64).cpu().numpy()

    def getState(self, getPositions: bool = False, getVelocities: bool = False, getEnergy: bool = False) -> State:
        pos = vel = ke = None
        squeeze = (lambda a: a[0]) if self.num_anchors == 1 else (lambda a: a)
        if getPositions:
            pos = squeeze(self.positions_host())
        if getVelocities or getEnergy:
            v = self.velocities_host()
            if getVelocities:
                vel = squeeze(v)
            if getEnergy:
                ke = squeeze(0.5 * np.einsum("anc,n->a", v * v, self.system.masses))
        return State(self.getTime(0), pos, vel, ke)

    # -- the harness "force field" -----------------------------------------------------------------
    def setRandomPool(self, pool: np.ndarray) -> None:
        """Inject Gaussian numbers (north star: "identical injected random numbers").  ``pool`` is
        float32 [A, steps*Np, 4] (or [steps*Np, 4], replicated); step s reads at s*Np + atom."""
        pool = np.asarray(pool, dtype=np.float32)
        if pool.ndim == 2:
            pool = np.broadcast_to(pool[None], (self.num_anchors,) + pool.shape)
        self.stream.synchronize()
        self.random = torch.as_tensor(np.array(pool, dtype=np.float32, order="C", copy=True), device=self.device)
        self.random_injected = True
        self.random_cursor = 0
        self.random_pool_steps = self.random.shape[1] // self.padded_num_atoms

    def prepareForces(self) -> None:
        """One-time set-up of the force providers (never inside a graph capture)."""
        if self._tether and self._x0 is None:
            f = self._tether[0]
            self.stream.synchronize()
            if f.reference_positions is not None:
                x0 = np.array(self._broadcast(f.reference_positions), copy=True)
                self._x0 = torch.zeros_like(self.posq)
                self._x0[:, : self.num_atoms, :3] = torch.as_tensor(x0, dtype=self.real_dtype, device=self.device)
            else:
                self._x0 = self.posq.clone()
            torch.cuda.synchronize(self.device)

    def computeForces(self) -> None:
        """Stand-in for ``context.calcForcesAndEnergy(true,false)`` (MmvtLangevinIntegrator.cpp:99)."""
        s = self.stream_ptr
        for f in self._tether:
            L.check(L.lib.seekr2_b200_harness_tether_force(
                self.precision, self.num_anchors, self.num_atoms, self.padded_num_atoms, self.posq.data_ptr(),
                self.posq_correction.data_ptr() if self.posq_correction is not None else None,
                self._x0.data_ptr(), f.k, self.force.data_ptr(), s))
        if self._bond_atoms is not None:
            L.check(L.lib.seekr2_b200_harness_bond_force(
                self.precision, self.num_atoms, self.padded_num_atoms, self.posq.data_ptr(), None,
                int(self._bond_atoms.shape[0]), self._bond_atoms.data_ptr(), self._bond_params.data_ptr(),
                self.force.data_ptr(), s))

    @property
    def has_force_provider(self) -> bool:
        return bool(self._tether) or self._bond_atoms is not None
