"""Host-side mirror of the reference's integrator classes for the B200 step.

Same names, argument meaning and error behaviour as
``seekr2plugin/openmmapi/include/MmvtLangevinIntegrator.h:60-226`` and
``ElberLangevinIntegrator.h:60-232`` (and the SWIG classes of ``python/seekr2plugin.i:443-657``),
so the parity tests read like the reference's own tests.  ``step(n)`` follows
``MmvtLangevinIntegrator::step`` (openmmapi/src/MmvtLangevinIntegrator.cpp:94-102):
per step ``updateContextState`` (nothing here), the force pass, ``kernel.execute`` -- which is one
``seekr2_b200_step`` launch instead of the reference's three kernels, generic energy pass and
blocking readback.  Crossing records are drained and appended to the output file when ``step``
returns (the file then holds exactly what the reference would have written).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib as L


def _i32_array(values: Sequence[int]):
    arr = (C.c_int32 * max(1, len(values)))(*values)
    return arr


class _B200LangevinBase:
    KIND = L.KIND_LANGEVIN
    SCHEME = L.SCHEME_LANGEVIN
    KERNEL_NAME = "IntegrateLangevinStep"

    def __init__(self, temperature: float, frictionCoeff: float, stepSize: float, fileName: str = ""):
        self.temperature = float(temperature)
        self.friction = float(frictionCoeff)
        self.stepSize = float(stepSize)
        self.randomNumberSeed = 0
        self.outputFileName = fileName
        self.saveStateFileName = ""
        self.constraintTolerance = 1e-5
        self.context = None
        self._handle = C.c_void_p()
        # B200 knobs (platform properties in the OpenMM adapter)
        self.variant = L.VARIANT_CUDA
        self.flags = 0
        self.stepMode = "fused"        # "fused" (one launch) or "two_pass" (kick / constraints / finish)
        self.useGraph = False
        self.stepsPerGraph = 32
        self.ringCapacity = 0
        self.saveStateSlots = 32       # crossing states kept on the device between drains (saveStateFileName set)
        self.drainInterval = 4096      # steps between asynchronous drains inside one step(n) call
        self.useRandomPool = False     # True: Gaussian pool refilled by a harness kernel (OpenMM style)
        self.rngAnchorBase = 0         # Philox stream of local anchor a = rngAnchorBase + a * rngAnchorStride = its GLOBAL id
        self.rngAnchorStride = 1       # (multi-GPU, anchors dealt round-robin: base = rank, stride = world size)
        self._steps_issued = 0
        self._prev = (None, None, None)
        self._graph_key = None
        self._records: List[L.Record] = []
        self._keep_refs = []

    # -- OpenMM::Integrator surface ------------------------------------------------------------
    def getTemperature(self): return self.temperature
    def setTemperature(self, temp): self.temperature = float(temp)
    def getFriction(self): return self.friction
    def setFriction(self, coeff): self.friction = float(coeff)
    def getStepSize(self): return self.stepSize
    def setStepSize(self, size): self.stepSize = float(size)
    def getRandomNumberSeed(self): return self.randomNumberSeed
    def setRandomNumberSeed(self, seed): self.randomNumberSeed = int(seed)
    def getConstraintTolerance(self): return self.constraintTolerance
    def setConstraintTolerance(self, tol): self.constraintTolerance = float(tol)
    def getOutputFileName(self): return self.outputFileName
    def setOutputFileName(self, fileName): self.outputFileName = fileName
    def getSaveStateFileName(self): return self.saveStateFileName
    def setSaveStateFileName(self, fileName): self.saveStateFileName = fileName
    def getKernelNames(self): return [self.KERNEL_NAME]

    # -- B200 extensions ---------------------------------------------------------------------------
    def setVariant(self, variant: str) -> None:
        self.variant = {"cuda": L.VARIANT_CUDA, "reference": L.VARIANT_REFERENCE}[variant]

    def setRestoreCorrection(self, on: bool) -> None:
        self.flags = (self.flags | L.FLAG_RESTORE_CORRECTION) if on else (self.flags & ~L.FLAG_RESTORE_CORRECTION)

    def setBoundaryValueInFloat(self, on: bool) -> None:
        """Evaluate the boundary value in single precision, restating OpenMM's CUDA ``CustomCentroidBondForce`` evaluation
        (``SEEKR2_B200_FLAG_CV_FLOAT``), instead of the default double / fixed-point contract.  Takes effect when the
        Context is created."""
        self.flags = (self.flags | L.FLAG_CV_FLOAT) if on else (self.flags & ~L.FLAG_CV_FLOAT)

    def setDecisionMode(self, mode: str) -> None:
        """How the fused step hands the crossing decision to the storing threads: "auto" (by launch size), "local"
        (every block evaluates the boundaries; latency-bound launches) or "handoff" (one decide block per anchor).
        Bit-identical results; takes effect when the Context is created."""
        bits = {"auto": 0, "local": L.FLAG_DECIDE_LOCAL, "handoff": L.FLAG_DECIDE_HANDOFF}[mode]
        self.flags = (self.flags & ~(L.FLAG_DECIDE_LOCAL | L.FLAG_DECIDE_HANDOFF)) | bits

    def getDecisionMode(self) -> str:
        """The mode the engine actually runs in (after the Context exists)."""
        return "local" if L.lib.seekr2_b200_decision_mode(self._handle) else "handoff"

    def setStepMode(self, mode: str) -> None:
        assert mode in ("fused", "two_pass")
        self.stepMode = mode

    def setUseGraph(self, on: bool, stepsPerGraph: int = 32) -> None:
        self.useGraph = bool(on)
        self.stepsPerGraph = int(stepsPerGraph)

    # -- lifetime ----------------------------------------------------------------------------------
    def _fill_config(self, cfg: L.Config, num_anchors: int) -> None:
        pass

    def makeConfig(self, system, precision: int, num_anchors: int = 1, device: int = -1, lowered=None) -> L.Config:
        """Lower integrator + System into the ``seekr2_b200_config`` the C ABI takes.  Needs no GPU
        (the CPU oracle is built from the very same struct in the tests)."""
        from .system import lower_boundaries
        low = lowered if lowered is not None else lower_boundaries(system, num_anchors)
        n = system.getNumParticles()
        cfg = L.Config()
        cfg.abi_version = L.ABI_VERSION
        cfg.kind = self.KIND
        cfg.precision = precision
        cfg.variant = self.variant
        cfg.flags = self.flags
        cfg.scheme = self.SCHEME
        cfg.device = device
        cfg.num_anchors = num_anchors
        cfg.num_atoms = n
        cfg.padded_num_atoms = (n + 31) // 32 * 32
        cfg.num_groups = low.num_groups
        offs = np.ascontiguousarray(low.group_offsets, dtype=np.int32)
        atoms = np.ascontiguousarray(low.group_atoms, dtype=np.int32)
        wts = np.ascontiguousarray(low.group_weights, dtype=np.float64)
        cfg.group_offsets = offs.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.group_atoms = atoms.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.group_weights = wts.ctypes.data_as(C.POINTER(C.c_double))
        nb = low.num_boundaries
        cfg.num_boundaries = nb
        flat = (L.Boundary * max(1, nb * num_anchors))()
        for a in range(num_anchors):
            for i, b in enumerate(low.boundaries[a]):
                flat[a * nb + i] = b
        cfg.boundaries = flat
        cfg.ring_capacity = self.ringCapacity
        cfg.save_state_slots = self.saveStateSlots if (self.saveStateFileName and self.KIND != L.KIND_LANGEVIN) else 0
        self._keep_refs = [offs, atoms, wts, flat]
        self._fill_config(cfg, num_anchors)
        return cfg

    def _initialize(self, ctx) -> None:
        """``Integrator::initialize`` -> ``kernel.initialize`` (MmvtLangevinIntegrator.cpp:73-78)."""
        self.context = ctx
        cfg = self.makeConfig(ctx.system, ctx.precision, ctx.num_anchors, ctx.device_index, ctx.lowered)
        self._config = cfg
        if self._handle:
            L.lib.seekr2_b200_destroy(self._handle)
            self._handle = C.c_void_p()
        L.check(L.lib.seekr2_b200_create(C.byref(cfg), C.byref(self._handle)))
        L.check(L.lib.seekr2_b200_set_rng(self._handle, self.randomNumberSeed, self.rngAnchorBase, self.rngAnchorStride))
        self._steps_issued = 0
        self._prev = (None, None, None)
        self._graph_key = None
        self._records = []
        self._write_headers()

    def cleanup(self) -> None:
        if self._handle:
            L.lib.seekr2_b200_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.cleanup()
        except Exception:
            pass

    # -- files ---------------------------------------------------------------------------------------
    def outputFileNames(self) -> List[str]:
        if not self.outputFileName:
            return []
        if isinstance(self.outputFileName, (list, tuple)):
            return list(self.outputFileName)
        A = self.context.num_anchors if self.context is not None else 1
        return [self.outputFileName] if A == 1 else [f"{self.outputFileName}.{a}" for a in range(A)]

    def _write_headers(self) -> None:
        if self.KIND == L.KIND_LANGEVIN:
            return
        for name in self.outputFileNames():
            L.check(L.lib.seekr2_b200_write_header(self.KIND, name.encode()))

    def _labels(self, anchor: int) -> List[int]:
        return []

    # -- stepping ------------------------------------------------------------------------------------
    def _update_params(self) -> None:
        key = (self.temperature, self.friction, self.stepSize)
        if key != self._prev:      # CudaSeekr2Kernels.cpp:188
            L.check(L.lib.seekr2_b200_set_params(self._handle, *key))
            self._prev = key
            self._graph_key = None

    def _prepare_random(self, steps_needed: int) -> int:
        """``integration.prepareRandomNumbers(paddedNumAtoms)`` (CudaSeekr2Kernels.cpp:214): returns
        the pool index of the next step.  Three sources of Gaussian numbers:
          - injected pool (``context.setRandomPool``): parity tests, "identical injected random numbers";
          - ``useRandomPool``: a device pool refilled with the Philox harness kernel when exhausted;
          - default: no pool at all, the step kernels generate the numbers themselves (same Philox)."""
        ctx = self.context
        Np = ctx.padded_num_atoms
        if ctx.random is None and not self.useRandomPool:
            return 0                              # in-kernel generator: the index is unused
        if ctx.random is None:
            ctx.random = torch.empty((ctx.num_anchors, ctx.random_pool_steps * Np, 4), dtype=torch.float32, device=ctx.device)
            ctx.random_cursor = ctx.random_pool_steps
        if ctx.random_cursor + steps_needed > ctx.random_pool_steps:
            if ctx.random_injected:
                raise L.Seekr2Error(L.ERR_INVALID, "injected random pool exhausted")
            L.check(L.lib.seekr2_b200_harness_fill_pool(ctx.random.data_ptr(), ctx.num_anchors, Np, int(ctx.random.shape[1]),
                                                        ctx.random_pool_steps, self.randomNumberSeed, self.rngAnchorBase, self.rngAnchorStride,
                                                        self._steps_issued, ctx.stream_ptr))
            ctx.random_cursor = 0
        return ctx.random_cursor * Np

    def _before_first_step(self) -> None:
        ctx = self.context
        if self.stepMode == "fused" and ctx.groups_dirty and self.KIND != L.KIND_LANGEVIN:
            L.check(L.lib.seekr2_b200_sync_groups(self._handle, C.byref(ctx.buffers()), ctx.stream_ptr))
            ctx.groups_dirty = False

    def _execute(self, bufs: L.Buffers, random_index: int) -> None:
        """``kernel.execute(context, integrator)`` for one step."""
        ctx = self.context
        if self.stepMode == "fused":
            L.check(L.lib.seekr2_b200_step(self._handle, C.byref(bufs), random_index, ctx.stream_ptr))
        else:
            L.check(L.lib.seekr2_b200_kick(self._handle, C.byref(bufs), random_index, ctx.stream_ptr))
            if self.SCHEME == L.SCHEME_MIDDLE:
                # integration.applyVelocityConstraints(tol) (CudaSeekr2Kernels.cpp:770): none in the harness
                L.check(L.lib.seekr2_b200_middle_drift(self._handle, C.byref(bufs), random_index, ctx.stream_ptr))
            # integration.applyConstraints(tol): the harness systems have no constraints
            L.check(L.lib.seekr2_b200_finish(self._handle, C.byref(bufs), ctx.stream_ptr))
            ctx.groups_dirty = True

    def step(self, steps: int) -> None:
        """``Integrator::step(n)``.  The n steps are queued in chunks of ``drainInterval``; after each chunk an
        asynchronous drain is queued behind it (``seekr2_b200_drain_begin``) and the NEXT chunk is queued before the
        host collects that drain, so appending to the crossings file overlaps the GPU's work on the next chunk and
        the rings never overflow however large n is.  With a state file name set the chunk is ``saveStateSlots``
        steps and each drain is collected at once: a step saves at most one crossing state per anchor, so none can
        be overwritten before it is fetched (the reference writes the file inside the step,
        CudaSeekr2Kernels.cpp:282-293)."""
        if self.context is None:
            raise L.Seekr2Error(L.ERR_INVALID, "This Integrator is not bound to a context!")
        ctx = self.context
        self._update_params()
        if self.stepMode == "two_pass":
            middle = self.SCHEME == L.SCHEME_MIDDLE
            ctx.ensure_two_pass_buffers(self.KIND == L.KIND_MMVT and (self.variant == L.VARIANT_REFERENCE or middle), middle)
        ctx.prepareForces()
        self._first_step_check()
        self._before_first_step()
        saving = bool(self.saveStateFileName) and self.KIND != L.KIND_LANGEVIN
        chunk = max(1, self.saveStateSlots if saving else self.drainInterval)
        done, pending, new = 0, False, []
        while done < steps:
            n = min(chunk, steps - done)
            self._issue(n)
            done += n
            if self.KIND == L.KIND_LANGEVIN:
                continue
            if pending:
                new += self._drain_collect()          # the previous chunk's records; the GPU is busy with this chunk
            self._drain_begin()
            pending = True
            if saving:
                new += self._drain_collect()
                pending = False
        if pending:
            new += self._drain_collect()
        self._after_steps(new)

    def _issue(self, steps: int) -> None:
        """Queue ``steps`` steps on the context's stream (graph replays where possible); never blocks."""
        ctx = self.context
        done = 0
        if self.useGraph and self.stepMode == "fused" and self.stepsPerGraph >= 4:
            S = self.stepsPerGraph // 4 * 4        # the launch phase (mod 4) is baked into every captured launch
            while steps - done >= S:
                ri = self._prepare_random(S)
                key = (ri, ctx.random.data_ptr() if ctx.random is not None else 0, self._prev, self._steps_issued % 4)   # launch phase is baked in
                if self._graph_key != key:
                    self._capture(S, ri)
                    self._graph_key = key
                L.check(L.lib.seekr2_b200_graph_launch(self._handle, ctx.stream_ptr))
                ctx.random_cursor += S
                self._steps_issued += S
                done += S
        while done < steps:
            ri = self._prepare_random(1)
            if ctx.has_force_provider:
                ctx.computeForces()
            self._execute(ctx.buffers(), ri)
            ctx.random_cursor += 1
            self._steps_issued += 1
            done += 1

    def _capture(self, S: int, random_index0: int) -> None:
        ctx = self.context
        if S % 4:
            raise L.Seekr2Error(L.ERR_INVALID, "stepsPerGraph must be a multiple of 4")
        Np = ctx.padded_num_atoms
        bufs = ctx.buffers()
        L.check(L.lib.seekr2_b200_graph_begin(self._handle, ctx.stream_ptr))
        try:
            for s in range(S):
                if ctx.has_force_provider:
                    ctx.computeForces()
                L.check(L.lib.seekr2_b200_step(self._handle, C.byref(bufs), random_index0 + s * Np, ctx.stream_ptr))
        finally:
            rc = L.lib.seekr2_b200_graph_end(self._handle, ctx.stream_ptr)
        L.check(rc)

    def _first_step_check(self) -> None:
        pass

    def _after_steps(self, new) -> None:
        pass

    # -- results -------------------------------------------------------------------------------------
    _DRAIN_CHUNK = 1 << 14

    def _drain_begin(self) -> None:
        L.check(L.lib.seekr2_b200_drain_begin(self._handle, self._DRAIN_CHUNK, self.context.stream_ptr))

    def _drain_collect(self) -> List[L.Record]:
        """Collect the drain in flight (and whatever did not fit into it), append the records to the output
        file(s), write the crossing states.  Waits only for the work queued before ``_drain_begin``."""
        if getattr(self, "_drain_buf", None) is None:
            self._drain_buf = (L.Record * self._DRAIN_CHUNK)()
        buf = self._drain_buf
        names = self.outputFileNames()
        new: List[L.Record] = []
        status = L.OK
        while True:
            n, waiting = C.c_int64(0), C.c_int64(0)
            rc = L.lib.seekr2_b200_drain_end(self._handle, buf, self._DRAIN_CHUNK, C.byref(n), C.byref(waiting))
            if rc not in (L.OK, L.ERR_RING_OVERFLOW):
                L.check(rc)
            status = status or rc
            got = [self._copy_record(buf[i]) for i in range(n.value)]
            new.extend(got)
            if names and got:
                for a in sorted({r.anchor for r in got}):
                    labels = self._labels(a)
                    L.check(L.lib.seekr2_b200_append_records(self.KIND, names[a].encode(), buf, n.value, a,
                                                             _i32_array(labels), len(labels), self._num_src()))
            if waiting.value <= 0:
                break
            self._drain_begin()
        self._records.extend(new)
        self._save_states(new)
        if status != L.OK:
            raise L.Seekr2Error(status, "crossing ring overflowed: records were lost; drain more often or raise ringCapacity")
        return new

    def drain(self) -> List[L.Record]:
        """Fetch new crossing records and append them to the output file(s) (blocks until every queued step ran)."""
        self._drain_begin()
        return self._drain_collect()

    def stateFileName(self, record: L.Record) -> str:
        """``saveStateFileName + "_" + counter + "_" + milestoneGroup`` (CudaSeekr2Kernels.cpp:285-287, :578-580, :838-840)."""
        labels = self._labels(record.anchor)
        idx = record.index + (self._num_src() if record.kind in (L.REC_ELBER_DEST, L.REC_ELBER_DEST_STAR) else 0)
        prefix = self.saveStateFileName
        if self.context is not None and self.context.num_anchors > 1:
            prefix = f"{prefix}.{record.anchor}"
        if self.KIND == L.KIND_ELBER and self.SCHEME == L.SCHEME_MIDDLE:
            return f"{prefix}_{labels[idx]}"        # the ElberLangevinMiddle kernel leaves the counter out (:1157)
        return f"{prefix}_{record.counter}_{labels[idx]}"

    def fetchCrossingState(self, record: L.Record):
        """(positions[N,3], velocities[N,3]) as they were when ``record``'s surface was crossed, before
        the bounce -- what the reference's getState(Positions | Velocities) returns at :283 / :576."""
        n = self.context.num_atoms
        x, v = np.zeros((n, 3)), np.zeros((n, 3))
        dp = C.POINTER(C.c_double)
        L.check(L.lib.seekr2_b200_fetch_snapshot(self._handle, record.anchor, record.snapshot, x.ctypes.data_as(dp),
                                                 v.ctypes.data_as(dp), self.context.stream_ptr))
        return x, v

    def _save_states(self, records) -> None:
        if not self.saveStateFileName:
            return
        seen = set()
        for r in records:
            if r.snapshot <= 0:
                continue
            # Elber: one state per ending step, named after the last surface of that step (:565, :578)
            key = (r.anchor, r.snapshot)
            if key in seen:
                continue
            seen.add(key)
            x, v = self.fetchCrossingState(r)
            dp = C.POINTER(C.c_double)
            box = self.context.system.getDefaultPeriodicBoxVectors() if hasattr(self.context.system, "getDefaultPeriodicBoxVectors") else None
            boxp = None
            if box is not None:
                boxa = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
                boxp = boxa.ctypes.data_as(dp)
            L.check(L.lib.seekr2_b200_write_state_xml(self.stateFileName(r).encode(), x.shape[0], x.ctypes.data_as(dp),
                                                      v.ctypes.data_as(dp), float(r.time), boxp))

    @staticmethod
    def _copy_record(r: L.Record) -> L.Record:
        out = L.Record()
        C.memmove(C.byref(out), C.byref(r), C.sizeof(L.Record))
        return out

    def _num_src(self) -> int:
        return 0

    def getRecords(self, anchor: Optional[int] = None) -> List[L.Record]:
        return [r for r in self._records if anchor is None or r.anchor == anchor]

    def getAnchorState(self, anchor: int = 0) -> L.AnchorState:
        st = L.AnchorState()
        L.check(L.lib.seekr2_b200_get_state(self._handle, anchor, C.byref(st), self.context.stream_ptr))
        return st


class LangevinIntegrator(_B200LangevinBase):
    """Plain Langevin step on the same kernels (KIND_LANGEVIN): the comparator for the north-star
    target "the fused MMVT step adds no more than 5% over a plain Langevin step"."""
    KIND = L.KIND_LANGEVIN


class MmvtLangevinIntegrator(_B200LangevinBase):
    """``Seekr2Plugin::MmvtLangevinIntegrator`` (MmvtLangevinIntegrator.h:60-226)."""
    KIND = L.KIND_MMVT
    KERNEL_NAME = "IntegrateMmvtLangevinStep"     # Seekr2Kernels.h:55-57

    def __init__(self, temperature, frictionCoeff, stepSize, fileName):
        super().__init__(temperature, frictionCoeff, stepSize, fileName)
        self.saveStatisticsFileName = ""
        self.milestoneGroups: List[int] = []
        self.anchorMilestoneGroups = {}
        self.bounceCounter = 0
        self.anchorBounceCounters = None

    def getSaveStatisticsFileName(self): return self.saveStatisticsFileName
    def setSaveStatisticsFileName(self, fileName): self.saveStatisticsFileName = fileName

    def getMilestoneGroup(self, index: int) -> int:
        if index < 0 or index >= len(self.milestoneGroups):      # ASSERT_VALID_INDEX
            raise L.Seekr2Error(L.ERR_INVALID, "Index out of range")
        return self.milestoneGroups[index]

    def getNumMilestoneGroups(self) -> int:
        return len(self.milestoneGroups)

    def addMilestoneGroup(self, milestoneGroup: int) -> int:
        self.milestoneGroups.append(int(milestoneGroup))
        return len(self.milestoneGroups) - 1

    def setAnchorMilestoneGroups(self, anchor: int, groups: Sequence[int]) -> None:
        """Labels written for anchor ``anchor`` (each Voronoi cell has its own neighbour ids)."""
        self.anchorMilestoneGroups[int(anchor)] = [int(g) for g in groups]

    def getBounceCounter(self): return self.bounceCounter
    def setBounceCounter(self, counter): self.bounceCounter = int(counter)
    def setAnchorBounceCounters(self, counters): self.anchorBounceCounters = [int(c) for c in counters]

    def _labels(self, anchor):
        return self.anchorMilestoneGroups.get(anchor, self.milestoneGroups)

    def _fill_config(self, cfg, num_anchors):
        cfg.num_milestone_groups = len(self.milestoneGroups)
        counters = self.anchorBounceCounters or [self.bounceCounter] * num_anchors
        arr = _i32_array(counters)
        self._keep_refs.append(arr)
        cfg.bounce_counter0 = arr

    def _first_step_check(self):
        # CudaSeekr2Kernels.cpp:205-210: only while stepCount <= 0
        ctx = self.context
        if getattr(self, "_first_checked", None) is self._handle.value:
            return
        if self.getAnchorState(0).step_count <= 0:
            L.check(L.lib.seekr2_b200_check_first_step(self._handle, C.byref(ctx.buffers()), None, ctx.stream_ptr))
            ctx.groups_dirty = False
        self._first_checked = self._handle.value

    def _after_steps(self, new):
        if self.saveStatisticsFileName and new:            # :313-345
            ctx = self.context
            for a in range(ctx.num_anchors):
                name = self.saveStatisticsFileName if ctx.num_anchors == 1 else f"{self.saveStatisticsFileName}.{a}"
                st = self.getAnchorState(a)
                labels = self._labels(a)
                L.check(L.lib.seekr2_b200_write_statistics(name.encode(), C.byref(st), _i32_array(labels), len(labels)))


class LangevinMiddleIntegrator(_B200LangevinBase):
    """Plain LangevinMiddle step (comparator for the Middle MMVT step)."""
    SCHEME = L.SCHEME_MIDDLE


class MmvtLangevinMiddleIntegrator(MmvtLangevinIntegrator):
    """``Seekr2Plugin::MmvtLangevinMiddleIntegrator`` -- what the reference's examples instantiate
    (mmvt_demonstration3_tryp_benz_spherical.py:59).  Kernel: langevinMiddle.cu:7-101; a bounce negates
    the velocity saved after the force kick (:41, :202-214)."""
    SCHEME = L.SCHEME_MIDDLE
    KERNEL_NAME = "IntegrateMmvtLangevinMiddleStep"     # Seekr2Kernels.h:117-119


class ElberLangevinIntegrator(_B200LangevinBase):
    """``Seekr2Plugin::ElberLangevinIntegrator`` (ElberLangevinIntegrator.h:60-232)."""
    KIND = L.KIND_ELBER
    KERNEL_NAME = "IntegrateElberLangevinStep"    # Seekr2Kernels.h:86-88

    def __init__(self, temperature, frictionCoeff, stepSize, fileName):
        super().__init__(temperature, frictionCoeff, stepSize, fileName)
        self.srcMilestoneGroups: List[int] = []
        self.destMilestoneGroups: List[int] = []
        self.crossingCounter = 0
        self.endOnSrcMilestone = False     # the reference leaves this uninitialised

    def getSrcMilestoneGroup(self, index):
        if index < 0 or index >= len(self.srcMilestoneGroups):
            raise L.Seekr2Error(L.ERR_INVALID, "Index out of range")
        return self.srcMilestoneGroups[index]

    def getNumSrcMilestoneGroups(self): return len(self.srcMilestoneGroups)

    def getDestMilestoneGroup(self, index):
        if index < 0 or index >= len(self.destMilestoneGroups):
            raise L.Seekr2Error(L.ERR_INVALID, "Index out of range")
        return self.destMilestoneGroups[index]

    def getNumDestMilestoneGroups(self): return len(self.destMilestoneGroups)

    def addSrcMilestoneGroup(self, group):
        self.srcMilestoneGroups.append(int(group))
        return len(self.srcMilestoneGroups) - 1

    def addDestMilestoneGroup(self, group):
        self.destMilestoneGroups.append(int(group))
        return len(self.destMilestoneGroups) - 1

    def getCrossingCounter(self): return self.crossingCounter
    def setCrossingCounter(self, counter): self.crossingCounter = int(counter)
    def getEndOnSrcMilestone(self): return self.endOnSrcMilestone
    def setEndOnSrcMilestone(self, endOnSrc): self.endOnSrcMilestone = bool(endOnSrc)

    def _labels(self, anchor):
        return self.srcMilestoneGroups + self.destMilestoneGroups

    def _num_src(self):
        return len(self.srcMilestoneGroups)

    def _fill_config(self, cfg, num_anchors):
        src, dst = _i32_array(self.srcMilestoneGroups), _i32_array(self.destMilestoneGroups)
        cc = _i32_array([self.crossingCounter] * num_anchors)
        self._keep_refs += [src, dst, cc]
        cfg.num_src, cfg.src_groups = len(self.srcMilestoneGroups), src
        cfg.num_dest, cfg.dest_groups = len(self.destMilestoneGroups), dst
        cfg.end_on_src_milestone = 1 if self.endOnSrcMilestone else 0
        cfg.crossing_counter0 = cc


class ElberLangevinMiddleIntegrator(ElberLangevinIntegrator):
    """``Seekr2Plugin::ElberLangevinMiddleIntegrator`` (langevinMiddle.cu:104-196; host logic identical to
    the non-Middle Elber kernel, CudaSeekr2Kernels.cpp:1098-1173)."""
    SCHEME = L.SCHEME_MIDDLE
    KERNEL_NAME = "IntegrateElberLangevinMiddleStep"    # Seekr2Kernels.h:148-150
