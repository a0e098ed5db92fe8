"""Harness stand-ins for the pieces of OpenMM the SEEKR2 integrators talk to.

OpenMM is not available in this environment, so the tests and benchmarks drive the B200 kernels
through a small mirror of the OpenMM objects the reference's examples use
(``seekr2plugin/examples/mmvt_demonstration*.py``): a ``System`` holding particle masses and forces,
a ``HarmonicBondForce`` (the reference's ``testSingleBond``), a harmonic ``TetherForce`` (the
synthetic force model of SURVEY.md 8d) and ``CustomCentroidBondForce`` restricted to the expression
forms the examples actually write, which are recognised and turned into typed boundary descriptors
(``include/seekr2_b200.h``: ``seekr2_b200_boundary``).  Anything that cannot be typed raises -- there
is no generic expression evaluator and no CPU evaluation path.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _lib as L


class System:
    """Mirror of ``OpenMM::System`` as far as the integrators need it (masses + forces)."""

    def __init__(self):
        self._masses: List[float] = []
        self._forces: list = []
        self._constraints: list = []
        self._box = np.array([[2.0, 0, 0], [0, 2.0, 0], [0, 0, 2.0]])     # OpenMM's default box

    def getDefaultPeriodicBoxVectors(self) -> np.ndarray:
        return self._box.copy()

    def setDefaultPeriodicBoxVectors(self, a, b, c) -> None:
        self._box = np.array([a, b, c], dtype=np.float64).reshape(3, 3)

    def addParticle(self, mass: float) -> int:
        self._masses.append(float(mass))
        return len(self._masses) - 1

    def getNumParticles(self) -> int:
        return len(self._masses)

    def getParticleMass(self, index: int) -> float:
        return self._masses[index]

    def setParticleMass(self, index: int, mass: float) -> None:
        self._masses[index] = float(mass)

    def addForce(self, force) -> int:
        self._forces.append(force)
        return len(self._forces) - 1

    def getNumForces(self) -> int:
        return len(self._forces)

    def getForce(self, index: int):
        return self._forces[index]

    def getNumConstraints(self) -> int:
        return len(self._constraints)

    @property
    def masses(self) -> np.ndarray:
        return np.asarray(self._masses, dtype=np.float64)


class Force:
    def __init__(self):
        self._force_group = 0

    def setForceGroup(self, group: int) -> None:
        self._force_group = int(group)

    def getForceGroup(self) -> int:
        return self._force_group


class HarmonicBondForce(Force):
    """E = k/2 (r - r0)^2 per bond; evaluated by ``seekr2_b200_harness_bond_force``."""

    def __init__(self):
        super().__init__()
        self.bonds: List[tuple] = []

    def addBond(self, particle1: int, particle2: int, length: float, k: float) -> int:
        self.bonds.append((int(particle1), int(particle2), float(length), float(k)))
        return len(self.bonds) - 1


class TetherForce(Force):
    """F = -k (x - x0): the synthetic force model of the benchmarks.

    ``x0`` defaults to the positions the context holds the first time forces are computed."""

    def __init__(self, k: float, reference_positions: Optional[np.ndarray] = None):
        super().__init__()
        self.k = float(k)
        self.reference_positions = None if reference_positions is None else np.asarray(reference_positions, dtype=np.float64)


@dataclass
class _TypedExpression:
    cv_type: int
    style: int
    axis: int = 0
    k_sign: float = 1.0          # "-k*(x1-bound)" flips the sign of k
    threshold_name: str = "radius"
    center_names: Sequence[str] = ()


def _recognise(expression: str, num_groups: int) -> _TypedExpression:
    """Map one of the reference examples' expression forms onto a typed CV."""
    e = re.sub(r"\s+", "", expression)
    style = L.STYLE_RAW
    m = re.fullmatch(r"bitcode\*step\((.*)\)", e)
    if m:
        style = L.STYLE_STEP
        e = m.group(1)
    sign = 1.0
    if e.startswith("-"):
        sign, e = -1.0, e[1:]
    m = re.fullmatch(r"k\*\((.*)\)", e)
    if not m:
        raise L.Seekr2Error(L.ERR_INVALID, f"cannot type boundary expression {expression!r}")
    u = m.group(1)
    if num_groups == 2 and re.fullmatch(r"distance\(g1,g2\)\^2-(\w+)\^2", u):
        name = re.fullmatch(r"distance\(g1,g2\)\^2-(\w+)\^2", u).group(1)
        return _TypedExpression(L.CV_SPHERE_PAIR, style, k_sign=sign, threshold_name=name)
    m = re.fullmatch(r"\(x1-(\w+)\)\^2\+\(y1-(\w+)\)\^2\+\(z1-(\w+)\)\^2-(\w+)\^2", u)
    if num_groups == 1 and m:
        return _TypedExpression(L.CV_SPHERE_FIXED, style, k_sign=sign, threshold_name=m.group(4),
                                center_names=(m.group(1), m.group(2), m.group(3)))
    m = re.fullmatch(r"([xyz])1-(\w+)", u)
    if num_groups == 1 and m:
        return _TypedExpression(L.CV_PLANE, style, axis="xyz".index(m.group(1)), k_sign=sign, threshold_name=m.group(2))
    m = re.fullmatch(r"distance\(g1,g2\)\^2((?:\+distance\(g\d,g\d\)\^2)*)-(\w+)\^2", u)
    if m and num_groups in (4, 6):
        pairs = re.findall(r"distance\(g(\d),g(\d)\)", u)
        if [int(x) for p in pairs for x in p] == list(range(1, num_groups + 1)):       # demo4:27
            return _TypedExpression(L.CV_PAIR_SUM, style, axis=num_groups // 2, k_sign=sign, threshold_name=m.group(2))
    m = re.fullmatch(r"angle\(g1,g2,g3\)-(\w+)", u)
    if num_groups == 3 and m:                                                           # demo5:24
        return _TypedExpression(L.CV_ANGLE, style, k_sign=sign, threshold_name=m.group(1))
    raise L.Seekr2Error(L.ERR_INVALID, f"cannot type boundary expression {expression!r}")


class CustomCentroidBondForce(Force):
    """The subset of OpenMM's ``CustomCentroidBondForce`` that SEEKR2 milestone definitions use.

    ``bitcode*step(k*(distance(g1,g2)^2-radius^2))`` (demo3:37), the fixed-centre sphere (demo1:59),
    axis planes (demo2:61), the sum of squared pair distances (demo4:27), the three-group angle
    (demo5:24) and their older non-``step`` forms.  Per-bond parameters may be given per
    anchor (``setAnchorBondParameters``) so one engine can run many Voronoi cells of one system."""

    def __init__(self, numGroups: int, expression: str):
        super().__init__()
        self.numGroupsPerBond = int(numGroups)
        self.expression = expression
        self._typed = _recognise(expression, self.numGroupsPerBond)
        self.groups: List[tuple] = []          # (indices, weights or None)
        self.param_names: List[str] = []
        self.bonds: List[tuple] = []           # (group indices, [params])
        self.anchor_params: Dict[tuple, List[float]] = {}

    def addGroup(self, particles: Sequence[int], weights: Optional[Sequence[float]] = None) -> int:
        self.groups.append((list(int(p) for p in particles), None if weights is None else [float(w) for w in weights]))
        return len(self.groups) - 1

    def addPerBondParameter(self, name: str) -> int:
        self.param_names.append(name)
        return len(self.param_names) - 1

    def addBond(self, groups: Sequence[int], parameters: Sequence[float]) -> int:
        if len(groups) != self.numGroupsPerBond:
            raise L.Seekr2Error(L.ERR_INVALID, "wrong number of groups for a bond")
        self.bonds.append((list(int(g) for g in groups), [float(p) for p in parameters]))
        return len(self.bonds) - 1

    def setAnchorBondParameters(self, anchor: int, bond: int, parameters: Sequence[float]) -> None:
        self.anchor_params[(int(anchor), int(bond))] = [float(p) for p in parameters]

    def getNumBonds(self) -> int:
        return len(self.bonds)

    # -- lowering ---------------------------------------------------------------------------
    def typed_boundaries(self, anchor: int, group_base, remap: Optional[Sequence[int]] = None) -> List[L.Boundary]:
        """``remap[g]`` = index of this Force's group g in the lowered (de-duplicated) group table; without it the
        groups sit at ``group_base + g``."""
        out = []
        t = self._typed
        for bi, (groups, params) in enumerate(self.bonds):
            params = self.anchor_params.get((anchor, bi), params)
            if len(params) != len(self.param_names):
                raise L.Seekr2Error(L.ERR_INVALID, "wrong number of per-bond parameters")
            p = dict(zip(self.param_names, params))
            b = L.Boundary()
            b.type = t.cv_type
            b.style = t.style
            for gi, g in enumerate(groups[:6]):
                b.groups[gi] = remap[g] if remap is not None else group_base + g
            b.axis = t.axis
            b.force_group = self.getForceGroup()
            # every name of the expression must be a per-bond parameter: no silent defaults (a global or misspelt
            # parameter would give a different surface than the reference evaluates)
            needed = (["bitcode"] if t.style == L.STYLE_STEP else []) + ["k", t.threshold_name] + list(t.center_names)
            for name in needed:
                if name not in p:
                    raise L.Seekr2Error(L.ERR_INVALID, f"'{name}' in boundary expression '{self.expression}' is not a per-bond parameter of the Force")
            b.bitcode = p["bitcode"] if t.style == L.STYLE_STEP else 1.0
            b.k = t.k_sign * p["k"]
            b.threshold = p[t.threshold_name]
            for c, name in enumerate(t.center_names):
                b.center[c] = p[name]
            out.append(b)
        return out


@dataclass
class LoweredBoundaries:
    """Groups and boundary terms of a System in the layout ``seekr2_b200_config`` wants."""
    group_offsets: np.ndarray
    group_atoms: np.ndarray
    group_weights: np.ndarray
    boundaries: List[List[L.Boundary]] = field(default_factory=list)   # [anchor][boundary]

    @property
    def num_groups(self) -> int:
        return len(self.group_offsets) - 1

    @property
    def num_boundaries(self) -> int:
        return len(self.boundaries[0]) if self.boundaries else 0


def lower_boundaries(system: System, num_anchors: int) -> LoweredBoundaries:
    """Walk the System's CustomCentroidBondForces and emit typed descriptors (what the OpenMM
    adapter's expression recogniser does, platforms/b200/B200Seekr2Kernels.cpp)."""
    offsets, atoms, weights = [0], [], []
    per_anchor: List[List[L.Boundary]] = [[] for _ in range(num_anchors)]
    masses = system.masses
    seen = {}            # identical groups (same atoms, same weights) are lowered once: the reference's examples give every
    for force in system._forces:      # milestone Force its own copies of the same two groups (demo3:37-45)
        if not isinstance(force, CustomCentroidBondForce):
            continue
        remap = []
        for indices, w in force.groups:
            # OpenMM: without explicit weights a centroid group is mass weighted
            ww = [float(x) for x in (w if w is not None else [masses[i] for i in indices])]
            key = (tuple(int(i) for i in indices), tuple(ww))
            if key not in seen:
                seen[key] = len(offsets) - 1
                atoms.extend(indices)
                weights.extend(ww)
                offsets.append(len(atoms))
            remap.append(seen[key])
        for a in range(num_anchors):
            per_anchor[a].extend(force.typed_boundaries(a, 0, remap))
    return LoweredBoundaries(np.asarray(offsets, dtype=np.int32), np.asarray(atoms, dtype=np.int32),
                             np.asarray(weights, dtype=np.float64), per_anchor)
