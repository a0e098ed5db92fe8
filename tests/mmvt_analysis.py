"""Minimal MMVT kinetics analysis -- TEST INFRASTRUCTURE.

The plugin's product is the crossings file (README.md:383-393 of the reference: boundary id, bounce index, time);
the rate constants are derived downstream by the separate `seekr2` package (not in the reference tree, absent
here).  This module restates the published MMVT equations (Vanden-Eijnden & Venturoli 2009; Votapka et al.,
SEEKR2, 2022) so that long-run runs of the B200 path and of the CPU oracle can be compared on the quantity users
care about (an MFPT / k_off), not only on raw counts:

  per anchor a:  k_ab = N_ab / T_a                     (bounces against the boundary shared with anchor b)
  stationary pi: sum_b pi_b k_ba = pi_a sum_b k_ab,  sum pi = 1
  milestones:    N_ij = sum_a pi_a N^a_ij / T_a,  R_i = sum_a pi_a R^a_i / T_a,  Q_ij = N_ij / R_i
  MFPT:          Q' tau = -1 with the absorbing milestone's row and column removed
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np


@dataclass
class AnchorStats:
    """What the reference's execute() accumulates per anchor (CudaSeekr2Kernels.cpp:294-309), keyed by
    milestone LABEL (the first column of the crossings file)."""
    N_alpha_beta: Dict[int, int] = field(default_factory=dict)
    N_ij: Dict[Tuple[int, int], int] = field(default_factory=dict)
    R_i: Dict[int, float] = field(default_factory=dict)
    T: float = 0.0

    def add(self, other: "AnchorStats") -> None:
        for k, v in other.N_alpha_beta.items():
            self.N_alpha_beta[k] = self.N_alpha_beta.get(k, 0) + v
        for k, v in other.N_ij.items():
            self.N_ij[k] = self.N_ij.get(k, 0) + v
        for k, v in other.R_i.items():
            self.R_i[k] = self.R_i.get(k, 0.0) + v
        self.T += other.T


def parse_crossings(path: str, dt: float) -> AnchorStats:
    """Replays the statistics block of the reference on a crossings file.  `time` on a line is the time of the step
    that bounced; incubationTime counts whole steps since the last change of milestone (:303,:364), i.e. the
    difference of the logged times."""
    st = AnchorStats()
    prev, prev_change_time, first_time, last_time = None, 0.0, 0.0, 0.0
    for line in open(path):
        if line.startswith("#") or not line.strip():
            continue
        label_s, _counter, time_s = line.strip().split(",")
        label, time = int(label_s), float(time_s)
        if prev is not None:
            st.N_alpha_beta[label] = st.N_alpha_beta.get(label, 0) + 1            # :294-296
        if prev != label:
            if prev is not None:                                                   # :298-300
                st.N_ij[(prev, label)] = st.N_ij.get((prev, label), 0) + 1
                st.R_i[prev] = st.R_i.get(prev, 0.0) + (time - prev_change_time)
            else:
                first_time = time                                                  # :302
            prev_change_time = time                                                # incubationTime = 0 (:304)
        last_time = time
        prev = label
    st.T = last_time - first_time                                                  # :306
    return st


def stats_from_state(state, labels: List[int], max_milestones: int) -> AnchorStats:
    """The same quantities from the engine's device-resident anchor state."""
    st = AnchorStats()
    m = len(labels)
    for i in range(m):
        if state.N_alpha_beta[i]:
            st.N_alpha_beta[labels[i]] = int(state.N_alpha_beta[i])
        if state.Ri_alpha[i]:
            st.R_i[labels[i]] = float(state.Ri_alpha[i])
        for j in range(m):
            n = state.Nij_alpha[i * max_milestones + j]
            if n:
                st.N_ij[(labels[i], labels[j])] = int(n)
    st.T = float(state.T_alpha)
    return st


def stationary_distribution(k: np.ndarray) -> np.ndarray:
    """pi with sum_b pi_b k[b, a] = pi_a sum_b k[a, b]."""
    n = k.shape[0]
    m = k.T - np.diag(k.sum(axis=1))
    m[-1, :] = 1.0
    rhs = np.zeros(n)
    rhs[-1] = 1.0
    return np.linalg.solve(m, rhs)


def mfpt(q: np.ndarray, start: int, absorbing: int) -> float:
    keep = [i for i in range(q.shape[0]) if i != absorbing]
    tau = np.linalg.solve(q[np.ix_(keep, keep)], -np.ones(len(keep)))
    return float(tau[keep.index(start)])


def shell_mfpt(anchors: List[AnchorStats], start_milestone: int, end_milestone: int) -> dict:
    """Concentric-shell (1-D chain) MMVT model: anchor a lies between milestone a (inner) and a + 1 (outer);
    crossing milestone a + 1 from anchor a leads to anchor a + 1."""
    n_anchor = len(anchors)
    k = np.zeros((n_anchor, n_anchor))
    for a, st in enumerate(anchors):
        for label, n in st.N_alpha_beta.items():
            b = a - 1 if label == a else a + 1
            if 0 <= b < n_anchor:
                k[a, b] = n / st.T
    pi = stationary_distribution(k)
    n_m = n_anchor + 1
    N = np.zeros((n_m, n_m))
    R = np.zeros(n_m)
    for a, st in enumerate(anchors):
        for (i, j), n in st.N_ij.items():
            N[i, j] += pi[a] * n / st.T
        for i, r in st.R_i.items():
            R[i] += pi[a] * r / st.T
    q = np.zeros((n_m, n_m))
    for i in range(n_m):
        if R[i] > 0:
            q[i] = N[i] / R[i]
        q[i, i] = -(q[i].sum() - q[i, i])
    # milestones nobody can leave through (the outermost / innermost wall) are reflecting ends of the chain
    live = [i for i in range(n_m) if R[i] > 0 or i == end_milestone]
    idx = {m: n for n, m in enumerate(live)}
    ql = q[np.ix_(live, live)]
    return {"pi": pi, "k": k, "Q": q, "mfpt": mfpt(ql, idx[start_milestone], idx[end_milestone])}
