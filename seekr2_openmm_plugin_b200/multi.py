"""Anchor-parallel multi-GPU driver pieces.

MMVT anchors (Voronoi cells) and Elber milestones are independent trajectories: the per-step path has
NO collective.  Anchors are dealt round-robin to ranks (one process per GPU); the only exchange is the
end-of-run gather of per-anchor crossing statistics (N_alpha_beta, N_ij_alpha, R_i_alpha, T_alpha,
steps, crossings), a few KB per anchor, done with one ``all_gather`` over NCCL (NVLink) -- or gloo in
the CPU tests.  The reference has no multi-GPU support at all (CudaSeekr2KernelFactory.cpp:74 hard-wires
``contexts[0]``).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import _lib as L

M = L.MAX_MILESTONES
STAT_WIDTH = 6 + M + M * M + M      # anchor id, steps, bounces, records, time, T_alpha | N_ab | Nij | Ri


def partition_anchors(num_anchors: int, world_size: int, rank: int) -> List[int]:
    """anchor a -> rank a mod world_size."""
    return [a for a in range(num_anchors) if a % world_size == rank]


def pack_state(global_anchor: int, st: L.AnchorState) -> np.ndarray:
    row = np.zeros(STAT_WIDTH, dtype=np.float64)
    row[0:6] = [global_anchor, st.step_count, st.bounce_counter, st.ring_head, st.time, st.T_alpha]
    row[6:6 + M] = st.N_alpha_beta[:]
    row[6 + M:6 + M + M * M] = st.Nij_alpha[:]
    row[6 + M + M * M:] = st.Ri_alpha[:]
    return row


def unpack_row(row: np.ndarray) -> Dict:
    return {"anchor": int(row[0]), "steps": int(row[1]), "bounce_counter": int(row[2]), "records": int(row[3]),
            "time": float(row[4]), "T_alpha": float(row[5]), "N_alpha_beta": row[6:6 + M].astype(np.int64),
            "Nij_alpha": row[6 + M:6 + M + M * M].astype(np.int64).reshape(M, M), "Ri_alpha": row[6 + M + M * M:].copy()}


def gather_rows(local_rows: np.ndarray, world_size: int, device=None) -> np.ndarray:
    """all_gather of [n_local, STAT_WIDTH] float64 blocks; returns rows sorted by anchor.  Ranks may hold different numbers
    of anchors (round-robin leaves the last ranks one short when num_anchors % world_size != 0): blocks are padded to the
    largest count with rows of anchor id -1, which are dropped after the gather."""
    if world_size <= 1:
        out = local_rows
    else:
        import torch
        import torch.distributed as dist
        count = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=device)
        dist.all_reduce(count, op=dist.ReduceOp.MAX)
        padded = np.full((int(count.item()), STAT_WIDTH), -1.0, dtype=np.float64)
        padded[: local_rows.shape[0]] = local_rows
        t = torch.from_numpy(padded)
        if device is not None:
            t = t.to(device)
        parts = [torch.empty_like(t) for _ in range(world_size)]
        dist.all_gather(parts, t)
        out = torch.cat(parts).cpu().numpy()
        out = out[out[:, 0] >= 0]
    return out[np.argsort(out[:, 0], kind="stable")]


def gather_statistics(integrator, world_size: int, global_anchor_ids=None) -> int:
    """End-of-run gather for an integrator whose context holds this rank's anchors.  Returns the number
    of anchors gathered and stores the merged table on ``integrator.gathered`` (rank-independent)."""
    ctx = integrator.context
    A = ctx.num_anchors
    rank = 0
    if world_size > 1:
        import torch.distributed as dist
        rank = dist.get_rank()
    ids = global_anchor_ids if global_anchor_ids is not None else [a * world_size + rank for a in range(A)]
    rows = np.stack([pack_state(ids[a], integrator.getAnchorState(a)) for a in range(A)])
    merged = gather_rows(rows, world_size, ctx.device if world_size > 1 else None)
    integrator.gathered = [unpack_row(r) for r in merged]
    return len(integrator.gathered)
