"""Anchor partitioning and the end-of-run statistics gather, world_size 2 over gloo on the CPU
(the N > 1 path of bench.py / multi.py; on the GPU box the same code runs over NCCL)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import helpers as H
from helpers import L
from seekr2_openmm_plugin_b200 import multi


def test_partition_is_round_robin_and_complete():
    for n, w in ((8, 1), (8, 2), (13, 4), (3, 8)):
        parts = [multi.partition_anchors(n, w, r) for r in range(w)]
        assert sorted(a for p in parts for a in p) == list(range(n))
        assert all(a % w == r for r, p in enumerate(parts) for a in p)


def _fake_state(anchor):
    st = L.AnchorState()
    st.step_count, st.bounce_counter, st.ring_head, st.time, st.T_alpha = 100 + anchor, anchor, 2 * anchor, 0.2 + anchor, 0.1 * anchor
    st.N_alpha_beta[0], st.N_alpha_beta[1] = anchor, anchor + 1
    st.Nij_alpha[0 * L.MAX_MILESTONES + 1] = 7 * anchor
    st.Ri_alpha[1] = 0.5 * anchor
    return st


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = multi.partition_anchors(7, world, rank)                  # uneven: rank 0 holds 4 anchors, rank 1 holds 3
    rows = np.stack([multi.pack_state(a, _fake_state(a)) for a in mine])
    merged = multi.gather_rows(rows, world)
    np.save(os.path.join(out_dir, f"merged{rank}.npy"), merged)
    dist.destroy_process_group()


def test_gather_statistics_world_size_2(tmp_path):
    port = 29500 + os.getpid() % 1000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    m0, m1 = np.load(tmp_path / "merged0.npy"), np.load(tmp_path / "merged1.npy")
    np.testing.assert_array_equal(m0, m1)                       # every rank ends with the same table
    assert list(m0[:, 0]) == [0, 1, 2, 3, 4, 5, 6]              # sorted by global anchor id, padding rows dropped
    for a, row in enumerate(m0):
        d = multi.unpack_row(row)
        assert (d["anchor"], d["steps"], d["bounce_counter"], d["records"]) == (a, 100 + a, a, 2 * a)
        assert d["N_alpha_beta"][1] == a + 1 and d["Nij_alpha"][0, 1] == 7 * a and d["Ri_alpha"][1] == 0.5 * a
        assert d["time"] == 0.2 + a and d["T_alpha"] == 0.1 * a


def test_single_rank_gather_is_identity():
    rows = np.stack([multi.pack_state(a, _fake_state(a)) for a in (2, 0, 1)])
    merged = multi.gather_rows(rows, 1)
    assert list(merged[:, 0]) == [0, 1, 2]


def test_rng_streams_are_keyed_by_the_global_anchor_id():
    """One seed for the whole job: local anchor j of rank r (round-robin) must draw from stream r + j * world -- the global
    anchor id -- so that no two anchors of the job share a stream (base + j alone collides: rank 0 local 1 == rank 1 local 0)."""
    world, n = 4, 13
    seen = {}
    for rank in range(world):
        mine = multi.partition_anchors(n, world, rank)
        base, stride = rank, world                                 # what bench.py sets on the integrator
        for j, a in enumerate(mine):
            stream = base + j * stride
            assert stream == a
            assert stream not in seen, (stream, seen[stream], (rank, j))
            seen[stream] = (rank, j)
    assert sorted(seen) == list(range(n))
