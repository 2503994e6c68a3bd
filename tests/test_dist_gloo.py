"""world_size-2 gloo test of the N>1 plumbing: DB broadcast + query sharding cover every query exactly once."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from foldseek_b200.dist import broadcast_bytes, shard_by_residues


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    db = torch.from_numpy(rng.integers(0, 21, 5000).astype(np.uint8)) if rank == 0 else None
    got = broadcast_bytes(dist, lambda n: torch.empty(n, dtype=torch.uint8), db, 0)
    lens = np.random.default_rng(2).integers(20, 900, 101)
    b, e = shard_by_residues(lens, world, rank)
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([b, e, int(got.sum())]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_share_db_and_split_queries(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0[2] == r1[2]                      # both ranks hold the same DB bytes
    assert r0[0] == 0 and r0[1] == r1[0] and r1[1] == 101   # contiguous, complete, disjoint
    lens = np.random.default_rng(2).integers(20, 900, 101)
    a, b = lens[r0[0]:r0[1]].sum(), lens[r1[0]:r1[1]].sum()
    assert abs(int(a) - int(b)) <= 900         # balanced by residues


def test_shard_edge_cases():
    assert shard_by_residues([], 4, 2) == (0, 0)
    assert shard_by_residues([10], 1, 0) == (0, 1)
    parts = [shard_by_residues([5, 5, 5], 8, r) for r in range(8)]
    assert parts[0][0] == 0 and parts[-1][1] == 3 and all(parts[i][1] == parts[i + 1][0] for i in range(7))


def test_ranks_then_lanes_cover_every_query_once():
    """bench.py / the host module split twice with the same rule: queries over the ranks, a rank's queries over its lanes (contexts
    on one GPU, fsb200_create_lane); every query lands in exactly one (rank, lane)"""
    lens = np.random.default_rng(3).integers(20, 900, 1000)
    for world in (1, 2, 8):
        for lanes in (1, 2, 3):
            seen = np.zeros(len(lens), np.int32)
            for rank in range(world):
                qb, qe = shard_by_residues(lens, world, rank)
                for k in range(lanes):
                    b, e = shard_by_residues(lens[qb:qe], lanes, k)
                    seen[qb + b:qb + e] += 1
            assert (seen == 1).all(), (world, lanes)
