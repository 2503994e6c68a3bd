"""Multi-GPU plumbing of the search path: the target DB is replicated (one broadcast at start-up), queries are sharded.

Mirrors how the reference splits work: queries are independent units (Prefiltering.cpp:848 `omp for` over queries,
query-split mode Prefiltering.cpp:781-786 via DBReader::decomposeDomainByAminoAcid); there is no data-path collective.
torch.distributed is plumbing only: NCCL on the GPUs, gloo in the CPU tests.
"""
import numpy as np


def shard_by_residues(lengths, world, rank):
    """Contiguous [begin, end) range of queries for `rank`, balanced by residue count like decomposeDomainByAminoAcid."""
    lengths = np.asarray(lengths, dtype=np.int64)
    n = len(lengths)
    if world <= 1:
        return 0, n
    cum = np.concatenate([[0], np.cumsum(lengths)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        b = int(np.searchsorted(cum, target, side="left"))
        bounds.append(max(bounds[-1], min(n, b)))
    bounds.append(n)
    return bounds[rank], bounds[rank + 1]


def broadcast_bytes(dist, tensor_factory, payload, src=0):
    """Broadcast a 1-D uint8 payload (given on `src` as a tensor) to every rank. tensor_factory(n) allocates the receive
    buffer on the right device. Returns the tensor on every rank."""
    import torch
    rank = dist.get_rank()
    dev = payload.device if (rank == src and payload is not None) else tensor_factory(1).device
    n = torch.zeros(1, dtype=torch.int64, device=dev)
    if rank == src:
        n[0] = payload.numel()
    dist.broadcast(n, src)
    buf = payload if rank == src else tensor_factory(int(n.item()))
    dist.broadcast(buf, src)
    return buf
