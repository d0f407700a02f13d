"""Multi-GPU plumbing (one process per GPU, torch.distributed): source partitioning, the single
all-gather that replicates the adjacency after a sharded makegraph, and the final result gather.
Device-agnostic torch code so that the logic is testable with gloo on CPU; on the GPU box the
tensors wrap the library's device pointers (zero copy) and the backend is NCCL over NVLink."""
from __future__ import annotations

import torch


def partition(n: int, world: int):
    """Contiguous x-major source ranges [(begin, end)] per rank (sizes differ by at most 1)."""
    return [((n * r) // world, (n * (r + 1)) // world) for r in range(world)]


class DevicePtr:
    """Zero-copy view of a raw CUDA pointer as a torch tensor via __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 2}


def wrap(ptr: int, nbytes: int, dtype, device):
    return torch.as_tensor(DevicePtr(ptr, nbytes), device=device).view(dtype)


def allgather_rows(rowptr_local: torch.Tensor, adj_local: torch.Tensor, dist, world: int):
    """rowptr_local: int64 [rows+1] starting at 0; adj_local: int32 [entries].  Returns the replicated
    (rowptr [N+1], adj [E]) in rank order.  One size exchange + two padded all-gathers."""
    device = rowptr_local.device
    rows = rowptr_local.numel() - 1
    ne = int(adj_local.numel())
    sizes = torch.tensor([rows, ne], dtype=torch.int64, device=device)
    all_sizes = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_sizes, sizes)
    all_sizes = all_sizes.cpu().view(world, 2)
    max_rows, max_ne = int(all_sizes[:, 0].max()), max(int(all_sizes[:, 1].max()), 1)
    rp_pad = torch.zeros(max_rows + 1, dtype=torch.int64, device=device)
    rp_pad[:rows + 1] = rowptr_local
    adj_pad = torch.zeros(max_ne, dtype=torch.int32, device=device)
    adj_pad[:ne] = adj_local[:ne]
    rp_all = torch.empty(world * (max_rows + 1), dtype=torch.int64, device=device)
    adj_all = torch.empty(world * max_ne, dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(rp_all, rp_pad)
    dist.all_gather_into_tensor(adj_all, adj_pad)
    rp_all = rp_all.view(world, max_rows + 1)
    adj_all = adj_all.view(world, max_ne)
    parts_rp, parts_adj, base = [], [], 0
    for r in range(world):
        rr, ee = int(all_sizes[r, 0]), int(all_sizes[r, 1])
        parts_rp.append(rp_all[r, :rr] + base)
        parts_adj.append(adj_all[r, :ee])
        base += ee
    rp_full = torch.cat(parts_rp + [torch.tensor([base], dtype=torch.int64, device=device)]).contiguous()
    adj_full = torch.cat(parts_adj).contiguous() if base > 0 else torch.zeros(1, dtype=torch.int32, device=device)
    return rp_full, adj_full, base


def gather_results(mine: torch.Tensor, counts, dist, rank: int, world: int, dst: int = 0):
    """mine: int64 [rows_r, width] per-source integers of this rank; returns [N, width] on dst (else None)."""
    mx = max(counts)
    width = mine.shape[1]
    padded = torch.zeros((mx, width), dtype=mine.dtype, device=mine.device)
    padded[:mine.shape[0]] = mine
    if rank == dst:
        bufs = [torch.empty_like(padded) for _ in range(world)]
        dist.gather(padded, bufs, dst=dst)
        return torch.cat([bufs[r][:counts[r]] for r in range(world)])
    dist.gather(padded, None, dst=dst)
    return None
