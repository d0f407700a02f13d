"""Multi-GPU plumbing (one process per GPU, torch.distributed): work-balanced source partition for the sharded
makegraph, the single exchange that replicates the graph -- as RUN-LENGTH rows (8 bytes per run: 1.3 GB at 10^6 cells
instead of 22 GB of entries), every rank's shard broadcast straight into its final place, no padding, no concatenation
-- and the final result gather.  Device-agnostic torch code so that the logic is testable with gloo on CPU; on the GPU
box the tensors wrap the library's device pointers (zero copy) and the backend is NCCL over NVLink."""
from __future__ import annotations

import numpy as np
import torch


def partition(n: int, world: int):
    """Contiguous ranges [(begin, end)] per rank (sizes differ by at most 1)."""
    return [((n * r) // world, (n * (r + 1)) // world) for r in range(world)]


def estimate_source_work(state: np.ndarray, cols: int, rows: int) -> np.ndarray:
    """Estimated construction work per filled cell (x-major order), from the open area around it:
    (free run to the left + right + 1) * (free run down + up + 1) of filled cells -- the visible set of a cell of a
    street or a room grows with the lengths of the free axis-parallel runs through it.  O(cells) numpy passes."""
    f = ((state.reshape(cols, rows) & 2) != 0)

    def run_before(a):  # along axis 0: number of consecutive filled cells strictly before each cell
        out = np.zeros(a.shape, np.int32)
        cur = np.zeros(a.shape[1], np.int32)
        for i in range(a.shape[0]):
            out[i] = cur
            cur = np.where(a[i], cur + 1, 0)
        return out
    left = run_before(f)
    right = run_before(f[::-1])[::-1]
    down = run_before(f.T).T
    up = run_before(f.T[::-1])[::-1].T
    w = (left + right + 1).astype(np.int64) * (down + up + 1).astype(np.int64)
    return w[f]  # boolean indexing of a (cols, rows) array walks x-major


def partition_by_work(weights: np.ndarray, world: int):
    """Contiguous ranges whose summed weights are as equal as a prefix-sum cut allows."""
    n = len(weights)
    if n == 0:
        return [(0, 0)] * world
    c = np.cumsum(weights.astype(np.float64))
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(c, c[-1] * r / world)))
    cuts.append(n)
    cuts = np.maximum.accumulate(np.array(cuts))
    return [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:])]


class DevicePtr:
    """Zero-copy view of a raw CUDA pointer as a torch tensor via __cuda_array_interface__."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (max(nbytes, 1),), "typestr": "|u1", "data": (ptr, False), "version": 2}


def wrap(ptr: int, nbytes: int, dtype, device):
    return torch.as_tensor(DevicePtr(ptr, nbytes), device=device).view(dtype)


def exchange_sizes(rows: int, nruns: int, dist, world: int, device):
    """[(rows, runs)] of every rank (one tiny all-gather)."""
    sizes = torch.tensor([rows, nruns], dtype=torch.int64, device=device)
    all_sizes = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(all_sizes, sizes)
    return [(int(a), int(b)) for a, b in all_sizes.cpu().view(world, 2).tolist()]


def allgather_runs(runptr_local: torch.Tensor, runs_local: torch.Tensor, deg_local: torch.Tensor, sizes, dist, rank: int,
                   world: int, out=None):
    """Replicates run-length rows.  runptr_local int64 [rows+1] starting at 0, runs_local int64 [nruns] (one (first,
    length) pair of u32 per element), deg_local int32 [rows]; sizes from exchange_sizes.  `out` = (runptr [N+1] int64,
    runs [R] int64, deg [N] int32) tensors to fill (e.g. views of the library's final allocation), allocated here when
    None.  Every rank's shard is broadcast straight into its slice; offsets are rebased in place."""
    device = runptr_local.device
    n = sum(s[0] for s in sizes)
    total = sum(s[1] for s in sizes)
    if out is None:
        out = (torch.empty(n + 1, dtype=torch.int64, device=device), torch.empty(max(total, 1), dtype=torch.int64, device=device),
               torch.empty(max(n, 1), dtype=torch.int32, device=device))
    rp_full, runs_full, deg_full = out
    row0 = run0 = 0
    for r, (rows, nr) in enumerate(sizes):
        rp_slice = rp_full[row0:row0 + rows]
        runs_slice = runs_full[run0:run0 + nr]
        deg_slice = deg_full[row0:row0 + rows]
        if r == rank:
            rp_slice.copy_(runptr_local[:rows])
            if nr:
                runs_slice.copy_(runs_local[:nr])
            if rows:
                deg_slice.copy_(deg_local[:rows])
        if world > 1:
            if rows:
                dist.broadcast(rp_slice, src=r)
                dist.broadcast(deg_slice, src=r)
            if nr:
                dist.broadcast(runs_slice, src=r)
        if rows:
            rp_slice += run0
        row0 += rows
        run0 += nr
    rp_full[n:n + 1] = total
    return rp_full, runs_full, deg_full, total


def replicate_graph(ctx, g, n: int, rows: int, dist, rank: int, world: int, device, stats=None):
    """The run-length rows of every rank's shard `g` (rows [lo, lo + rows) of the n cells), broadcast straight into a
    BFS / local graph allocated by the library (vga_graph_runs_alloc): one pass, no padding, no staging copy.  Returns the
    replicated capi.Graph (no entries: it serves vga_global, vga_global_sources and vga_local)."""
    rp_ptr, runs_ptr, nr = g.device_runs()
    deg_ptr = g.device_degrees()
    sizes = exchange_sizes(rows, nr, dist, world, device)
    total_runs = sum(s[1] for s in sizes)
    full, f_rp, f_runs, f_deg = ctx.graph_runs_alloc(n, g.ghosts, total_runs)
    out = (wrap(f_rp, (n + 1) * 8, torch.int64, device), wrap(f_runs, max(total_runs, 1) * 8, torch.int64, device),
           wrap(f_deg, max(n, 1) * 4, torch.int32, device))
    allgather_runs(wrap(rp_ptr, (rows + 1) * 8, torch.int64, device), wrap(runs_ptr, max(nr, 1) * 8, torch.int64, device),
                   wrap(deg_ptr, max(rows, 1) * 4, torch.int32, device), sizes, dist, rank, world, out=out)
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    full.runs_commit()
    full.set_cell_refs(g.cell_refs())  # coordinates -> spatially coherent BFS batches
    if stats is not None:
        stats["exchange_bytes"] = int(total_runs) * 8 + (n + 1) * 8 + n * 4
    return full


def gather_results(mine: torch.Tensor, counts, dist, rank: int, world: int, dst: int = 0):
    """mine: int64 [rows_r, width] per-source integers of this rank; returns [sum(counts), width] on dst (else None)."""
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
