"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: work-balanced source partition, the exchange of run-length
row shards (every rank's shard broadcast into its slice of the final buffers) and the result gather.  Shard rows come
from the oracle (stand-in for each rank's GPU shard)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden


def runs_of_rows(rp, col, n):
    """Run-length form of sorted ordinal rows: (runptr int64 [rows+1], runs int64 [R] = first | length << 32), ghost
    columns (>= n) excluded -- what vga_graph_device_runs hands out on the GPU."""
    runptr = [0]
    runs = []
    for v in range(len(rp) - 1):
        c = np.sort(col[int(rp[v]):int(rp[v + 1])].astype(np.int64))
        c = c[c < n]
        if len(c):
            brk = np.flatnonzero(np.diff(c) != 1) + 1
            starts = np.concatenate([[0], brk])
            ends = np.concatenate([brk, [len(c)]])
            for s, e in zip(starts, ends):
                runs.append(int(c[s]) | (int(e - s) << 32))
        runptr.append(len(runs))
    return np.array(runptr, np.int64), np.array(runs, np.int64)


def expand_runs(runptr, runs):
    rp = [0]
    col = []
    for v in range(len(runptr) - 1):
        for r in runs[int(runptr[v]):int(runptr[v + 1])]:
            first, length = int(r) & 0xffffffff, int(r) >> 32
            col.extend(range(first, first + length))
        rp.append(len(col))
    return np.array(rp, np.uint64), np.array(col, np.uint32)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from depthmapx_b200 import multi
    from oracle import pyoracle as po
    from test_oracle import ordinal_csr
    fx = golden("office24")
    grid = po.Grid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                   fx["state"], fx["line_off"], fx["lines"])
    n = grid.n_filled
    w = multi.estimate_source_work(grid.state, grid.cols, grid.rows)
    lo, hi = multi.partition_by_work(w, world)[rank]
    og = po.OracleGraph(grid, src_range=(lo, hi))

    class F:  # ordinal_csr only needs these
        state, cols, rows = grid.state, grid.cols, grid.rows
    rp, col, allrefs = ordinal_csr(F, og)
    rp_local = (rp[lo:hi + 1] - rp[lo]).astype(np.int64)
    col_local = col[int(rp[lo]):int(rp[hi])]
    runptr, runs = runs_of_rows(rp_local, col_local, n)
    deg = np.diff(rp_local).astype(np.int32)
    sizes = multi.exchange_sizes(hi - lo, len(runs), dist, world, torch.device("cpu"))
    rp_full, runs_full, deg_full, total = multi.allgather_runs(torch.from_numpy(runptr), torch.from_numpy(runs),
                                                               torch.from_numpy(deg), sizes, dist, rank, world)
    # every rank now holds the full run-length graph: its share of the BFS sources runs on it
    frp, fcol = expand_runs(rp_full.numpy(), runs_full.numpy()[:total])
    share = np.arange(n)[rank::world]
    tn, td, hist = po.global_csr(n, frp, fcol, share, -1, maxl=16)
    pack = np.concatenate([share[:, None].astype(np.int64), tn[:, None].astype(np.int64), td[:, None], hist.astype(np.int64)], axis=1)
    counts = [len(np.arange(n)[r::world]) for r in range(world)]
    res = multi.gather_results(torch.from_numpy(pack), counts, dist, rank, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "res.npz"), rp=frp, col=fcol, deg=deg_full.numpy(), res=res.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_run_exchange_and_result_gather(tmp_path):
    from oracle import pyoracle as po
    from test_oracle import ordinal_csr
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    out = np.load(os.path.join(str(tmp_path), "res.npz"))
    fx = golden("office24")
    grid = po.Grid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                   fx["state"], fx["line_off"], fx["lines"])
    og = po.OracleGraph(grid)
    n = og.n

    class F:
        state, cols, rows = grid.state, grid.cols, grid.rows
    rp, col, _ = ordinal_csr(F, og)
    # the exchanged runs expand to the sorted rows without the ghost columns
    want_rp, want_col = [0], []
    for v in range(n):
        c = np.sort(col[int(rp[v]):int(rp[v + 1])])
        want_col.extend(c[c < n].tolist())
        want_rp.append(len(want_col))
    assert np.array_equal(out["rp"], np.array(want_rp, np.uint64))
    assert np.array_equal(out["col"], np.array(want_col, np.uint32))
    assert np.array_equal(out["deg"], np.diff(rp).astype(np.int32))
    tn, td, hist, nl = og.global_ints(-1, maxl=16)
    res = out["res"]
    order = np.argsort(res[:, 0])
    assert np.array_equal(res[order, 0], np.arange(n))
    assert np.array_equal(res[order, 1], tn)
    assert np.array_equal(res[order, 2], td)
    assert np.array_equal(res[order, 3:], hist)


def test_partitions():
    from depthmapx_b200 import multi
    for n in (0, 1, 7, 64, 65537):
        for w in (1, 2, 3, 8):
            p = multi.partition(n, w)
            assert p[0][0] == 0 and p[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(p[:-1], p[1:]))
            assert max(e - s for s, e in p) - min(e - s for s, e in p) <= 1
    rng = np.random.RandomState(0)
    for n in (0, 1, 5, 1000):
        wts = rng.randint(1, 1000, n)
        for w in (1, 2, 3, 8):
            p = multi.partition_by_work(wts, w)
            assert len(p) == w and p[0][0] == 0 and p[-1][1] == n
            assert all(a[1] == b[0] and a[0] <= a[1] for a, b in zip(p[:-1], p[1:]))
            if n == 1000:
                sums = [wts[a:b].sum() for a, b in p]
                assert max(sums) - min(sums) <= 2 * wts.max()


def test_work_estimate_follows_open_area():
    """A cell in a long corridor gets more estimated work than a cell in a small closed room."""
    from depthmapx_b200 import multi
    cols, rows = 12, 8
    st = np.zeros((cols, rows), np.uint16)
    st[1:11, 1] = 2          # a corridor of 10 cells
    st[2:4, 4:6] = 2         # a 2 x 2 room
    w = multi.estimate_source_work(st.reshape(-1), cols, rows)
    filled = np.argwhere((st & 2) != 0)  # x-major order
    assert len(w) == len(filled)
    corridor = w[[i for i, (x, y) in enumerate(filled) if y == 1]]
    room = w[[i for i, (x, y) in enumerate(filled) if y >= 4]]
    assert corridor.min() == 10 and room.max() == 4
