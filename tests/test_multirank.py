"""World-size-2 gloo test (CPU) of the multi-GPU host logic: source partition, padded all-gather of row
shards, result gather.  Shard rows come from the oracle (stand-in for each rank's GPU shard)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from depthmapx_b200 import multi
    from oracle import pyoracle as po
    fx = golden("office24")
    grid = po.Grid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                   fx["state"], fx["line_off"], fx["lines"])
    n = grid.n_filled
    lo, hi = multi.partition(n, world)[rank]
    og = po.OracleGraph(grid, src_range=(lo, hi))
    rp, ref, b = og.iter_rows()
    rp_local = torch.from_numpy((rp[lo:hi + 1] - rp[lo]).astype(np.int64))
    adj_local = torch.from_numpy(ref[int(rp[lo]):int(rp[hi])].astype(np.int32))
    rp_full, adj_full, total = multi.allgather_rows(rp_local, adj_local, dist, world)
    # every rank now holds the full adjacency: run its share of the BFS on it
    full = po.OracleGraph(grid, edges=(rp_full.numpy().astype(np.uint64), adj_full.numpy()))
    tn, td, hist, nl = full.global_ints(-1, (lo, hi), maxl=16)
    pack = np.concatenate([tn[:, None].astype(np.int64), td[:, None], hist.astype(np.int64)], axis=1)
    counts = [e - s for s, e in multi.partition(n, world)]
    res = multi.gather_results(torch.from_numpy(pack), counts, dist, rank, world)
    if rank == 0:
        np.savez(os.path.join(out_dir, "res.npz"), rp=rp_full.numpy(), adj=adj_full.numpy(), res=res.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_gather(tmp_path):
    from oracle import pyoracle as po
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    out = np.load(os.path.join(str(tmp_path), "res.npz"))
    fx = golden("office24")
    assert np.array_equal(out["rp"].astype(np.uint64), fx["rowptr"])
    assert np.array_equal(out["adj"], fx["ref"])
    grid = po.Grid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                   fx["state"], fx["line_off"], fx["lines"])
    og = po.OracleGraph(grid)
    tn, td, hist, nl = og.global_ints(-1, maxl=16)
    assert np.array_equal(out["res"][:, 0], tn)
    assert np.array_equal(out["res"][:, 1], td)
    assert np.array_equal(out["res"][:, 2:], hist)


def test_partition_covers_everything():
    from depthmapx_b200 import multi
    for n in (0, 1, 7, 64, 65537):
        for w in (1, 2, 3, 8):
            p = multi.partition(n, w)
            assert p[0][0] == 0 and p[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(p[:-1], p[1:]))
            assert max(e - s for s, e in p) - min(e - s for s, e in p) <= 1
