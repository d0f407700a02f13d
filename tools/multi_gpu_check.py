"""Multi-GPU parity check (run under torchrun on N GPUs): sharded makegraph (work-balanced source ranges) + exchange of
the run-length rows + partitioned BFS / local must equal the single-GPU result bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from depthmapx_b200 import capi, plans, multi

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
name = sys.argv[1] if len(sys.argv) > 1 else "office:128:128:5"
flat = capi.prepare(plans.by_name(name))
ctx = capi.Context(lr)
n = flat.n_filled
parts = multi.partition_by_work(multi.estimate_source_work(flat.state, flat.cols, flat.rows), world)
lo, hi = parts[rank]
g = ctx.build(flat, (lo, hi))
full = multi.replicate_graph(ctx, g, n, hi - lo, dist, rank, world, dev)
order = full.batch_order()
a, b = multi.partition(n, world)[rank]
mine = order[a:b].astype(np.int64)
tn, td, hist, used = full.global_ints(-1, sources=mine)
L = 32
pack = np.zeros((len(mine), L + 3), np.int64)
pack[:, 0] = mine; pack[:, 1] = tn; pack[:, 2] = td; pack[:, 3:3 + min(L, hist.shape[1])] = hist[:, :L]
counts = [e - s for s, e in multi.partition(n, world)]
res = multi.gather_results(torch.from_numpy(pack).to(dev), counts, dist, rank, world)
cells = (lo, min(hi, lo + 64))
cl, kk, tot, ctl = full.local_ints(cells)
ok = True
single = ctx.build(flat)
scl, skk, stot, sctl = single.local_ints(cells)
ok &= bool(np.array_equal(cl, scl) and np.array_equal(kk, skk) and np.array_equal(tot, stot) and np.array_equal(ctl, sctl))
srp, scol, sb, sacc = single.csr()
mrp, mcol, mb, macc = g.csr()
ok &= bool(np.array_equal(mrp + srp[lo], srp[lo:hi + 1]) and np.array_equal(mcol, scol[int(srp[lo]):int(srp[hi])])
           and np.array_equal(mb, sb[int(srp[lo]):int(srp[hi])]))
if rank == 0:
    res = res.cpu().numpy()
    stn, std_, shist, sused = single.global_ints(-1)
    o = np.argsort(res[:, 0])
    ok &= bool(np.array_equal(res[o, 0], np.arange(n)) and np.array_equal(res[o, 1], stn) and np.array_equal(res[o, 2], std_))
    Lc = min(L, shist.shape[1])
    ok &= bool(np.array_equal(res[o, 3:3 + Lc], shist[:, :Lc]))
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
ok = bool(flag.item())
if rank == 0:
    print(f"multi-gpu check world={world} plan={name} N={n} parts={parts}: {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
