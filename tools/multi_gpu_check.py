"""Multi-GPU parity check (run under torchrun on N GPUs): sharded makegraph + all-gather + partitioned
BFS must equal the single-GPU result bit for bit."""
import os, sys, time
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
lo, hi = multi.partition(n, world)[rank]
g = ctx.build(flat, (lo, hi))
rp_ptr, adj_ptr, ne = g.device_rows()
rp_local = multi.wrap(rp_ptr, (hi - lo + 1) * 8, torch.int64, dev)
adj_local = multi.wrap(adj_ptr, ne * 4, torch.int32, dev)[:ne]
rp_full, adj_full, total = multi.allgather_rows(rp_local, adj_local, dist, world)
torch.cuda.synchronize(dev)
full = ctx.graph_from_device_rows(n, g.ghosts, rp_full.data_ptr(), adj_full.data_ptr(), total)
full.set_cell_refs(g.cell_refs())
tn, td, hist, used = full.global_ints(-1, (lo, hi))
L = 32
pack = np.zeros((hi - lo, L + 2), np.int64)
pack[:, 0] = tn; pack[:, 1] = td; pack[:, 2:2 + min(L, hist.shape[1])] = hist[:, :L]
counts = [e - s for s, e in multi.partition(n, world)]
res = multi.gather_results(torch.from_numpy(pack).to(dev), counts, dist, rank, world)
cl, kk, tot, ctl = full.local_ints((lo, min(hi, lo + 64)))
ok = True
if rank == 0:
    res = res.cpu().numpy()
    single = ctx.build(flat)
    stn, std_, shist, sused = single.global_ints(-1)
    ok &= bool(np.array_equal(res[:, 0], stn) and np.array_equal(res[:, 1], std_))
    Lc = min(L, shist.shape[1])
    ok &= bool(np.array_equal(res[:, 2:2 + Lc], shist[:, :Lc]))
    srp, scol, sb, sacc = single.csr()
    frp, fcol, fb, facc = full.csr()
    ok &= bool(np.array_equal(srp, frp) and np.array_equal(scol, fcol) and np.array_equal(sb, fb) and np.array_equal(sacc, facc))
    scl, skk, stot, sctl = single.local_ints((lo, min(hi, lo + 64)))
    ok &= bool(np.array_equal(cl, scl) and np.array_equal(kk, skk) and np.array_equal(tot, stot) and np.array_equal(ctl, sctl))
    print(f"multi-gpu check world={world} plan={name} N={n} E={total}: {'OK' if ok else 'MISMATCH'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
