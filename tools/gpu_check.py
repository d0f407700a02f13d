"""Ad-hoc GPU parity check (development aid): CUDA path vs oracle on a list of plans."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import pyoracle as po
from depthmapx_b200 import plans, capi


def check(name, radius=-1, local_n=64, mode=None):
    p = plans.by_name(name)
    t0 = time.time()
    f = capi.prepare(p)
    t_prep = time.time() - t0
    og_grid = po.Grid(f.cols, f.rows, f.spacing, f.bl_x, f.bl_y, f.state, f.line_off, f.lines, f.maxdist)
    ctx = capi.Context(0)
    if mode is not None:
        ctx.set_option("bfs_mode", mode)
    t0 = time.time()
    g = ctx.build(f)
    t_build = time.time() - t0
    tb = ctx.timing()
    print(f"[{name}] N={g.n} ghosts={g.ghosts} E={g.entries} prep={t_prep:.2f}s build={t_build:.3f}s timing={tb}", flush=True)
    t0 = time.time()
    og = po.OracleGraph(og_grid)
    print(f"  oracle makegraph {time.time()-t0:.2f}s", flush=True)
    rp, col, b, acc = g.csr()
    orp, oref, ob = og.iter_rows()
    refs = g.cell_refs()
    ok_rp = np.array_equal(rp, orp)
    # oracle rows are in bin order; sort each row by (x,y) == ordinal order
    ok_adj = False
    if ok_rp:
        oref_sorted = np.empty_like(oref)
        ob_sorted = np.empty_like(ob)
        key = oref.astype(np.int64)
        rowid = np.repeat(np.arange(len(orp) - 1), np.diff(orp).astype(np.int64))
        order = np.lexsort((key, rowid))
        oref_sorted = oref[order]; ob_sorted = ob[order]
        ok_adj = np.array_equal(refs[col], oref_sorted) and np.array_equal(b, ob_sorted)
    st = g.node_stats()
    a = og.node_attrs()
    ok_stats = (np.array_equal(st["connectivity"].astype(np.float32), a["connectivity"]) and
                np.array_equal(st["sum_d"].astype(np.float32), a["first_moment"]) and
                np.array_equal(st["sum_d2"].astype(np.float32), a["second_moment"]) and
                np.array_equal(st["far"], a["far"]) and
                np.array_equal(st["bin_count"].astype(np.uint16), a["bin_count"]) and
                np.array_equal(st["gridconn"], a["gridconn"]))
    print(f"  makegraph parity: rowptr={ok_rp} adj={ok_adj} stats={ok_stats}", flush=True)
    # global
    t0 = time.time()
    tn, td, dist, used = g.global_ints(radius)
    t_g = time.time() - t0
    tg = ctx.timing()
    print(f"  global radius={radius}: {t_g:.3f}s levels={used} timing={tg}", flush=True)
    nsamp = min(g.n, 256)
    rng = np.random.RandomState(1)
    samp = np.sort(rng.choice(g.n, nsamp, replace=False))
    ok_g = True
    t0 = time.time()
    for s in samp:
        otn, otd, odist, onl = og.global_ints(radius, (int(s), int(s) + 1), maxl=max(used, 64))
        L = dist.shape[1]
        if otn[0] != tn[s] or otd[0] != td[s] or not np.array_equal(odist[0, :L], dist[s]) or odist[0, L:].any():
            ok_g = False
            print("   MISMATCH source", s, otn[0], tn[s], otd[0], td[s])
            break
    print(f"  global parity ({nsamp} sampled sources, oracle {time.time()-t0:.1f}s): {ok_g}", flush=True)
    # local
    ln = min(g.n, local_n)
    lo = (g.n // 2) - ln // 2
    t0 = time.time()
    cl, kk, tot, ctl = g.local_ints((lo, lo + ln))
    t_l = time.time() - t0
    ocl, okk, otot, octl = og.local_ints((lo, lo + ln))
    ok_l = np.array_equal(cl, ocl) and np.array_equal(kk, okk) and np.array_equal(tot, otot) and np.array_equal(ctl, octl)
    print(f"  local parity ({ln} cells, gpu {t_l:.3f}s): {ok_l}", flush=True)
    if os.environ.get("VGA_FULL_LOCAL"):
        t0 = time.time()
        g.local_ints()
        print(f"  local full: {time.time()-t0:.3f}s timing={ctx.timing()}", flush=True)
    return ok_rp and ok_adj and ok_stats and ok_g and ok_l


if __name__ == "__main__":
    names = sys.argv[1:] or ["oblique:30:30:7", "oblique:30:30:8:0.7", "office:64:64:1", "C1"]
    allok = True
    for n in names:
        rad = -1
        if "@" in n:
            n, r = n.split("@")
            rad = int(r)
        allok &= check(n, rad)
    print("ALL OK" if allok else "FAILURES")
    sys.exit(0 if allok else 1)
