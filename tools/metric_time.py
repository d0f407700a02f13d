"""Development aid: time vga_metric / vga_angular on a plan (sampled sources), check a few sources against the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from depthmapx_b200 import plans, capi

name = sys.argv[1]
nsrc = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
check = int(sys.argv[3]) if len(sys.argv) > 3 else 0
flat = capi.prepare(plans.by_name(name))
ctx = capi.Context(0)
for kv in sys.argv[4:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
g = ctx.build(flat)
ba = capi.blocked_adjacent(flat)
print(f"{name}: N={g.n} E={g.entries} expanding cells {int(ba.sum())} ({ba.mean():.3f})", flush=True)
src = np.sort(np.random.RandomState(3).choice(g.n, min(g.n, nsrc), replace=False))
for what in ("metric", "angular"):
    for rep in range(2):
        t0 = time.time()
        out = g.metric(ba, flat.spacing, -1.0, src) if what == "metric" else g.angular(ba, -1.0, src)
        dt = time.time() - t0
        tm = ctx.timing()
        print(f"{what} rep{rep}: {len(src)} sources wall {dt*1e3:.1f} ms kernels {tm['kernel_ms']:.1f} ms -> {len(src)/dt:.0f} sources/s, "
              f"full map {g.n/len(src)*dt:.2f} s; unsafe angle evaluations {out[-1]}; mean count {out[-2].mean():.1f}", flush=True)
    if check:
        from oracle import pyoracle as po
        og = po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines))
        bad = 0
        t0 = time.time()
        for i in range(0, len(src), max(1, len(src) // check)):
            s = int(src[i])
            o = og.metric(flat.spacing, -1.0, (s, s + 1)) if what == "metric" else og.angular(-1.0, (s, s + 1))
            for a, b in zip(o, out[:-1]):
                if a.view(np.int32)[0] != b.view(np.int32)[i]:
                    bad += 1
                    print("  MISMATCH", what, s, a[0], b[i])
        print(f"  oracle check of {what}: {bad} mismatching values; oracle {(time.time()-t0)/max(1,check)*1e3:.1f} ms per source (1 core)", flush=True)
