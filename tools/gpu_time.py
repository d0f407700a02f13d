"""Development aid: time the GPU stages on a plan (no oracle)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from depthmapx_b200 import plans, capi

name = sys.argv[1]
what = sys.argv[2] if len(sys.argv) > 2 else "build,global"
p = plans.by_name(name)
t0 = time.time(); f = capi.prepare(p); print(f"prep {time.time()-t0:.2f}s N={f.n_filled}", flush=True)
ctx = capi.Context(0)
for kv in sys.argv[3:]:
    k, v = kv.split("=")
    ctx.set_option(k, int(v))
dg = ctx.upload(f)
for rep in range(2):
    t0 = time.time(); g = ctx.build(dg); dt = time.time() - t0
    print(f"build rep{rep}: wall {dt*1e3:.1f} ms N={g.n} E={g.entries} {ctx.timing()}", flush=True)
    if rep == 0: g.free()
if "global" in what:
    for rad in [int(x) for x in os.environ.get("VGA_TIME_RADII", "-1,3").split(",")]:
        for rep in range(int(os.environ.get("VGA_TIME_REPS", "2"))):
            lim = int(os.environ.get("VGA_TIME_SRC", "0"))
            src = None if lim <= 0 else (g.n // 2 - lim // 2, g.n // 2 - lim // 2 + lim)
            t0 = time.time(); tn, td, dist, used = g.global_ints(rad, src); dt = time.time() - t0
            print(f"global r={rad} rep{rep}: wall {dt*1e3:.1f} ms levels={used} meandepth={(td/np.maximum(tn-1,1)).mean():.3f} {ctx.timing()}", flush=True)
if "local" in what:
    t0 = time.time(); cl, kk, tot, ctl = g.local_ints(); dt = time.time() - t0
    print(f"local: wall {dt*1e3:.1f} ms {ctx.timing()} sum2paths={int((kk.astype(np.int64)).sum())}", flush=True)
