"""Executed-work counters of vga_global under the SIMT emulation (tests/emu/): how many adjacency entries, atomics,
run records and pyramid nodes a BFS schedule really touches on a plan, early exit and pruning included.

    python tests/emu/build_emu.py
    python tools/emu_work_counts.py office:64:64:1 "" "bfs_pull=1,bfs_push=1,bfs_coarse=0"

Each further argument is a comma-separated option set ("" = defaults); the first one is the reference the others must
reproduce bit for bit.  Memory operations ~ 2 per adjacency entry or node-list entry (id + word), 1 per atomic / run
record / pyramid node reached through a run, 15 per pyramid build or down group, 3 per k_update word.  Test infrastructure: it loads the emulation build, never the product
library."""
import os, sys, json, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from depthmapx_b200 import capi, plans
capi.LIBDIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'emu', '_build')
name = sys.argv[1]
flat = capi.prepare(plans.by_name(name))
L = capi.abi(); L.simt_counters_dump.restype = C.c_char_p; L.simt_counters_dump.argtypes = [C.c_int]
base = None
for opts in sys.argv[2:]:
    ctx = capi.Context(0)
    for kv in opts.split(","):
        if kv: k, v = kv.split("="); ctx.set_option(k, int(v))
    g = ctx.build(flat)
    L.simt_counters_dump(1)
    res = g.global_ints(-1)
    c = json.loads(L.simt_counters_dump(1).decode())
    if base is None: base = res
    same = all(np.array_equal(a, b) for a, b in zip(base[:3], res[:3]))
    mem = c.get("push_entries",0)*2 + c.get("push_atomics",0) + c.get("pull_entries",0)*2 + c.get("ppush_runs",0) + c.get("ppush_nodes",0) + c.get("ppush_atomics",0) \
          + c.get("pyr_down_groups",0)*15 + c.get("pyr_build_groups",0)*15 + c.get("ppull_runs",0) + c.get("ppull_nodes",0) + c.get("update_words",0)*3 \
          + c.get("npush_nodes",0)*2 + c.get("npush_atomics",0) + c.get("npull_nodes",0)*2
    print(f"{name} N={g.n} E={g.entries} [{opts or 'default'}] same={same} memory-ops~{mem:.3e}  {c}")
    ctx.close()
