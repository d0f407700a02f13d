"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libdmxref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    python tests/golden/make_golden.py

Each fixture holds the inputs (walls, spacing, seeds) and what the reference produced for them:
grid geometry, Point::m_state, per-cell clipped wall lines, the iterated adjacency of every Node
(bin order), bin counts / far distances / grid connections, and every attribute column written by
sparkGraph2, VGAVisualLocal::run and VGAVisualGlobal::run (radius n and radius 3).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from depthmapx_b200 import plans  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def box2x2():
    # salaTest/testpointmap.cpp:313-345: 1.5 x 1.5 box, spacing 0.5 -> 2x2 filled cells
    r = 1.5
    p = plans.Plan("box2x2", 2, 2, [(0, 0, 0, r), (0, r, r, r), (r, r, r, 0), (r, 0, 0, 0)], [], 0.5)
    return p


def fixture(plan, seeds=None):
    rm = po.RefMap(plan.walls, plan.spacing)
    if seeds is None:
        seeds = plan.seeds
    for s in seeds:
        assert rm.fill(*s)
    g = rm.grid()
    out = dict(walls=np.array(plan.walls, np.float64), spacing=plan.spacing, seeds=np.array(seeds, np.float64),
               cols=rm.cols, rows=rm.rows, bl_x=rm.bl_x, bl_y=rm.bl_y, state=g.state, line_off=g.line_off,
               lines=g.lines)
    rm.makegraph()
    rp, ref, b = rm.edges()
    cnt, dist, gc = rm.bins()
    out.update(rowptr=rp, ref=ref, bin=b, bin_count=cnt, bin_dist=dist, gridconn=gc)
    rm.vga_local()
    rm.vga_global(-1.0)
    rm.vga_global(3.0)
    cols = rm.columns()
    out["columns"] = np.array(cols)
    for i, c in enumerate(cols):
        out[f"attr_{i}"] = rm.attr(c)
    return out


def main():
    cases = {
        "box2x2": (box2x2(), None),
        "oblique20": (plans.oblique(20, 20, 11, n_axis=6, n_oblique=3), None),
        "oblique16s07": (plans.oblique(16, 16, 5, n_axis=5, n_oblique=2, spacing=0.7), None),
        "office24": (plans.office(24, 24, 2, room_size=6, corridor=2, door=2), None),
    }
    for name, (plan, seeds) in cases.items():
        if name == "box2x2":
            # midpoint seed as in the reference test
            rm = po.RefMap(plan.walls, plan.spacing)
            gx = rm.bl_x - plan.spacing / 2
            gy = rm.bl_y - plan.spacing / 2
            import math
            seeds = [(gx + plan.spacing * (math.floor(rm.cols * 0.5) + 0.5), gy + plan.spacing * (math.floor(rm.rows * 0.5) + 0.5))]
        fx = fixture(plan, seeds)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **fx)
        print(name, "N", int(((fx["state"] & 2) != 0).sum()), "E", len(fx["ref"]), os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
