"""Generate tests/golden/metric_angular.npz from the UNMODIFIED reference (oracle/_ref/libdmxref.so): the columns
VGAMetric::run (radius n and a finite radius) and VGAAngular::run (radius n and a finite radius) write for a few
by-name plans (depthmapx_b200/plans.by_name), so that the GPU box -- which has no /root/reference -- can compare with the
reference itself.  Run in the build container only:

    python tests/golden/make_golden_metric.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402
from depthmapx_b200 import plans  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
PLANS = ["oblique:30:30:7", "office:40:40:1", "oblique:24:24:5:0.7", "urban:60:60:4"]
METRIC = ["Metric Mean Shortest-Path Angle", "Metric Mean Shortest-Path Distance", "Metric Mean Straight-Line Distance",
          "Metric Node Count"]
ANGULAR = ["Angular Mean Depth", "Angular Total Depth", "Angular Node Count"]


def main():
    out = {"plans": np.array(PLANS)}
    for name in PLANS:
        p = plans.by_name(name)
        rm = po.RefMap(p.walls, p.spacing)
        for s in p.seeds:
            assert rm.fill(*s)
        rm.makegraph()
        mr = 7.25 * p.spacing
        ar = 1.5
        out[name + "/metric_radius"] = np.float64(mr)
        out[name + "/angular_radius"] = np.float64(ar)
        for tag, r in (("n", -1.0), ("r", mr)):
            assert rm.vga_metric(r) >= 0
            sfx = "" if r == -1.0 else " R%.2f" % r
            for col, key in zip(METRIC, ("angle", "path", "line", "count")):
                out[f"{name}/metric_{tag}/{key}"] = rm.attr(col + sfx)
        for tag, r in (("n", -1.0), ("r", ar)):
            assert rm.vga_angular(r) >= 0
            sfx = "" if r == -1.0 else " R%.2f" % r
            for col, key in zip(ANGULAR, ("mean", "total", "count")):
                out[f"{name}/angular_{tag}/{key}"] = rm.attr(col + sfx)
        print(name, "cells", rm.n)
    np.savez_compressed(os.path.join(HERE, "metric_angular.npz"), **out)


if __name__ == "__main__":
    main()
