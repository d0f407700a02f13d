"""GPU: visual step depth (vga_step_depth) against the reference's golden column
(VGAVisualGlobalDepth::run) and the oracle, plus consistency with the all-sources BFS histogram."""
import numpy as np
import pytest

from conftest import golden
from depthmapx_b200 import capi, plans

pytestmark = pytest.mark.gpu


def test_step_depth_vs_reference_golden():
    ctx = capi.Context(0)
    sd = golden("stepdepth")
    for k in sorted(x for x in sd.files if x.endswith("__src")):
        name, i, _ = k.split("__")
        fx = golden(name)
        flat = capi.FlatGrid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                             fx["state"], fx["line_off"], fx["lines"])
        g = ctx.build(flat)
        d = g.step_depth(sd[k])
        assert np.array_equal(d.astype(np.float32), sd[f"{name}__{i}__depth"]), k
    ctx.close()


def test_step_depth_vs_oracle_office():
    from oracle import pyoracle as po
    flat = capi.prepare(plans.by_name("office:64:64:1"))
    ctx = capi.Context(0)
    g = ctx.build(flat)
    og = po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines))
    for src in ([0], [10, 2000, 4095], list(range(0, 4096, 97))):
        assert np.array_equal(g.step_depth(src), og.step_depth(src))
    # consistency with the all-sources BFS: depth from {s} has the histogram vga_global reports for s
    tn, td, dist, used = g.global_ints(-1, (123, 124))
    d = g.step_depth([123])
    assert np.array_equal(np.bincount(d[d >= 0], minlength=dist.shape[1])[:dist.shape[1]], dist[0])
    ctx.close()
