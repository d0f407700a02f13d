"""CPU: the oracle (oracle/vga_oracle.c) and the host pre-steps (depthmapx_b200/host) against the UNMODIFIED
reference compiled into oracle/_ref/libdmxref.so, on seeded random plans (oblique walls, spacings 0.5..2.0).
This is what "parity pinned" rests on besides the committed golden fixtures.  Skipped when the reference
library has not been built (make -C oracle ref needs /root/reference)."""
import random

import numpy as np
import pytest

from depthmapx_b200 import capi, plans
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libdmxref.so not built")


@pytest.mark.parametrize("seed", range(16))
def test_random_plan_bit_equal_to_reference(seed):
    rng = random.Random(1000 + seed)
    w, h = rng.randrange(8, 22), rng.randrange(8, 22)
    sp = rng.choice([1.0, 0.7, 1.3, 0.5, 2.0])
    p = plans.oblique(w, h, seed, n_axis=rng.randrange(2, 8), n_oblique=rng.randrange(0, 5), spacing=sp)
    rm = po.RefMap(p.walls, p.spacing)
    if not rm.fill(*p.seeds[0]):
        pytest.skip("seed cell not fillable in this plan")
    g = rm.grid()
    # host pre-steps (setGrid, blockLines, makePoints) == reference
    hm = capi.HostMap(p.walls, p.spacing)
    assert hm.fill(*p.seeds[0])
    f = hm.flat()
    assert np.array_equal(g.state, f.state) and np.array_equal(g.line_off, f.line_off) and np.array_equal(g.lines, f.lines)
    # oracle makegraph == reference sparkGraph2
    og = po.OracleGraph(g)
    rm.makegraph()
    rp, ref, b = rm.edges()
    orp, oref, ob = og.iter_rows()
    assert np.array_equal(rp, orp) and np.array_equal(ref, oref) and np.array_equal(b, ob)
    a = og.node_attrs()
    cnt, dist, gc = rm.bins()
    assert np.array_equal(cnt, a["bin_count"]) and np.array_equal(dist, a["far"]) and np.array_equal(gc, a["gridconn"])
    assert np.array_equal(rm.attr("Connectivity"), a["connectivity"])
    assert np.array_equal(rm.attr("Point First Moment"), a["first_moment"])
    assert np.array_equal(rm.attr("Point Second Moment"), a["second_moment"])
    # oracle global / local == reference VGAVisualGlobal / VGAVisualLocal (all columns, float32 bit-equal)
    radius = rng.choice([-1, 2, 3])
    rm.vga_global(float(radius))
    tn, td, d, nl = og.global_ints(radius)
    sfx = "" if radius == -1 else f" R{radius}"
    for k, v in po.global_formulas(tn, td, d, nl).items():
        assert np.array_equal(rm.attr(k + sfx), v), k
    if rm.n <= 300:  # the reference's local analysis is O(k^3)
        rm.vga_local()
        for k, v in po.local_formulas(*og.local_ints()).items():
            assert np.array_equal(rm.attr(k), v), k
    # oracle metric / angular == reference VGAMetric / VGAAngular (row f4), radius n and a finite radius, float32 bit-equal
    if rm.n > 800:
        return  # both searches are O(n^2 log n) per map; the larger plans are covered by tests/golden/metric_angular.npz
    for mr in (-1.0, rng.choice([3.0, 5.5, 8.0]) * sp):
        assert rm.vga_metric(mr) >= 0
        sfx = "" if mr == -1.0 else " R%.2f" % mr
        names = ["Metric Mean Shortest-Path Angle", "Metric Mean Shortest-Path Distance", "Metric Mean Straight-Line Distance",
                 "Metric Node Count"]
        for k, v in zip(names, og.metric(p.spacing, mr)):
            assert np.array_equal(rm.attr(k + sfx).view(np.int32), v.view(np.int32)), (k, mr)
    for ar in (-1.0, rng.choice([0.5, 1.0, 2.0])):
        assert rm.vga_angular(ar) >= 0
        sfx = "" if ar == -1.0 else " R%.2f" % ar
        for k, v in zip(["Angular Mean Depth", "Angular Total Depth", "Angular Node Count"], og.angular(ar)):
            assert np.array_equal(rm.attr(k + sfx).view(np.int32), v.view(np.int32)), (k, ar)


@pytest.mark.ref
@pytest.mark.parametrize("name", ["oblique:20:20:11", "office:24:24:2"])
def test_context_filled_semantics_oracle_vs_reference(name):
    """Semi-fill (fill_type 1: FILLED | CONTEXTFILLED): cells that are not "even" are skipped as sources by global and
    local, counted but not expanded under a radius, and not expanded beyond level 0 by step depth -- the oracle's
    restatement against the reference's VGAVisualGlobal / VGAVisualLocal / VGAVisualGlobalDepth on the same map."""
    if not po.have_ref():
        pytest.skip("compiled reference not present")
    from depthmapx_b200 import plans
    plan = plans.by_name(name)
    r = po.RefMap(plan.walls, plan.spacing)
    for s in plan.seeds:
        assert r.fill(*s, fill_type=1)
    grid = r.grid()
    assert ((grid.state & 8) != 0).sum() == r.n
    og = po.OracleGraph(grid)
    r.makegraph()
    r.vga_local()
    r.vga_global(-1.0)
    r.vga_global(3.0)
    for radius, suffix in ((-1, ""), (3, " R3")):
        tn, td, dist, nl = og.global_ints(radius)
        skipped = tn == -1
        assert 0 < skipped.sum() < len(tn)
        for k, v in po.global_formulas(tn, td, dist, nl).items():
            assert np.array_equal(np.where(skipped, -1.0, v).astype(np.float32), r.attr(k + suffix)), k + suffix
    cl, k, tot, ctl = og.local_ints()
    for n_, v in po.local_formulas(cl, np.where(k < 0, 0, k), tot, ctl).items():
        assert np.array_equal(np.where(k < 0, -1.0, v).astype(np.float32), r.attr(n_)), n_
    for src in ([5, 40], [0], [17, 18, 19]):
        assert np.array_equal(og.step_depth(src).astype(np.float32), r.step_depth(src))


@pytest.mark.parametrize("seed", range(6))
def test_metric_angular_with_merge_links_against_the_reference(seed):
    """Row f4 with merge links (vgametric.cpp:96-104, vgaangular.cpp:91-99): three random pairs of merged cells, radius n
    and a finite radius, every column float32 bit-equal to the reference's."""
    rng = random.Random(3000 + seed)
    w, h = rng.randrange(10, 24), rng.randrange(10, 24)
    sp = rng.choice([1.0, 0.7, 1.3])
    p = plans.oblique(w, h, seed, n_axis=rng.randrange(2, 8), n_oblique=rng.randrange(0, 5), spacing=sp)
    rm = po.RefMap(p.walls, p.spacing)
    if not rm.fill(*p.seeds[0]):
        pytest.skip("seed cell not fillable in this plan")
    g = rm.grid()
    og = po.OracleGraph(g)
    rm.makegraph()
    refs = og.cell_refs()
    if og.n < 10:
        pytest.skip("too few cells")
    partner = np.full(og.n, -1, np.int32)
    cells = rng.sample(range(og.n), 6)
    for a, b in zip(cells[0::2], cells[1::2]):
        ra, rb = int(refs[a]), int(refs[b])
        assert rm.merge(g.bl_x + (ra >> 16) * sp, g.bl_y + (ra & 0xffff) * sp, g.bl_x + (rb >> 16) * sp, g.bl_y + (rb & 0xffff) * sp)
        partner[a], partner[b] = b, a
    for mr in (-1.0, 6.0 * sp):
        assert rm.vga_metric(mr) >= 0
        sfx = "" if mr == -1.0 else " R%.2f" % mr
        names = ["Metric Mean Shortest-Path Angle", "Metric Mean Shortest-Path Distance", "Metric Mean Straight-Line Distance",
                 "Metric Node Count"]
        for k, v in zip(names, og.metric(p.spacing, mr, partner=partner)):
            assert np.array_equal(rm.attr(k + sfx).view(np.int32), v.view(np.int32)), (k, mr)
    for ar in (-1.0, 1.0):
        assert rm.vga_angular(ar) >= 0
        sfx = "" if ar == -1.0 else " R%.2f" % ar
        for k, v in zip(["Angular Mean Depth", "Angular Total Depth", "Angular Node Count"], og.angular(ar, partner=partner)):
            assert np.array_equal(rm.attr(k + sfx).view(np.int32), v.view(np.int32)), (k, ar)
