"""GPU: metric / angular VGA (csrc/metric.cu) through the C ABI against the oracle's restatement of
VGAMetric::run / VGAAngular::run (oracle/vga_oracle.c vgao_metric / vgao_angular, pinned bit-equal against the unmodified
reference by tests/test_oracle.py::test_metric_angular_*), and against the reference's own columns committed as goldens
(tests/golden/metric_angular.npz, made by tests/golden/make_golden_metric.py).

Bar: float32 BIT-equal for every column.  The node counts and both distance columns depend on IEEE operations only; the
angle sums also go through acos, whose last bits may differ between CUDA and glibc -- the library counts the evaluations
whose float rounding could be affected (`unsafe`), and the assertion on the angle columns is made when that count is 0
(it is on every plan here)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from depthmapx_b200 import capi, plans

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]

PLANS = ["oblique:30:30:7", "office:40:40:1", "oblique:24:24:5:0.7", "urban:60:60:4"]


def bits(a):
    return np.asarray(a, np.float32).view(np.int32)


def oracle_for(name):
    from oracle import pyoracle as po
    flat = capi.prepare(plans.by_name(name))
    og = po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines))
    return flat, og


@pytest.mark.parametrize("name", PLANS)
@pytest.mark.parametrize("radius_cells", [-1.0, 6.5])
def test_metric_vs_oracle(name, radius_cells):
    flat, og = oracle_for(name)
    radius = radius_cells if radius_cells < 0 else radius_cells * flat.spacing
    c = capi.Context(0)
    g = c.build(flat)
    ba = capi.blocked_adjacent(flat)
    rng = np.random.RandomState(5)
    src = np.sort(rng.choice(g.n, min(g.n, 96), replace=False))
    angle, path, line, count, unsafe = g.metric(ba, flat.spacing, radius, src)
    for i, s in enumerate(src):
        oa, op, ol, oc = og.metric(flat.spacing, radius, (int(s), int(s) + 1))
        assert bits(oc)[0] == bits(count)[i] and bits(op)[0] == bits(path)[i] and bits(ol)[0] == bits(line)[i], (name, s)
        if unsafe == 0:
            assert bits(oa)[0] == bits(angle)[i], (name, s)
        else:
            assert abs(float(oa[0]) - float(angle[i])) <= 1e-5 * max(1.0, abs(float(oa[0])))
    c.close()


@pytest.mark.parametrize("name", PLANS)
@pytest.mark.parametrize("radius", [-1.0, 1.5])
def test_angular_vs_oracle(name, radius):
    flat, og = oracle_for(name)
    c = capi.Context(0)
    g = c.build(flat)
    ba = capi.blocked_adjacent(flat)
    rng = np.random.RandomState(6)
    src = np.sort(rng.choice(g.n, min(g.n, 96), replace=False))
    mean, total, count, unsafe = g.angular(ba, radius, src)
    for i, s in enumerate(src):
        om, ot, oc = og.angular(radius, (int(s), int(s) + 1))
        if unsafe == 0:
            assert bits(oc)[0] == bits(count)[i] and bits(ot)[0] == bits(total)[i] and bits(om)[0] == bits(mean)[i], (name, s)
        else:
            assert abs(float(ot[0]) - float(total[i])) <= 1e-5 * max(1.0, abs(float(ot[0])))
    c.close()


def test_whole_map_and_source_order():
    """sources = None runs every cell in x-major order; an explicit permuted list gives the same values in list order; one
    slot (metric_slots = 4) and many slots agree."""
    flat, og = oracle_for("oblique:20:20:3")
    ba = capi.blocked_adjacent(flat)
    c = capi.Context(0)
    g = c.build(flat)
    full = g.metric(ba, flat.spacing)
    perm = np.random.RandomState(1).permutation(g.n)
    part = g.metric(ba, flat.spacing, -1.0, perm)
    for a, b in zip(full[:4], part[:4]):
        assert np.array_equal(bits(a)[perm], bits(b))
    c.set_option("metric_slots", 4)
    few = g.metric(ba, flat.spacing)
    for a, b in zip(full[:4], few[:4]):
        assert np.array_equal(bits(a), bits(b))
    oa, op, ol, oc = og.metric(flat.spacing)
    assert np.array_equal(bits(oc), bits(full[3])) and np.array_equal(bits(op), bits(full[1])) and np.array_equal(bits(ol), bits(full[2]))
    if full[4] == 0:
        assert np.array_equal(bits(oa), bits(full[0]))
    fa = g.angular(ba)
    om, ot, ocn = og.angular()
    if fa[3] == 0:
        assert np.array_equal(bits(om), bits(fa[0])) and np.array_equal(bits(ot), bits(fa[1])) and np.array_equal(bits(ocn), bits(fa[2]))
    c.close()


@pytest.mark.parametrize("name", ["oblique:20:20:3", "office:24:24:2"])
def test_merge_links_vs_oracle(name):
    """Merge links (Point::m_merge): the partner of a finalised cell is expanded from the same key and finalised without
    being counted (vgametric.cpp:96-104, vgaangular.cpp:91-99); oracle pinned against the reference with merges by
    tests/test_oracle_vs_reference.py::test_metric_angular_with_merge_links_against_the_reference."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    g = c.build(flat)
    ba = capi.blocked_adjacent(flat)
    rng = np.random.RandomState(8)
    cells = rng.choice(g.n, 8, replace=False)
    partner = np.full(g.n, -1, np.int32)
    for a, b in zip(cells[0::2], cells[1::2]):
        partner[a], partner[b] = b, a
    for radius in (-1.0, 5.0 * flat.spacing):
        got = g.metric(ba, flat.spacing, radius, None, partner)
        want = og.metric(flat.spacing, radius, partner=partner)
        for i, (a, b) in enumerate(zip(want, got[:4])):
            if i == 0 and got[4] != 0:
                continue
            assert np.array_equal(bits(a), bits(b)), (name, radius, i)
    for radius in (-1.0, 1.0):
        got = g.angular(ba, radius, None, partner)
        want = og.angular(radius, partner=partner)
        if got[3] == 0:
            for i, (a, b) in enumerate(zip(want, got[:3])):
                assert np.array_equal(bits(a), bits(b)), (name, radius, i)
    bad = partner.copy()
    bad[cells[0]] = cells[2]  # asymmetric
    with pytest.raises(capi.VgaError):
        g.metric(ba, flat.spacing, -1.0, None, bad)
    c.close()


def test_errors():
    flat, _ = oracle_for("oblique:20:20:3")
    c = capi.Context(0)
    g = c.build(flat)
    ba = capi.blocked_adjacent(flat)
    with pytest.raises(capi.VgaError):
        g.metric(ba, flat.spacing, -1.0, np.array([g.n], np.int64))
    rp, col, b, acc = g.csr()
    bare = c.graph_from_csr(g.n, g.ghosts, rp, col, b)  # no coordinates
    with pytest.raises(capi.VgaError):
        bare.metric(ba, flat.spacing)
    bare.set_cell_refs(g.cell_refs())
    got = bare.metric(ba, flat.spacing, -1.0, np.arange(8))
    want = g.metric(ba, flat.spacing, -1.0, np.arange(8))
    for a, b2 in zip(got[:4], want[:4]):
        assert np.array_equal(bits(a), bits(b2))
    c.close()


def test_reference_golden_columns():
    """The reference's own columns (unmodified VGAMetric / VGAAngular through libdmxref.so), committed as fixtures."""
    path = os.path.join(ROOT, "tests", "golden", "metric_angular.npz")
    z = np.load(path, allow_pickle=False)
    for name in [str(x) for x in z["plans"]]:
        flat = capi.prepare(plans.by_name(name))
        ba = capi.blocked_adjacent(flat)
        c = capi.Context(0)
        g = c.build(flat)
        for tag, radius in (("n", -1.0), ("r", float(z[name + "/metric_radius"]))):
            got = g.metric(ba, flat.spacing, radius)
            for col, a in zip(("angle", "path", "line", "count"), got[:4]):
                if col == "angle" and got[4] != 0:
                    continue
                assert np.array_equal(bits(z[f"{name}/metric_{tag}/{col}"]), bits(a)), (name, tag, col)
        for tag, radius in (("n", -1.0), ("r", float(z[name + "/angular_radius"]))):
            got = g.angular(ba, radius)
            if got[3] == 0:
                for col, a in zip(("mean", "total", "count"), got[:3]):
                    assert np.array_equal(bits(z[f"{name}/angular_{tag}/{col}"]), bits(a)), (name, tag, col)
        c.close()
