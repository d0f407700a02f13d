"""CPU tests of the oracle (oracle/vga_oracle.c): it must reproduce the reference's own known-answer
tests and the golden fixtures generated from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import GOLDEN, golden
from oracle import pyoracle as po


def grid_of(fx, maxdist=-1.0):
    return po.Grid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                   fx["state"], fx["line_off"], fx["lines"], maxdist)


def attr(fx, name):
    cols = [str(c) for c in fx["columns"]]
    return fx[f"attr_{cols.index(name)}"]


# ---- salaTest/testsparksieve.cpp:21-83 -----------------------------------------------------------

def test_sieve_one_block_garbage():
    g = po.sieve_kat(1, 1, 4, [[0.5, 0.2, 0.5, 0.7]])
    assert len(g) == 1 and g[0][0] == 0 and g[0][1] == pytest.approx(0.625)


def test_sieve_shift_start_and_end():
    g = po.sieve_kat(1, 1, 4, [[0.5, 0.2, 0.5, 0.7], [0.5, 0.1, 1.1, 0.9]])
    assert len(g) == 1 and g[0][0] == pytest.approx(0.55555555555) and g[0][1] == pytest.approx(0.625)


def test_sieve_delete_gap():
    assert len(po.sieve_kat(1, 1, 4, [[1.1, 0.2, 0.5, 0.7]])) == 0


def test_sieve_add_gap():
    g = po.sieve_kat(1, 1, 4, [[0.5, 0.2, 0.5, 0.1], [0.5, 0.3, 0.5, 0.7]])
    assert len(g) == 2
    assert g[0][0] == 0 and g[0][1] == pytest.approx(0.55555555555)
    assert g[1][0] == pytest.approx(0.625) and g[1][1] == pytest.approx(0.71428571)


# ---- salaTest/testpointmap.cpp:313-453 (2x2 connections) -------------------------------------------

def test_box2x2_connections_text():
    fx = golden("box2x2")
    og = po.OracleGraph(grid_of(fx))
    refs = og.cell_refs()
    assert list(refs) == [65537, 65538, 131073, 131074]
    rp, ref, b = og.iter_rows()
    rows = [list(ref[int(rp[i]):int(rp[i + 1])]) for i in range(4)]
    # expected "connections [" blocks of PointMap::outputConnections
    assert rows == [[131073, 131074, 65538], [131074, 65537, 131073], [131074, 65538, 65537], [65538, 65537, 131073]]


# ---- golden fixtures from the reference --------------------------------------------------------------

@pytest.mark.parametrize("name", GOLDEN)
def test_makegraph_matches_reference(name):
    fx = golden(name)
    og = po.OracleGraph(grid_of(fx))
    rp, ref, b = og.iter_rows()
    assert np.array_equal(rp, fx["rowptr"])
    assert np.array_equal(ref, fx["ref"])
    assert np.array_equal(b, fx["bin"])
    a = og.node_attrs()
    assert np.array_equal(a["bin_count"], fx["bin_count"])
    assert np.array_equal(a["far"], fx["bin_dist"])
    assert np.array_equal(a["gridconn"], fx["gridconn"])
    assert np.array_equal(a["connectivity"], attr(fx, "Connectivity"))
    assert np.array_equal(a["first_moment"], attr(fx, "Point First Moment"))
    assert np.array_equal(a["second_moment"], attr(fx, "Point Second Moment"))


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("radius", [-1, 3])
def test_global_matches_reference(name, radius):
    fx = golden(name)
    og = po.OracleGraph(grid_of(fx))
    tn, td, dist, nl = og.global_ints(radius)
    f = po.global_formulas(tn, td, dist, nl)
    sfx = "" if radius == -1 else f" R{radius}"
    for k, v in f.items():
        assert np.array_equal(v, attr(fx, k + sfx)), k


@pytest.mark.parametrize("name", GOLDEN)
def test_local_matches_reference(name):
    fx = golden(name)
    og = po.OracleGraph(grid_of(fx))
    f = po.local_formulas(*og.local_ints())
    for k, v in f.items():
        assert np.array_equal(v, attr(fx, k)), k


def test_graph_from_edges_equals_built():
    fx = golden("oblique20")
    g = grid_of(fx)
    og = po.OracleGraph(g)
    og2 = po.OracleGraph(g, edges=(fx["rowptr"], fx["ref"]))
    a = og.global_ints(-1)
    b = og2.global_ints(-1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_maxdist_limits_targets():
    fx = golden("oblique20")
    og = po.OracleGraph(grid_of(fx, maxdist=4.0))
    a = og.node_attrs()
    full = po.OracleGraph(grid_of(fx)).node_attrs()
    assert (a["connectivity"] <= full["connectivity"]).all()
    assert a["far"].max() <= 4.0 + 1e-6


def test_step_depth_matches_reference():
    """oracle step depth == the reference's "Visual Step Depth" column (golden)."""
    sd = golden("stepdepth")
    keys = sorted(k for k in sd.files if k.endswith("__src"))
    assert keys
    cache = {}
    for k in keys:
        name, i, _ = k.split("__")
        if name not in cache:
            cache[name] = po.OracleGraph(grid_of(golden(name)))
        d = cache[name].step_depth(sd[k])
        assert np.array_equal(d.astype(np.float32), sd[f"{name}__{i}__depth"]), k


# ---- the CSR form of the oracle used by the full-size GPU parity tests -------------------------------------------

def ordinal_csr(flat, og):
    """The oracle's iterated adjacency as CSR of vertex ordinals: N cells in x-major order, then all unfilled cells
    in x-major order (the library's numbering, include/vga_b200.h)."""
    rp, ref, _ = og.iter_rows()
    st = flat.state.reshape(flat.cols, flat.rows)
    gx, gy = np.nonzero((st & 2) == 0)
    allrefs = np.concatenate([og.cell_refs(), ((gx.astype(np.int64) << 16) + gy).astype(np.int32)])
    key = allrefs.astype(np.uint32).astype(np.int64)
    order = np.argsort(key, kind="stable")
    col = order[np.searchsorted(key[order], ref.astype(np.uint32).astype(np.int64))].astype(np.uint32)
    return rp, col, allrefs


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:48:48:2", "room:24:24:1+holes"])
def test_csr_oracle_equals_graph_oracle(name):
    from depthmapx_b200 import capi, plans
    flat = capi.prepare(plans.by_name(name.split("+")[0]))
    if name.endswith("+holes"):  # unfilled cells inside diagonal runs -> ghost vertices
        st = flat.state.copy()
        for d in (3, 5):
            st[(2 + d) * flat.rows + (2 + d)] &= ~np.uint16(2)
        flat.state = st
    og = po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines))
    n = og.n
    rp, col, allrefs = ordinal_csr(flat, og)
    if name.endswith("+holes"):
        assert (col >= n).any()
    src = np.random.RandomState(1).choice(n, min(n, 40), replace=False)
    for radius in (-1, 2):
        tn, td, dist = po.global_csr(n, rp, col, src, radius)
        for i, s in enumerate(src):
            otn, otd, odist, _ = og.global_ints(radius, (int(s), int(s) + 1), maxl=64)
            assert otn[0] == tn[i] and otd[0] == td[i] and np.array_equal(odist[0], dist[i])
    # packed form (col << 6 | flags), as the library stores it
    tn2, td2, dist2 = po.global_csr(n, rp, (col << 6) | np.uint32(37), src, -1, shift=6)
    tn1, td1, dist1 = po.global_csr(n, rp, col, src, -1)
    assert np.array_equal(tn1, tn2) and np.array_equal(td1, td2) and np.array_equal(dist1, dist2)
    cl, kk, tot, ctl = po.local_csr(n, len(allrefs), rp, col, allrefs, src)
    for i, s in enumerate(src):
        o = og.local_ints((int(s), int(s) + 1))
        assert o[0][0] == cl[i] and o[1][0] == kk[i] and o[2][0] == tot[i] and o[3][0] == ctl[i]


# ---- metric / angular VGA: the oracle's vgao_metric / vgao_angular against the reference's own columns
# (tests/golden/metric_angular.npz, written from libdmxref.so by tests/golden/make_golden_metric.py)

@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:40:40:1", "oblique:24:24:5:0.7", "urban:60:60:4"])
def test_metric_angular_oracle_matches_reference_columns(name):
    import os
    from conftest import ROOT
    from depthmapx_b200 import capi, plans
    z = np.load(os.path.join(ROOT, "tests", "golden", "metric_angular.npz"), allow_pickle=False)
    flat = capi.prepare(plans.by_name(name))
    og = po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines))
    # all sources of the small plans, three slices of the larger ones (one C call per slice)
    slices = [(0, og.n)] if og.n <= 1200 else [(0, 60), (og.n // 2, og.n // 2 + 60), (og.n - 60, og.n)]
    for tag, radius in (("n", -1.0), ("r", float(z[name + "/metric_radius"]))):
        for lo, hi in slices:
            got = og.metric(flat.spacing, radius, (lo, hi))
            for col, a in zip(("angle", "path", "line", "count"), got):
                assert np.array_equal(z[f"{name}/metric_{tag}/{col}"][lo:hi].view(np.int32), a.view(np.int32)), (tag, col, lo)
    for tag, radius in (("n", -1.0), ("r", float(z[name + "/angular_radius"]))):
        for lo, hi in slices:
            got = og.angular(radius, (lo, hi))
            for col, a in zip(("mean", "total", "count"), got):
                assert np.array_equal(z[f"{name}/angular_{tag}/{col}"][lo:hi].view(np.int32), a.view(np.int32)), (tag, col, lo)
