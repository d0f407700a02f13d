"""GPU: every schedule of the all-sources BFS (csrc/bfs.cu) against the oracle -- top-down only, bottom-up only and the
direction-optimising hybrid, every word width, radius limits, explicit source lists -- plus the two row-ordering paths
of makegraph (own bitmap-rank kernel = default, CUB segmented sort = fallback for rows that exceed shared memory).

The in-rows come from the O(runs) transposition of the run-length out-rows (k_trans_events / k_trans_pair): the
bottom-up-only schedule (bfs_mode = 1) reads nothing else after level 0, so it pins that transposition."""
import numpy as np
import pytest

from depthmapx_b200 import capi, plans

pytestmark = [pytest.mark.timeout(900, method="thread"), pytest.mark.gpu]

_CACHE = {}


def oracle_for(name):
    """(flat grid, oracle graph) of a plan, built once per test session (the oracle of urban:120 takes seconds)."""
    if name not in _CACHE:
        from oracle import pyoracle as po
        flat = capi.prepare(plans.by_name(name.split("+")[0]))
        if name.endswith("+holes"):  # unfilled cells inside diagonal runs -> ghost columns
            st = flat.state.copy()
            for d in (3, 5):
                st[(2 + d) * flat.rows + (2 + d)] &= ~np.uint16(2)
            flat.state = st
        _CACHE[name] = (flat, po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state,
                                                     flat.line_off, flat.lines)))
    return _CACHE[name]


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4", "room:24:24:1+holes"])
@pytest.mark.parametrize("radius", [-1, 2])
@pytest.mark.parametrize("mode,words,order", [(0, 1, 0), (1, 1, 1), (2, 1, 2), (0, 4, 2), (1, 2, 2), (2, 4, 2), (2, 8, 2),
                                              (1, 8, 0), (2, 0, 2)])
def test_schedules_vs_oracle(name, radius, mode, words, order):
    """One lane per node (bfs_coop = 0) for every width; words = 0 = the automatic choice with the default kernels."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    for k, v in (("bfs_mode", mode), ("bfs_words", words), ("bfs_order", order), ("bfs_coop", 1 if words == 0 else 0)):
        c.set_option(k, v)
    g = c.build(flat)
    tn, td, dist, used = g.global_ints(radius)
    rng = np.random.RandomState(9)
    for s in rng.choice(g.n, min(g.n, 48), replace=False):
        otn, otd, odist, onl = og.global_ints(radius, (int(s), int(s) + 1), maxl=64)
        L = dist.shape[1]
        assert otn[0] == tn[s] and otd[0] == td[s]
        assert np.array_equal(odist[0, :L], dist[s]) and not odist[0, L:].any()
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "room:24:24:1+holes"])
@pytest.mark.parametrize("radius", [-1, 2])
@pytest.mark.parametrize("mode,words,unroll", [(0, 4, 2), (1, 4, 2), (2, 4, 2), (0, 8, 2), (1, 8, 4), (2, 8, 2), (1, 4, 4), (2, 4, 4)])
def test_lane_cooperative_kernels_vs_oracle(name, radius, mode, words, unroll):
    """bfs_coop = 1 (default): W/2 lanes share a pyramid node, 16 bytes each (k_push_nodes_coop / k_pull_nodes_coop with 2 or
    4 node loads per lane between early-exit checks)."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    for k, v in (("bfs_coop", 1), ("bfs_mode", mode), ("bfs_words", words), ("bfs_pull_unroll", unroll)):
        c.set_option(k, v)
    g = c.build(flat)
    tn, td, dist, used = g.global_ints(radius)
    otn, otd, odist, onl = og.global_ints(radius, maxl=max(dist.shape[1], 8))
    L = dist.shape[1]
    assert np.array_equal(tn, otn) and np.array_equal(td, otd)
    assert np.array_equal(dist, odist[:, :L]) and not odist[:, L:].any()
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1"])
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_whole_map_every_source(name, mode):
    """Every source of the map, every level count (not a sample), for the three directions."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    c.set_option("bfs_mode", mode)
    g = c.build(flat)
    for radius in (-1, 3):
        tn, td, dist, used = g.global_ints(radius)
        otn, otd, odist, onl = og.global_ints(radius, maxl=dist.shape[1])
        assert np.array_equal(tn, otn) and np.array_equal(td, otd) and np.array_equal(dist, odist)
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "urban:120:120:4"])
def test_explicit_source_lists(name):
    """vga_global_sources: scattered, unsorted and repeated sources; outputs in list order."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    g = c.build(flat)
    rng = np.random.RandomState(4)
    for count, radius in ((1, -1), (7, -1), (200, 2), (min(g.n, 700), -1)):
        src = rng.choice(g.n, count, replace=False)
        if count == 7:
            src[3] = src[0]  # the same source twice
        tn, td, dist, used = g.global_ints(radius, sources=src)
        for i in range(0, count, max(1, count // 40)):
            s = int(src[i])
            otn, otd, odist, onl = og.global_ints(radius, (s, s + 1), maxl=64)
            L = dist.shape[1]
            assert otn[0] == tn[i] and otd[0] == td[i]
            assert np.array_equal(odist[0, :L], dist[i]) and not odist[0, L:].any()
    assert len(g.global_ints(-1, sources=np.zeros(0, np.int64))[0]) == 0
    order = g.batch_order()
    assert np.array_equal(np.sort(order), np.arange(g.n))
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4", "room:24:24:1+holes"])
@pytest.mark.parametrize("mode,words", [(0, 4), (1, 4), (2, 0), (1, 1)])
def test_hybrid_x_or_y_major_lists_vs_oracle(name, mode, words):
    """bfs_hybrid = 1 (default): every row is walked as the shorter of its x-major and y-major pyramid-node lists (second
    pyramid over the y-major vertex order).  Never longer than x-major only, and the same integers as the oracle in every
    direction (top-down only pins the y-major out-lists, bottom-up only the y-major in-lists)."""
    flat, og = oracle_for(name)
    sizes = {}
    res = {}
    for hyb in (0, 1):
        c = capi.Context(0)
        for k, v in (("bfs_hybrid", 2 * hyb), ("bfs_mode", mode), ("bfs_words", words)):
            c.set_option(k, v)
        g = c.build(flat)
        res[hyb] = g.global_ints(-1)
        sizes[hyb] = g.list_sizes()
        c.close()
    assert sizes[1]["out_nodes"] <= sizes[0]["out_nodes"] and sizes[1]["out_runs"] == sizes[0]["out_runs"]
    if mode != 0:
        assert sizes[1]["in_nodes"] <= sizes[0]["in_nodes"]
    for a, b in zip(res[0][:3], res[1][:3]):
        assert np.array_equal(a, b)
    tn, td, dist, _ = res[1]
    rng = np.random.RandomState(11)
    for s in rng.choice(len(tn), min(len(tn), 32), replace=False):
        otn, otd, odist, onl = og.global_ints(-1, (int(s), int(s) + 1), maxl=64)
        L = dist.shape[1]
        assert otn[0] == tn[s] and otd[0] == td[s]
        assert np.array_equal(odist[0, :L], dist[s]) and not odist[0, L:].any()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4", "room:24:24:1+holes"])
@pytest.mark.parametrize("mode,words,unroll,hyb", [(0, 1, 4, 0), (0, 2, 2, 2), (2, 4, 4, 2), (0, 4, 1, 2), (0, 8, 2, 0), (2, 0, 4, 1),
                                                   (1, 4, 2, 2), (1, 1, 1, 0), (2, 8, 1, 2)])
def test_delta_push_vs_full_push_and_oracle(name, mode, words, unroll, hyb):
    """bfs_delta = 1 (default): out-rows stored as [new nodes | nodes whose cells row u-1 holds as well]; a vertex pushes
    its whole vector only to the new nodes and F[u] & ~F[u-1] to the others (k_push_delta).  Same integers as every vertex
    pushing its whole row (bfs_delta = 0, k_push_nodes[_coop]) and as the oracle; top-down only pins it at every level."""
    flat, og = oracle_for(name)
    for radius in (-1, 2):
        res = {}
        for delta in (0, 1):
            c = capi.Context(0)
            for k, v in (("bfs_delta", delta), ("bfs_mode", mode), ("bfs_words", words),
                         ("bfs_push_unroll", unroll), ("bfs_delta_unroll", unroll), ("bfs_hybrid", hyb)):
                c.set_option(k, v)
            g = c.build(flat)
            res[delta] = g.global_ints(radius)
            c.close()
        for a, b in zip(res[0][:3], res[1][:3]):
            assert np.array_equal(a, b)
        tn, td, dist, _ = res[1]
        rng = np.random.RandomState(13)
        for s in rng.choice(len(tn), min(len(tn), 24), replace=False):
            otn, otd, odist, onl = og.global_ints(radius, (int(s), int(s) + 1), maxl=64)
            L = dist.shape[1]
            assert otn[0] == tn[s] and otd[0] == td[s]
            assert np.array_equal(odist[0, :L], dist[s]) and not odist[0, L:].any()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4", "room:24:24:1+holes"])
def test_row_ordering_kernels_vs_oracle(name):
    """build_sort = 1 (default): rows ordered by bitmap rank in shared memory (k_rank_sort); build_sort = 0: CUB segmented
    radix sort.  The sorted adjacency incl. bins, accepted flags and ghost columns must equal the oracle's and each other."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    c.set_option("build_sort", 1)
    g = c.build(flat)
    rp, col, b, acc = g.csr()
    orp, oref, ob = og.iter_rows()
    assert np.array_equal(rp, orp)
    if not (col >= g.n).any():  # with ghosts the column order is "filled cells, then ghosts": compared with the other build below
        refs = g.cell_refs()
        rowid = np.repeat(np.arange(len(orp) - 1), np.diff(orp).astype(np.int64))
        order = np.lexsort((oref.astype(np.int64), rowid))
        assert np.array_equal(refs[col], oref[order]) and np.array_equal(b, ob[order])
    else:
        assert name.endswith("+holes")
    d = capi.Context(0)
    d.set_option("build_sort", 0)
    rp2, col2, b2, acc2 = d.build(flat).csr()
    assert np.array_equal(col, col2) and np.array_equal(b, b2) and np.array_equal(acc, acc2)
    c.close()
    d.close()


def test_run_length_rows_round_trip():
    """vga_graph_device_runs -> vga_graph_from_device_runs (what a multi-GPU run exchanges): the adopted BFS-only graph
    gives the integers of the graph it came from."""
    flat, og = oracle_for("office:64:64:1")
    c = capi.Context(0)
    g = c.build(flat)
    ref = g.global_ints(-1)
    rp, runs, nr = g.device_runs()
    assert nr > 0
    h = c.graph_from_device_runs(g.n, g.ghosts, rp, runs, nr)
    h.set_cell_refs(g.cell_refs())
    got = h.global_ints(-1)
    assert ref[3] == got[3]
    for x, y in zip(ref[:3], got[:3]):
        assert np.array_equal(x, y)
    c.close()
