"""GPU: opt-in EXPERIMENTAL code paths prepared for the next round.  They are not defaults and
have not run on a GPU yet, so every test here is xfail(strict=False): green or red, they cannot mask or break
the parity suite; they exist so that the first GPU call of the next round validates them.  On CPU the same tests run
against the SIMT emulation of the CUDA sources (tests/test_emulated_kernels.py, with --runxfail) and pass."""
import numpy as np
import pytest

from depthmapx_b200 import capi, plans

_CACHE = {}


def oracle_for(name):
    """(flat grid, oracle graph) of a plan, built once per test session (the oracle of urban:120 takes seconds)."""
    if name not in _CACHE:
        from oracle import pyoracle as po
        flat = capi.prepare(plans.by_name(name))
        _CACHE[name] = (flat, po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state,
                                                     flat.line_off, flat.lines)))
    return _CACHE[name]


# a kernel that never returns would hang the GPU box: give up on the whole run instead (these modules run last)
pytestmark = [pytest.mark.timeout(900, method="thread"), pytest.mark.gpu, pytest.mark.xfail(strict=False, reason="experimental opt-in path: passes under SIMT emulation on CPU, first B200 run pending")]


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "room:40:40:5"])
def test_local_bit_sliced_counters_vs_oracle(name):
    """local_mode = 3 (bit-sliced per-source counters in k_lb_expand_sliced) must equal the oracle."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    c.set_option("local_mode", 3)
    g = c.build(flat)
    hi = min(g.n, 1500)
    for x, y in zip(g.local_ints((0, hi)), og.local_ints((0, hi))):
        assert np.array_equal(x, y)
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4"])
@pytest.mark.parametrize("radius", [-1, 2])
@pytest.mark.parametrize("mode,words,coarse", [(1, 1, 0), (1, 2, 1), (2, 1, 1), (2, 4, 1), (2, 0, 1)])
def test_pyramid_pull_vs_oracle(name, radius, mode, words, coarse):
    """bfs_pull = 1: bottom-up step as range-OR queries over an OR-pyramid of the frontier with run-length in-rows
    (k_pyr_build / k_pull_pyr, csrc/pyramid.cuh; index logic pinned on CPU by tests/test_pyramid_logic.py)."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    for k, v in (("bfs_pull", 1), ("bfs_mode", mode), ("bfs_words", words), ("bfs_coarse", coarse)):
        c.set_option(k, v)
    g = c.build(flat)
    tn, td, dist, used = g.global_ints(radius)
    rng = np.random.RandomState(3)
    for s in rng.choice(g.n, min(g.n, 40), replace=False):
        otn, otd, odist, onl = og.global_ints(radius, (int(s), int(s) + 1), maxl=64)
        L = dist.shape[1]
        assert otn[0] == tn[s] and otd[0] == td[s]
        assert np.array_equal(odist[0, :L], dist[s]) and not odist[0, L:].any()
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1"])
@pytest.mark.parametrize("mode,words", [(0, 1), (0, 2), (2, 4), (2, 0)])
def test_unrolled_push_vs_oracle(name, mode, words):
    """bfs_push_unroll = 4: four adjacency entries per lane and round in the top-down step."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    for k, v in (("bfs_push_unroll", 4), ("bfs_mode", mode), ("bfs_words", words)):
        c.set_option(k, v)
    g = c.build(flat)
    tn, td, dist, used = g.global_ints(-1)
    otn, otd, odist, onl = og.global_ints(-1, maxl=dist.shape[1])
    assert np.array_equal(tn, otn) and np.array_equal(td, otd) and np.array_equal(dist, odist)
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4"])
@pytest.mark.parametrize("radius", [-1, 2])
@pytest.mark.parametrize("push,pull,mode,words", [(2, 0, 0, 1), (2, 0, 2, 2), (2, 1, 2, 4), (1, 1, 2, 0), (1, 0, 2, 1)])
def test_pyramid_push_vs_oracle(name, radius, push, pull, mode, words):
    """bfs_push: top-down step as range-OR updates over the runs of the out-rows through a pyramid of `next`
    (k_push_pyr + k_pyr_down); 2 forces it for every top-down step after level 0, 1 lets the cost model choose."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    for k, v in (("bfs_push", push), ("bfs_pull", pull), ("bfs_mode", mode), ("bfs_words", words)):
        c.set_option(k, v)
    g = c.build(flat)
    tn, td, dist, used = g.global_ints(radius)
    rng = np.random.RandomState(5)
    for s in rng.choice(g.n, min(g.n, 40), replace=False):
        otn, otd, odist, onl = og.global_ints(radius, (int(s), int(s) + 1), maxl=64)
        L = dist.shape[1]
        assert otn[0] == tn[s] and otd[0] == td[s]
        assert np.array_equal(odist[0, :L], dist[s]) and not odist[0, L:].any()
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4", "room:24:24:1+holes"])
def test_rank_sort_vs_oracle(name):
    """build_sort = 1: rows ordered by bitmap rank in shared memory (k_rank_sort) instead of the segmented radix sort;
    the sorted adjacency incl. bins, accepted flags and ghost columns must equal the oracle's."""
    from oracle import pyoracle as po
    flat = capi.prepare(plans.by_name(name.split("+")[0]))
    if name.endswith("+holes"):  # unfilled cells inside diagonal runs -> ghost columns (numbered after the filled cells)
        st = flat.state.copy()
        for d in (3, 5):
            st[(2 + d) * flat.rows + (2 + d)] &= ~np.uint16(2)
        flat.state = st
    og = po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines))
    c = capi.Context(0)
    c.set_option("build_sort", 1)
    g = c.build(flat)
    rp, col, b, acc = g.csr()
    orp, oref, ob = og.iter_rows()
    assert np.array_equal(rp, orp)
    if not (col >= g.n).any():  # with ghosts the column order is "filled cells, then ghosts": compared with the default build below
        refs = g.cell_refs()
        rowid = np.repeat(np.arange(len(orp) - 1), np.diff(orp).astype(np.int64))
        order = np.lexsort((oref.astype(np.int64), rowid))
        assert np.array_equal(refs[col], oref[order]) and np.array_equal(b, ob[order])
    else:
        assert name.endswith("+holes")
    d = capi.Context(0)
    rp2, col2, b2, acc2 = d.build(flat).csr()
    assert np.array_equal(col, col2) and np.array_equal(b, b2) and np.array_equal(acc, acc2)
    c.close()
    d.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4"])
@pytest.mark.parametrize("radius", [-1, 2])
@pytest.mark.parametrize("push,pull,mode,words,coarse", [(1, 1, 2, 0, 0), (2, 0, 0, 1, 1), (0, 1, 1, 2, 1), (2, 1, 2, 4, 0)])
def test_pyramid_node_lists_vs_oracle(name, radius, push, pull, mode, words, coarse):
    """bfs_pyr_nodes = 1: rows as lists of pyramid node ids, walked by k_push_nodes / k_pull_nodes (the inner loops of the
    entry kernels) instead of per-run decompositions."""
    flat, og = oracle_for(name)
    c = capi.Context(0)
    for k, v in (("bfs_pyr_nodes", 1), ("bfs_push", push), ("bfs_pull", pull), ("bfs_mode", mode), ("bfs_words", words),
                 ("bfs_coarse", coarse)):
        c.set_option(k, v)
    g = c.build(flat)
    tn, td, dist, used = g.global_ints(radius)
    rng = np.random.RandomState(9)
    for s in rng.choice(g.n, min(g.n, 40), replace=False):
        otn, otd, odist, onl = og.global_ints(radius, (int(s), int(s) + 1), maxl=64)
        L = dist.shape[1]
        assert otn[0] == tn[s] and otd[0] == td[s]
        assert np.array_equal(odist[0, :L], dist[s]) and not odist[0, L:].any()
    c.close()


def test_pyramid_paths_equal_default_on_c2():
    """Full-size C2: pyramid push + pull together must give exactly the integers of the default schedule."""
    flat = capi.prepare(plans.by_name("C2"))
    a = capi.Context(0)
    ref = a.build(flat).global_ints(-1)
    a.close()
    b = capi.Context(0)
    b.set_option("bfs_pull", 1)
    b.set_option("bfs_push", 1)
    got = b.build(flat).global_ints(-1)
    b.close()
    assert ref[3] == got[3]
    for x, y in zip(ref[:3], got[:3]):
        assert np.array_equal(x, y)
    c = capi.Context(0)
    for k, v in (("bfs_pull", 1), ("bfs_push", 1), ("bfs_coarse", 0), ("bfs_pyr_nodes", 1)):
        c.set_option(k, v)
    got = c.build(flat).global_ints(-1)
    c.close()
    assert ref[3] == got[3]
    for x, y in zip(ref[:3], got[:3]):
        assert np.array_equal(x, y)


def test_pyramid_pull_equals_default_on_c2():
    """Full-size C2: the pyramid pull must give exactly the integers of the default schedule."""
    flat = capi.prepare(plans.by_name("C2"))
    a = capi.Context(0)
    ga = a.build(flat)
    ref = ga.global_ints(-1)
    a.close()
    b = capi.Context(0)
    b.set_option("bfs_pull", 1)
    gb = b.build(flat)
    got = gb.global_ints(-1)
    b.close()
    assert ref[3] == got[3]
    for x, y in zip(ref[:3], got[:3]):
        assert np.array_equal(x, y)
