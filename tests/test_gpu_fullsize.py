"""GPU parity at the BASELINE.json configuration sizes (C2 256x256 office, C4 512x512 gallery, C5 1024x1024 urban),
against the oracle on the same inputs -- not only through invariants.

What runs only at these sizes and is pinned here: W = 2 / W = 4 word widths chosen from the graph, several BFS chunks
under the real memory budget, adjacency slices with more than 2^31 entries, 64-bit row offsets (E = 5.5e9 at C5).

  makegraph   C2: the complete adjacency (columns, bins) and every node statistic against the oracle's makegraph;
              C4 / C5: row samples against vgao_makegraph_range (the oracle needs ~1 ms per row, the full C5 graph hours);
  VGA global  sampled sources (radius n and 3) against the oracle's BFS over the SAME adjacency (exported CSR);
  VGA local   sampled cells against the oracle's local counts over the same adjacency.

The oracle side is oracle/vga_oracle.c (vgao_global_csr / vgao_local_csr, pinned against the graph form of the
oracle by tests/test_oracle.py, which in turn is pinned against the compiled reference)."""
import numpy as np
import pytest

from depthmapx_b200 import capi, plans

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(2400, method="thread")]


def oracle_grid(flat):
    from oracle import pyoracle as po
    return po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines, flat.maxdist)


def sorted_rows(rowptr, ref, b):
    rowid = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr).astype(np.int64))
    order = np.lexsort((ref.astype(np.int64), rowid))
    return ref[order], b[order]


def sample_blocks(n, blocks, width, seed):
    """`blocks` runs of `width` consecutive ordinals at random places (the C ABI takes source ranges)."""
    rng = np.random.RandomState(seed)
    starts = np.sort(rng.choice(max(1, n - width), blocks, replace=False))
    return [(int(s), int(min(n, s + width))) for s in starts]


def check_rows_vs_oracle(g, flat, ranges):
    """Rows [lo, hi) of the library's graph (columns as PixelRefs, bins, node statistics) against the oracle's makegraph."""
    from oracle import pyoracle as po
    refs = g.cell_refs()
    st = g.node_stats()
    for lo, hi in ranges:
        sh = g.ctx.build(flat, (lo, hi))  # the shard build is checked against the oracle ...
        rp, col, b, acc = sh.csr()
        og = po.OracleGraph(oracle_grid(flat), src_range=(lo, hi))
        orp, oref, ob = og.iter_rows()
        orp = orp[lo:hi + 1] - orp[lo]
        assert np.array_equal(rp, orp)
        eref, ebin = sorted_rows(orp, oref, ob)
        assert np.array_equal(refs[col], eref) and np.array_equal(b, ebin)
        a = og.node_attrs()
        sst = sh.node_stats()
        for mine in (sst, {k: v[lo:hi] for k, v in st.items()}):  # ... and so are the statistics of the full build
            assert np.array_equal(mine["connectivity"].astype(np.float32), a["connectivity"][lo:hi])
            assert np.array_equal(mine["sum_d"].astype(np.float32), a["first_moment"][lo:hi])
            assert np.array_equal(mine["sum_d2"].astype(np.float32), a["second_moment"][lo:hi])
            assert np.array_equal(mine["far"], a["far"][lo:hi])
            assert np.array_equal(mine["gridconn"], a["gridconn"][lo:hi])
        sh.free()


def check_analyses_vs_oracle(g, rp, col, src_ranges, cell_ranges, radii=(-1, 3)):
    from oracle import pyoracle as po
    n = g.n
    for radius in radii:
        for lo, hi in src_ranges:
            tn, td, dist, used = g.global_ints(radius, (lo, hi))
            otn, otd, odist = po.global_csr(n, rp, col, np.arange(lo, hi), radius, maxl=64)
            L = dist.shape[1]
            assert np.array_equal(tn, otn) and np.array_equal(td, otd), (radius, lo)
            assert np.array_equal(odist[:, :L], dist) and not odist[:, L:].any(), (radius, lo)
    refs = g.cell_refs()
    for lo, hi in cell_ranges:
        mine = g.local_ints((lo, hi))
        theirs = po.local_csr(n, n + g.ghosts, rp, col, refs, np.arange(lo, hi))
        for x, y in zip(mine, theirs):
            assert np.array_equal(x, y), lo


def test_c2_complete_adjacency_statistics_and_sampled_analyses():
    """BASELINE config 2/3 (the 256x256 office plan, 65,536 cells, E = 2.8e7): the whole graph against the oracle's
    makegraph, 64 sampled sources of VGA global (radius n and 3) and 64 sampled cells of VGA local."""
    from oracle import pyoracle as po
    flat = capi.prepare(plans.by_name("C2"))
    c = capi.Context(0)
    c.set_option("bfs_hybrid", 2)  # x-major / y-major lists whatever the number of sources of the first call
    g = c.build(flat)
    og = po.OracleGraph(oracle_grid(flat))
    rp, col, b, acc = g.csr()
    orp, oref, ob = og.iter_rows()
    assert np.array_equal(rp, orp)
    eref, ebin = sorted_rows(orp, oref, ob)
    assert np.array_equal(g.cell_refs()[col], eref) and np.array_equal(b, ebin)
    st, a = g.node_stats(), og.node_attrs()
    assert np.array_equal(st["connectivity"].astype(np.float32), a["connectivity"])
    assert np.array_equal(st["sum_d"].astype(np.float32), a["first_moment"])
    assert np.array_equal(st["sum_d2"].astype(np.float32), a["second_moment"])
    assert np.array_equal(st["far"], a["far"])
    assert np.array_equal(st["bin_count"].astype(np.uint16), a["bin_count"])
    assert np.array_equal(st["gridconn"], a["gridconn"])
    check_analyses_vs_oracle(g, rp, col, sample_blocks(g.n, 8, 8, 1), sample_blocks(g.n, 8, 8, 2))
    # the complete run (all 65,536 sources in one call, as bench.py does) agrees with the sampled calls
    tn, td, dist, used = g.global_ints(-1)
    for lo, hi in sample_blocks(g.n, 8, 8, 1):
        otn, otd, odist = po.global_csr(g.n, rp, col, np.arange(lo, hi), -1, maxl=64)
        assert np.array_equal(tn[lo:hi], otn) and np.array_equal(td[lo:hi], otd)
        assert np.array_equal(dist[lo:hi], odist[:, :dist.shape[1]])
    c.close()


def test_c4_sampled_rows_sources_and_cells():
    """BASELINE config 4 (512x512 gallery, 262,144 cells, E = 3.9e9, W = 4): row samples against the oracle's makegraph,
    64 sampled sources (radius n and 3) and 32 sampled cells of VGA local against the oracle on the exported adjacency."""
    flat = capi.prepare(plans.by_name("C4"))
    c = capi.Context(0)
    g = c.build(flat)
    check_rows_vs_oracle(g, flat, sample_blocks(g.n, 3, 48, 5))
    rp, col, _, _ = g.csr(bins=False)
    check_analyses_vs_oracle(g, rp, col, sample_blocks(g.n, 8, 8, 3), sample_blocks(g.n, 4, 8, 4))
    c.close()


@pytest.mark.slow
def test_c5_sampled_rows_sources_and_cells():
    """BASELINE config 5 (1024x1024 urban, 1,048,576 cells, E = 5.5e9; the north-star configuration): W = 2 batches,
    several BFS chunks, adjacency slices beyond 2^31 entries and 64-bit row offsets, pinned against the oracle."""
    flat = capi.prepare(plans.by_name("C5"))
    c = capi.Context(0)
    c.set_option("bfs_hybrid", 2)  # the y-major lists at full size (C4 above runs x-major only: few sources per call)
    g = c.build(flat)
    assert g.entries > 2 ** 32
    check_rows_vs_oracle(g, flat, sample_blocks(g.n, 4, 32, 7) + [(g.n - 16, g.n)])
    rp, col, _, _ = g.csr(bins=False)
    check_analyses_vs_oracle(g, rp, col, sample_blocks(g.n, 4, 8, 8) + [(g.n - 4, g.n)], sample_blocks(g.n, 4, 8, 9))
    # a multi-chunk call (more batches than fit in the BFS state budget at once) over a slice of sources
    from oracle import pyoracle as po
    c.set_option("bfs_chunk", 8)
    lo = g.n // 3
    tn, td, dist, used = g.global_ints(-1, (lo, lo + 4096))
    pick = np.arange(lo, lo + 4096, 257)
    otn, otd, odist = po.global_csr(g.n, rp, col, pick, -1, maxl=64)
    assert np.array_equal(tn[pick - lo], otn) and np.array_equal(td[pick - lo], otd)
    assert np.array_equal(dist[pick - lo], odist[:, :dist.shape[1]])
    c.close()
