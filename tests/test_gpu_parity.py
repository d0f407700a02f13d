"""GPU parity tests (run with -m gpu on the B200 box).  Everything goes through the C ABI
(libvga_b200.so) or the host layer above it (libvga_host.so); the oracle and the golden fixtures
(generated from the unmodified reference) are only the checkers.

Bar: adjacency, node statistics, BFS integers and local integers bit-exact; float attributes
bit-equal as float32 (tolerance of the north star: 1e-9 relative -- asserted as exact equality,
which is stricter)."""
import numpy as np
import pytest

from conftest import GOLDEN, golden
from depthmapx_b200 import capi, plans

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    if capi.device_count() < 1:
        pytest.fail("no CUDA device: the GPU tests must run on the B200 box (there is no CPU fallback)")
    c = capi.Context(0)
    yield c
    c.close()


def flat_of(fx, maxdist=-1.0):
    return capi.FlatGrid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                         fx["state"], fx["line_off"], fx["lines"], maxdist)


def attr(fx, name):
    cols = [str(c) for c in fx["columns"]]
    return fx[f"attr_{cols.index(name)}"]


def sorted_rows(rowptr, ref, b):
    rowid = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr).astype(np.int64))
    order = np.lexsort((ref.astype(np.int64), rowid))
    return ref[order], b[order]


def my_rows(g):
    """CSR of the library with columns translated to packed PixelRefs and rows re-sorted by PixelRef
    (the library sorts by ordinal: filled cells in PixelRef order, ghost vertices last)."""
    rp, col, b, acc = g.csr()
    ref = g.cell_refs()[col] if len(col) else np.zeros(0, np.int32)
    r, bb = sorted_rows(rp, ref, b)
    return rp, r, bb


def oracle_graph(flat, **kw):
    from oracle import pyoracle as po
    return po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off,
                                  flat.lines, flat.maxdist), **kw)


# ---- against the reference's own outputs (golden fixtures) -----------------------------------------

@pytest.mark.parametrize("name", GOLDEN)
def test_makegraph_vs_reference_golden(ctx, name):
    fx = golden(name)
    g = ctx.build(flat_of(fx))
    rp, col, b, acc = g.csr()
    assert np.array_equal(rp, fx["rowptr"])
    refs = g.cell_refs()
    eref, ebin = sorted_rows(fx["rowptr"], fx["ref"], fx["bin"])
    assert np.array_equal(refs[col], eref)
    assert np.array_equal(b, ebin)
    st = g.node_stats()
    assert np.array_equal(st["connectivity"].astype(np.float32), attr(fx, "Connectivity"))
    assert np.array_equal(st["sum_d"].astype(np.float32), attr(fx, "Point First Moment"))
    assert np.array_equal(st["sum_d2"].astype(np.float32), attr(fx, "Point Second Moment"))
    assert np.array_equal(st["far"], fx["bin_dist"])
    assert np.array_equal(st["gridconn"], fx["gridconn"])
    # accepted pixels per bin == stored node counts for non-diagonal bins; diagonal bins store pixels.size()
    assert np.array_equal(st["bin_count"].astype(np.uint16), fx["bin_count"])


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("radius", [-1, 3])
def test_global_vs_reference_golden(ctx, name, radius):
    fx = golden(name)
    g = ctx.build(flat_of(fx))
    tn, td, dist, used = g.global_ints(radius)
    out = capi.global_attributes(tn, td, dist)
    sfx = "" if radius == -1 else f" R{radius}"
    for k, v in out.items():
        assert np.array_equal(v, attr(fx, k + sfx)), k


@pytest.mark.parametrize("name", GOLDEN)
def test_local_vs_reference_golden(ctx, name):
    fx = golden(name)
    g = ctx.build(flat_of(fx))
    out = capi.local_attributes(*g.local_ints())
    for k, v in out.items():
        assert np.array_equal(v, attr(fx, k)), k


@pytest.mark.parametrize("name", ["oblique20", "office24"])
def test_host_layer_columns_vs_reference_golden(name):
    """dmx::PointMap::sparkGraph2 + VGAVisualLocal::run + VGAVisualGlobal::run: every column the reference writes."""
    fx = golden(name)
    m = capi.HostMap(fx["walls"], float(fx["spacing"]))
    for s in fx["seeds"]:
        assert m.fill(float(s[0]), float(s[1]))
    assert m.make_graph()
    assert m.vga_local()
    assert m.vga_global(-1.0)
    assert m.vga_global(3.0)
    cols = [str(c) for c in fx["columns"]]
    assert sorted(m.columns()) == sorted(cols)
    for c in cols:
        assert np.array_equal(m.attr(c), attr(fx, c)), c
    assert np.array_equal(m.grid_connections(), fx["gridconn"])


def test_from_csr_adopts_reference_adjacency(ctx):
    """-m VGA on a loaded .graph: adjacency flattened from the reference's Nodes -> same analysis."""
    fx = golden("oblique16s07")
    state = fx["state"]
    rows = int(fx["rows"])
    ordmap = np.full(state.shape[0], -1, np.int64)
    filled = (state & 2) != 0
    ordmap[filled] = np.arange(filled.sum())
    ref = fx["ref"].astype(np.int64)
    cell = (ref >> 16) * rows + (ref & 0xffff)
    col = ordmap[cell]
    assert (col >= 0).all()
    n = int(filled.sum())
    g = ctx.graph_from_csr(n, 0, fx["rowptr"], col.astype(np.uint32), fx["bin"])
    tn, td, dist, used = g.global_ints(-1)
    out = capi.global_attributes(tn, td, dist)
    for k, v in out.items():
        assert np.array_equal(v, attr(fx, k)), k
    out = capi.local_attributes(*g.local_ints())
    for k, v in out.items():
        assert np.array_equal(v, attr(fx, k)), k


# ---- against the oracle on seeded plans ---------------------------------------------------------------

PLANS = ["oblique:30:30:7", "oblique:30:30:8:0.7", "oblique:40:25:21:1.3", "office:64:64:1", "room:40:40:5",
         "gallery:160:160:3", "urban:120:120:4"]


@pytest.mark.parametrize("name", PLANS)
def test_makegraph_vs_oracle(ctx, name):
    flat = capi.prepare(plans.by_name(name))
    g = ctx.build(flat)
    og = oracle_graph(flat)
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
    assert np.array_equal(st["gridconn"], a["gridconn"])


_ORACLE_CACHE = {}


def cached_oracle(name):
    if name not in _ORACLE_CACHE:
        flat = capi.prepare(plans.by_name(name))
        _ORACLE_CACHE[name] = (flat, oracle_graph(flat))
    return _ORACLE_CACHE[name]


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1"])
@pytest.mark.parametrize("radius", [-1, 2])
@pytest.mark.parametrize("mode,words,order", [(0, 1, 0), (1, 1, 1), (2, 1, 2), (0, 4, 2), (1, 2, 2), (2, 4, 2), (2, 0, 2)])
def test_global_vs_oracle_all_modes(name, radius, mode, words, order):
    c = capi.Context(0)
    c.set_option("bfs_mode", mode)
    c.set_option("bfs_words", words)
    c.set_option("bfs_order", order)
    flat, og = cached_oracle(name)
    g = c.build(flat)
    tn, td, dist, used = g.global_ints(radius)
    rng = np.random.RandomState(3)
    for s in rng.choice(g.n, min(g.n, 96), replace=False):
        otn, otd, odist, onl = og.global_ints(radius, (int(s), int(s) + 1), maxl=64)
        L = dist.shape[1]
        assert otn[0] == tn[s] and otd[0] == td[s]
        assert np.array_equal(odist[0, :L], dist[s]) and not odist[0, L:].any()
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1", "urban:120:120:4"])
@pytest.mark.parametrize("local_mode", [0, 1, 2])
def test_local_vs_oracle(name, local_mode):
    """The local kernels (2 = default: per-cell bitmaps fed by run-length rows; 0: fed by entries; 1: bit-parallel
    batches) against the oracle."""
    flat, og = cached_oracle(name)
    c = capi.Context(0)
    c.set_option("local_mode", local_mode)
    g = c.build(flat)
    lo = max(0, g.n // 2 - 100)
    hi = min(g.n, lo + 200)
    a = g.local_ints((lo, hi))
    b = og.local_ints((lo, hi))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    if g.n <= 5000:  # whole map through the batched path (several batches, partial last batch)
        for x, y in zip(g.local_ints(), og.local_ints()):
            assert np.array_equal(x, y)
    c.close()


@pytest.mark.skipif(bool(__import__("os").environ.get("VGA_EMU_LIBDIR")), reason="the SIMT emulation has no tensor cores")
@pytest.mark.parametrize("name", ["room:40:40:5", "oblique:30:30:7", "room:100:100:0", "office:64:64:1"])
def test_local_tensor_core_kernel_vs_oracle(name):
    """local_mode = 4: cluster / total as a masked int8 product on the tensor cores (tcgen05.mma kind::i8, local_tc.cu),
    operands expanded from bit matrices into shared memory, empty K chunks skipped; every cell of the map (several row
    and column tiles, partial last tiles) and a slice against the oracle.  room:100 = BASELINE config 1 (auto mode picks
    this kernel there: deg/N = 0.32)."""
    flat, og = cached_oracle(name)
    c = capi.Context(0)
    c.set_option("local_mode", 4)
    g = c.build(flat)
    whole = g.n <= 5000
    rng = (0, g.n) if whole else (g.n // 2 - 300, g.n // 2 + 300)
    for x, y in zip(g.local_ints(rng), og.local_ints(rng)):
        assert np.array_equal(x, y)
    for x, y in zip(g.local_ints((7, 8)), og.local_ints((7, 8))):  # a single cell
        assert np.array_equal(x, y)
    c.close()


def test_local_tensor_core_kernel_with_ghost_columns():
    """Unfilled cells inside diagonal runs are columns of the product (they count in k / total) but never middle vertices."""
    if __import__("os").environ.get("VGA_EMU_LIBDIR"):
        pytest.skip("the SIMT emulation has no tensor cores")
    from oracle import pyoracle as po
    flat = capi.prepare(plans.by_name("room:24:24:1"))
    st = flat.state.copy()
    for d in (3, 5):
        st[(2 + d) * flat.rows + (2 + d)] &= ~np.uint16(2)
    flat.state = st
    og = oracle_graph(flat)
    c = capi.Context(0)
    c.set_option("local_mode", 4)
    g = c.build(flat)
    assert g.ghosts > 0
    for x, y in zip(g.local_ints(), og.local_ints()):
        assert np.array_equal(x, y)
    c.close()


@pytest.mark.parametrize("name", ["oblique:30:30:7", "office:64:64:1"])
def test_local_runs_in_several_passes(name):
    """The run-length local kernel covering the vertex universe in several column ranges (as at 10^6 cells, where two
    bitmaps of the whole universe exceed shared memory): runs clipped at the range borders, counts summed over the passes."""
    flat, og = cached_oracle(name)
    c = capi.Context(0)
    c.set_option("local_span", 320)
    g = c.build(flat)
    hi = min(g.n, 400)
    for x, y in zip(g.local_ints((0, hi)), og.local_ints((0, hi))):
        assert np.array_equal(x, y)
    c.close()


def test_maxdist_vs_oracle(ctx):
    flat = capi.prepare(plans.by_name("oblique:30:30:7"), maxdist=6.5)
    g = ctx.build(flat)
    og = oracle_graph(flat)
    rp, col, b, acc = g.csr()
    orp, oref, ob = og.iter_rows()
    assert np.array_equal(rp, orp)
    eref, ebin = sorted_rows(orp, oref, ob)
    assert np.array_equal(g.cell_refs()[col], eref)


def test_boundary_graph_vs_oracle(ctx):
    """-pb: only EDGE cells keep FILLED before construction (pointdata.cpp:1254-1264)."""
    flat = capi.prepare(plans.by_name("office:48:48:2"))
    st = flat.state.copy()
    drop = ((st & 2) != 0) & ((st & 0x20) == 0)
    st[drop] &= ~np.uint16(2)
    flat.state = st
    g = ctx.build(flat)
    og = oracle_graph(flat)
    rp, ref, b = my_rows(g)
    orp, oref, ob = og.iter_rows()
    assert np.array_equal(rp, orp)
    eref, ebin = sorted_rows(orp, oref, ob)
    assert np.array_equal(ref, eref) and np.array_equal(b, ebin)
    # cells between accepted diagonal cells may now be unfilled: ghosts are allowed to appear
    assert g.ghosts >= 0
    tn, td, dist, used = g.global_ints(-1)
    otn, otd, odist, onl = og.global_ints(-1, maxl=dist.shape[1])
    assert np.array_equal(tn, otn) and np.array_equal(td, otd) and np.array_equal(dist, odist)
    lo, hi = 0, min(g.n, 150)
    for x, y in zip(g.local_ints((lo, hi)), og.local_ints((lo, hi))):
        assert np.array_equal(x, y)


def test_ghosts_diagonal_fill_vs_oracle(ctx):
    """Unfilled cells inside a diagonal bin's first..last run are part of the iterated adjacency
    (Bin::make, ngraph.cpp:243-259) and of the local measures' k."""
    flat = capi.prepare(plans.by_name("room:24:24:1"))
    st = flat.state.copy()
    rows = flat.rows
    # punch unfilled holes on a diagonal of an open area, away from walls
    for d in (3, 5):
        st[(2 + d) * rows + (2 + d)] &= ~np.uint16(2)
    flat.state = st
    g = ctx.build(flat)
    og = oracle_graph(flat)
    assert g.ghosts > 0
    rp, ref, b = my_rows(g)
    orp, oref, ob = og.iter_rows()
    assert np.array_equal(rp, orp)
    eref, ebin = sorted_rows(orp, oref, ob)
    assert np.array_equal(ref, eref) and np.array_equal(b, ebin)
    tn, td, dist, used = g.global_ints(-1)
    otn, otd, odist, onl = og.global_ints(-1, maxl=dist.shape[1])
    assert np.array_equal(tn, otn) and np.array_equal(td, otd) and np.array_equal(dist, odist)
    for x, y in zip(g.local_ints(), og.local_ints()):
        assert np.array_equal(x, y)


def test_empty_and_tiny_inputs(ctx):
    # no filled cell at all
    flat = capi.prepare(plans.by_name("room:12:12:0"))
    flat.state = (flat.state & ~np.uint16(2)).astype(np.uint16)
    g = ctx.build(flat)
    assert g.n == 0 and g.entries == 0
    tn, td, dist, used = g.global_ints(-1)
    assert len(tn) == 0
    # a single filled cell: Node Count 1, everything else -1
    flat = capi.prepare(plans.by_name("room:12:12:0"))
    st = (flat.state & ~np.uint16(2)).astype(np.uint16)
    st[5 * flat.rows + 5] |= 2
    flat.state = st
    g = ctx.build(flat)
    assert g.n == 1 and g.entries == 0
    tn, td, dist, used = g.global_ints(-1)
    out = capi.global_attributes(tn, td, dist)
    assert out["Visual Node Count"][0] == 1.0 and out["Visual Mean Depth"][0] == -1.0
    lo = capi.local_attributes(*g.local_ints())
    assert lo["Visual Control"][0] == -1.0


def test_overflow_path_small_capacity():
    """Tasks that overflow the shared-memory gap/block slices are re-run with global scratch and
    give the same rows."""
    flat = capi.prepare(plans.by_name("urban:120:120:4"))
    a = capi.Context(0)
    g1 = a.build(flat)
    b = capi.Context(0)
    b.set_option("sieve_mode", 0)
    b.set_option("sieve_gcap", 4)
    b.set_option("sieve_bcap", 8)
    g2 = b.build(flat)
    for x, y in zip(g1.csr(), g2.csr()):
        assert np.array_equal(x, y)
    s1, s2 = g1.node_stats(), g2.node_stats()
    for k in s1:
        assert np.array_equal(s1[k], s2[k]), k


@pytest.mark.parametrize("name", ["oblique:30:30:7", "oblique:40:25:21:1.3", "office:64:64:1", "urban:120:120:4"])
def test_sieve_kernels_agree_and_match_oracle(name):
    """The thread-per-octant kernel (default, falls back to the warp kernel on overflow) and the
    warp-per-octant kernel must produce identical rows and statistics, equal to the oracle's."""
    flat = capi.prepare(plans.by_name(name))
    og = oracle_graph(flat)
    orp, oref, ob = og.iter_rows()
    eref, ebin = sorted_rows(orp, oref, ob)
    a = og.node_attrs()
    for mode in (0, 1):
        c = capi.Context(0)
        c.set_option("sieve_mode", mode)
        g = c.build(flat)
        rp, ref, b = my_rows(g)
        assert np.array_equal(rp, orp)
        assert np.array_equal(ref, eref) and np.array_equal(b, ebin)
        st = g.node_stats()
        assert np.array_equal(st["sum_d"].astype(np.float32), a["first_moment"])
        assert np.array_equal(st["sum_d2"].astype(np.float32), a["second_moment"])
        assert np.array_equal(st["far"], a["far"])
        c.close()


# ---- size-independent properties at larger sizes -----------------------------------------------------

def test_shards_equal_full_build(ctx):
    flat = capi.prepare(plans.by_name("office:96:96:3"))
    full = ctx.build(flat)
    n = full.n
    frp, fcol, fb, facc = full.csr()
    cuts = [0, n // 3, n // 3 + 1, n]
    base = 0
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        sh = ctx.build(flat, (lo, hi))
        rp, col, b, acc = sh.csr()
        assert np.array_equal(rp + frp[lo], frp[lo:hi + 1])
        assert np.array_equal(col, fcol[int(frp[lo]):int(frp[hi])])
        base += len(col)
    assert base == len(fcol)


def test_properties_c1_full_size(ctx):
    """BASELINE config 1 at full size: histogram sums, symmetry-implied reachability, determinism."""
    flat = capi.prepare(plans.by_name("C1"))
    g = ctx.build(flat)
    tn, td, dist, used = g.global_ints(-1)
    assert (dist.sum(axis=1) == tn).all()
    assert ((dist * np.arange(dist.shape[1])).sum(axis=1) == td).all()
    assert (tn == g.n).all()  # one connected open plan
    tn3, td3, dist3, used3 = g.global_ints(3)
    assert used3 <= 4 and (tn3 <= tn).all()
    L = min(dist.shape[1], dist3.shape[1], 3)
    assert np.array_equal(dist3[:, :L], dist[:, :L])  # radius only truncates
    g2 = ctx.build(flat)
    for x, y in zip(g.csr(), g2.csr()):
        assert np.array_equal(x, y)
    st = g.node_stats()
    rp = g.csr()[0]
    assert np.array_equal(np.diff(rp).astype(np.int64), st["connectivity"].astype(np.int64))  # no fill-ins here
    # Connectivity must equal the degree the BFS sees at level 1
    assert np.array_equal(dist[:, 1].astype(np.int64), st["connectivity"].astype(np.int64))


def test_properties_c2_full_size(ctx):
    """BASELINE config 2 (the bench workload) at full size: histogram identities, full reachability, and
    equality of two entirely different BFS schedules (default: coherent clusters + direction-optimising hybrid
    vs. plain x-major top-down-only single-word batches).  Both local kernels must agree on a slice (config 3)."""
    flat = capi.prepare(plans.by_name("C2"))
    g = ctx.build(flat)
    tn, td, dist, used = g.global_ints(-1)
    assert (dist.sum(axis=1) == tn).all()
    assert ((dist * np.arange(dist.shape[1])).sum(axis=1) == td).all()
    assert (tn == g.n).all()
    c2 = capi.Context(0)
    for k, v in (("bfs_mode", 0), ("bfs_words", 1), ("bfs_order", 0), ("local_mode", 1)):
        c2.set_option(k, v)
    g2 = c2.build(flat)
    tn2, td2, dist2, used2 = g2.global_ints(-1)
    assert used2 == used
    assert np.array_equal(tn, tn2) and np.array_equal(td, td2) and np.array_equal(dist, dist2)
    a = g.local_ints((1000, 1256))    # default: per-cell bitmaps at this degree
    b = g2.local_ints((1000, 1256))   # bit-parallel batches
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    c2.close()


def test_properties_c4_radius_truncates(ctx):
    """BASELINE config 4 (512x512 gallery, 3.9e9 edges): radius 3 vs radius n on a slice of sources --
    a radius only truncates the level histogram; the four-word batches are exercised at scale."""
    flat = capi.prepare(plans.by_name("C4"))
    g = ctx.build(flat)
    src = (g.n // 2, g.n // 2 + 2048)
    tn, td, dist, used = g.global_ints(-1, src)
    tn3, td3, dist3, used3 = g.global_ints(3, src)
    assert used3 <= 4
    assert (tn == g.n).all() and (tn3 <= tn).all()
    assert np.array_equal(dist3[:, :3], dist[:, :3])
    assert (dist.sum(axis=1) == tn).all() and (dist3.sum(axis=1) == tn3).all()
    g.free()


@pytest.mark.parametrize("words,chunk", [(1, 2), (2, 4), (0, 3)])
def test_global_in_several_chunks(words, chunk):
    """The batches of one call processed in several chunks (as at C5 scale, where the BFS state of all batches does not
    fit at once): per-chunk state, statistics and the early retirement of finished batches must not leak across chunks."""
    flat, og = cached_oracle("oblique:30:30:7")
    c = capi.Context(0)
    c.set_option("bfs_words", words)
    c.set_option("bfs_chunk", chunk)
    g = c.build(flat)
    for radius in (-1, 2):
        tn, td, dist, used = g.global_ints(radius)
        otn, otd, odist, onl = og.global_ints(radius, maxl=dist.shape[1])
        assert np.array_equal(tn, otn) and np.array_equal(td, otd) and np.array_equal(dist, odist)
    c.close()
