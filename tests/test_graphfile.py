"""CPU: the host layer's own .graph codec (depthmapx_b200/host/graphio.cpp -- GraphFile, PointMap::read /
write, the Node run-length codec, attribute-table statistics) against files the UNMODIFIED reference CLI wrote
(tests/golden/graphfiles.npz, made by tests/golden/make_golden_graphs.py), and -- where the compiled reference is present
-- against the reference's own MetaGraph reading the same files.

The attribute stages are fed with integers from the oracle here (the checker standing in for libvga_b200, exactly the
role it has in the GPU parity tests); tests/test_zzz_graphfile_gpu.py runs the same pipelines with the GPU library."""
import os

import numpy as np
import pytest

from conftest import golden
from depthmapx_b200 import capi
from oracle import pyoracle as po

CASES = ["oblique12", "oblique10s07", "office16"]
KINDS = ["plan", "fill", "prep", "prep_pb", "vga", "vga3", "sd", "prep_l", "vga_l", "vga_l2", "sd_l"]


def bfs_ints(rp, col, n, radius=-1, maxl=64):
    """Plain level BFS from every vertex of a small CSR (checker for the contracted adjacency): Node Count, total depth
    and level histogram with the reference's radius rule (cells at the radius are counted, not expanded)."""
    tn = np.zeros(n, np.int32)
    td = np.zeros(n, np.int64)
    dist = np.zeros((n, maxl), np.int32)
    for s in range(n):
        lev = np.full(n, -1, np.int64)
        lev[s] = 0
        fr, level = [s], 0
        while fr:
            dist[s, level] = len(fr)
            tn[s] += len(fr)
            td[s] += level * len(fr)
            nx = []
            if radius == -1 or level < radius:
                for u in fr:
                    for c in col[rp[u]:rp[u + 1]]:
                        if c < n and lev[c] < 0:
                            lev[c] = level + 1
                            nx.append(int(c))
            fr, level = nx, level + 1
    return tn, td, dist


def level_from(rp, col, seeds):
    """Level of every vertex in a BFS from the set `seeds` (-1 = unreached)."""
    n = len(rp) - 1
    lev = np.full(n, -1, np.int32)
    fr = [int(s) for s in seeds]
    for s in fr:
        lev[s] = 0
    level = 0
    while fr:
        nx = []
        for u in fr:
            for c in col[rp[u]:rp[u + 1]]:
                if c < n and lev[c] < 0:
                    lev[c] = level + 1
                    nx.append(int(c))
        fr, level = nx, level + 1
    return lev


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("graphs")
    fx = golden("graphfiles")
    for k in fx.files:
        if not k.endswith("__args"):
            open(os.path.join(d, k + ".graph"), "wb").write(fx[k].tobytes())
    return str(d), {c: [str(x) for x in fx[c + "__args"]] for c in CASES}


def data(path):
    return open(path, "rb").read()


def oracle_of(m):
    flat = m.flat()
    return po.OracleGraph(po.Grid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, flat.state, flat.line_off, flat.lines))


def accepted_flags(og):
    rp, ref, b = og.iter_rows()
    arp, aref, ab = og.acc_rows()

    def keys(rp, ref, b):
        row = np.repeat(np.arange(len(rp) - 1, dtype=np.int64), np.diff(rp).astype(np.int64))
        return (row << 40) | (b.astype(np.int64) << 32) | (ref.astype(np.int64) & 0xffffffff)
    return rp, ref, b, np.isin(keys(rp, ref, b), keys(arp, aref, ab)).astype(np.uint8)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kind", ["fill", "prep", "prep_pb", "vga"])
def test_read_write_round_trip_is_byte_identical(files, tmp_path, case, kind):
    d, _ = files
    src = os.path.join(d, f"{case}__{kind}.graph")
    g = capi.GraphFile(src)
    assert g.num_maps == 1 and g.displayed_map == 0
    out = str(tmp_path / "rt.graph")
    g.save(out)
    assert data(out) == data(src)


@pytest.mark.parametrize("case", CASES)
def test_loaded_map_matches_what_was_stored(files, case):
    """Grid, states, attribute columns and adjacency of a loaded map: self-consistency with the file's own columns
    (Connectivity = number of accepted neighbours = sum of the stored bin counts)."""
    d, _ = files
    m = capi.GraphFile(os.path.join(d, f"{case}__vga.graph")).map()
    n = m.n
    assert n == int(((m.state() & 2) != 0).sum())
    rp, ref, b = m.flat_rows()
    cnt, dist = m.bins()
    assert len(rp) == n + 1 and rp[-1] == len(ref)
    assert np.array_equal(m.attr("Connectivity"), cnt.sum(axis=1).astype(np.float32))
    # iterated pixels per bin >= stored count (diagonal runs may cover fill-in pixels), equal for the other bins
    per_bin = np.zeros((n, 32), np.int64)
    row = np.repeat(np.arange(n), np.diff(rp).astype(np.int64))
    np.add.at(per_bin, (row, b), 1)
    diag = np.isin(np.arange(32), [4, 12, 20, 28])
    assert np.array_equal(per_bin[:, ~diag], cnt[:, ~diag]) and (per_bin[:, diag] >= cnt[:, diag]).all()
    assert {"Visual Integration [HH]", "Visual Control", "Point First Moment"} <= set(m.columns())


@pytest.mark.parametrize("case", CASES)
def test_node_encoder_reproduces_the_reference_runs(files, tmp_path, case):
    """Flatten the stored nodes, shuffle every row, re-encode (Node::make / Bin::make): the written bytes must be the
    reference's.  This is the encoder the host layer uses after a GPU build."""
    d, _ = files
    src = os.path.join(d, f"{case}__prep.graph")
    g = capi.GraphFile(src)
    m = g.map()
    rp, ref, b = m.flat_rows()
    cnt, dist = m.bins()
    # which pixels of a diagonal run are fill-ins (not accepted by the sieve, excluded from the stored count) is not in
    # the file: take the flags from the oracle built from the same plan (the stored map has no wall lists any more)
    _, args = files
    from depthmapx_b200 import plans
    spec, grid, seed, _sdp = args[case][:4]
    plan = plans.by_name(spec)
    hm = capi.HostMap(plan.walls, float(grid))
    assert hm.fill(*[float(x) for x in seed.split(",")])
    hm.begin_graph(False)
    orp, oref, ob, acc = accepted_flags(oracle_of(hm))
    assert np.array_equal(orp, rp) and np.array_equal(oref, ref) and np.array_equal(ob, b)  # oracle == stored adjacency
    rng = np.random.default_rng(5)
    perm = np.arange(len(ref))
    for v in range(len(rp) - 1):
        lo, hi = int(rp[v]), int(rp[v + 1])
        perm[lo:hi] = lo + rng.permutation(hi - lo)
    m.encode_nodes(rp, ref[perm], b[perm], acc[perm], dist)
    out = str(tmp_path / "enc.graph")
    g.save(out)
    assert data(out) == data(src)


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("boundary", [False, True])
def test_pipeline_from_the_drawing_writes_the_reference_bytes(files, tmp_path, case, boundary):
    """plan.graph -> new map, grid, fill -> makegraph halves + node encoder -> local + global (n, then 3) -> step depth:
    every saved file byte-identical to what the reference CLI wrote for the same commands."""
    d, args = files
    spec, grid, seed, sdp = args[case][:4]
    g = capi.GraphFile(os.path.join(d, f"{case}__plan.graph"))
    assert g.num_maps == 0 and len(g.walls()) > 0
    m = g.new_map(float(grid))
    assert m.fill(*[float(x) for x in seed.split(",")])
    out = str(tmp_path / "o.graph")
    if not boundary:
        g.save(out)
        assert data(out) == data(os.path.join(d, f"{case}__fill.graph"))
    m.begin_graph(boundary)
    og = oracle_of(m)
    na = og.node_attrs()
    rp, ref, b, acc = accepted_flags(og)
    m.finish_graph(boundary, na["connectivity"].astype(np.int32), na["first_moment"].astype(np.float64),
                   na["second_moment"].astype(np.float64), na["gridconn"])
    m.encode_nodes(rp, ref, b, acc, na["far"])
    g.graph_made()
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__prep_pb.graph" if boundary else f"{case}__prep.graph"))
    if boundary:
        return
    m.write_local(False, *og.local_ints())
    tn, td, dist, _ = og.global_ints(-1)
    m.write_global(-1.0, False, tn, td, dist)
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga.graph"))
    tn, td, dist, _ = og.global_ints(3)
    m.write_global(3.0, False, tn, td, dist)
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga3.graph"))
    # step depth on the loaded prep file, as the CLI does it
    g2 = capi.GraphFile(os.path.join(d, f"{case}__prep.graph"))
    m2 = g2.map()
    m2.select([[float(x) for x in sdp.split(",")]])
    sel = m2.selection()
    assert len(sel) == 1
    src = [int(np.searchsorted(og.cell_refs(), s)) for s in sel]
    m2.write_step_depth(og.step_depth(src))
    g2.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__sd.graph"))


@pytest.mark.parametrize("case", CASES)
def test_merge_links_write_the_reference_bytes(files, tmp_path, case):
    """-m LINK, then VGA global (radius n and 2) + local and STEPDEPTH on the linked map.  The BFS analyses
    run on the contracted adjacency (a merged pair is one vertex), radius-limited runs add the at-the-radius correction;
    every file byte-identical to the reference CLI's."""
    d, args = files
    out = str(tmp_path / "o.graph")
    links = [[float(x) for x in l.split(",")] for l in args[case][4:]]
    g = capi.GraphFile(os.path.join(d, f"{case}__prep.graph"))
    m = g.map()
    for l in links:
        assert m.merge(*l)
    with pytest.raises(RuntimeError):
        m.merge(*links[0])  # already linked
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__prep_l.graph"))
    n = m.n
    rp, col, primary = m.contracted_rows()
    assert int((primary != np.arange(n)).sum()) == len(links)
    rp, col = rp.astype(np.int64), col.astype(np.int64)
    # local is computed on the original adjacency (merges are ignored there): integers from the oracle of the same plan
    from depthmapx_b200 import plans
    spec, grid, seed = args[case][:3]
    hm = capi.HostMap(plans.by_name(spec).walls, float(grid))
    assert hm.fill(*[float(x) for x in seed.split(",")])
    hm.begin_graph(False)
    og = oracle_of(hm)
    m.write_local(False, *og.local_ints())
    tn, td, dist = bfs_ints(rp, col, n)
    assert set(np.unique(tn[primary])) == {n - len(links)}  # every pair counts once
    m.write_global(-1.0, False, tn[primary], td[primary], dist[primary])
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga_l.graph"))
    # radius 2 on the linked map
    g = capi.GraphFile(os.path.join(d, f"{case}__prep_l.graph"))
    m = g.map()
    tn, td, dist = bfs_ints(rp, col, n, 2)
    m.radius_correction(2, level_from, tn, td, dist)
    m.write_global(2.0, False, tn[primary], td[primary], dist[primary])
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga_l2.graph"))
    # step depth on the linked map
    g = capi.GraphFile(os.path.join(d, f"{case}__prep_l.graph"))
    m = g.map()
    m.select([[float(x) for x in args[case][3].split(",")]])
    src = [int(primary[int(np.searchsorted(og.cell_refs(), s))]) for s in m.selection()]
    m.write_step_depth(level_from(rp, col, src)[primary])
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__sd_l.graph"))


def test_gates_only_writes_columns_only(files):
    """gates_only makes the reference skip every cell (vgavisualglobal.cpp:75-78, vgavisuallocal.cpp:43-46)."""
    d, _ = files
    m = capi.GraphFile(os.path.join(d, "oblique12__prep.graph")).map()
    from depthmapx_b200.capi import host
    import ctypes as C
    assert host().dmxh_map_write_global(m.h, -1.0, 0, None, None, None, 0) == 1
    assert host().dmxh_map_write_local(m.h, 0, None, None, None, None) == 1
    assert len(m.columns()) == 13
    for c in m.columns()[3:]:
        assert (m.attr(c) == -1).all()


def test_simple_version_and_radius_columns(files):
    d, _ = files
    m = capi.GraphFile(os.path.join(d, "oblique12__prep.graph")).map()
    n = m.n
    m.write_global(5.0, True, np.full(n, 3, np.int32), np.full(n, 4, np.int64), np.tile(np.array([1, 1, 1, 0], np.int32), (n, 1)))
    assert m.columns() == ["Connectivity", "Point First Moment", "Point Second Moment", "Visual Integration [HH] R5"]
    m.write_local(True, np.zeros(n, np.int64), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32))
    assert len(m.columns()) == 4


def test_rejects_what_is_not_a_graph(tmp_path):
    p = tmp_path / "x.graph"
    p.write_bytes(b"not a graph file")
    with pytest.raises(RuntimeError):
        capi.GraphFile(str(p))
    p.write_bytes(b"grf" + (500).to_bytes(4, "little") + bytes(16))
    with pytest.raises(RuntimeError):
        capi.GraphFile(str(p))
    fx = golden("graphfiles")
    p.write_bytes(fx["oblique12__prep"].tobytes()[:20000])  # truncated inside the point map
    with pytest.raises(RuntimeError):
        capi.GraphFile(str(p))


def test_vga_on_loaded_map_needs_the_gpu(files):
    """No CPU compute path: the analyses of a loaded map throw without a device."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    d, _ = files
    m = capi.GraphFile(os.path.join(d, "oblique12__prep.graph")).map()
    for call in (lambda: m.vga_global(-1.0), lambda: m.vga_local(), lambda: m.step_depth([(3, 3)])):
        with pytest.raises(RuntimeError):
            call()


# ---- against the compiled reference reading the same files (only where oracle/_ref/libdmxref.so exists) ----------------

@pytest.mark.ref
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kind", KINDS)
def test_rewrite_equals_the_references_rewrite(files, tmp_path, case, kind):
    """read + write with nothing in between must give what MetaGraph::readFromFile + write gives, including the quirks:
    SELECTED flags dropped, displayed-attribute index re-mapped, unnamed drawings renamed "<unknown>"."""
    if not po.have_ref():
        pytest.skip("compiled reference not present")
    d, _ = files
    src = os.path.join(d, f"{case}__{kind}.graph")
    ours, theirs = str(tmp_path / "a.graph"), str(tmp_path / "b.graph")
    capi.GraphFile(src).save(ours)
    assert po.ref_graph_rewrite(src, theirs)
    assert data(ours) == data(theirs)


@pytest.mark.ref
@pytest.mark.parametrize("case,seed", [("office16", 1), ("oblique12", 2), ("oblique10s07", 3)])
def test_random_merge_links_against_the_reference(files, case, seed):
    """Random link pairs and radii: the reference's VGAVisualGlobal::run on its own merged map against the contracted
    adjacency + at-the-radius correction of the host layer."""
    if not po.have_ref():
        pytest.skip("compiled reference not present")
    d, _ = files
    src = os.path.join(d, f"{case}__prep.graph")
    rng = np.random.default_rng(seed)
    corrected = 0
    for trial in range(16):
        m = capi.GraphFile(src).map()
        r = po.RefMap(graph_file=src)
        n = m.n
        st = m.state().reshape(m.cols, m.rows)
        filled = np.argwhere((st & 2) != 0)
        npairs, done = int(rng.integers(1, 5)), 0
        while done < npairs:
            a, b = filled[rng.integers(len(filled))], filled[rng.integers(len(filled))]
            if (a == b).all():
                continue
            pa = (m.bl_x + a[0] * m.spacing, m.bl_y + a[1] * m.spacing)
            pb = (m.bl_x + b[0] * m.spacing, m.bl_y + b[1] * m.spacing)
            if r.merge(*pa, *pb):
                assert m.merge(*pa, *pb)
                done += 1
        radius = int(rng.choice([-1, 1, 1, 2, 3]))
        r.vga_global(float(radius))
        rp, col, primary = m.contracted_rows()
        tn, td, dist = bfs_ints(rp.astype(np.int64), col.astype(np.int64), n, radius)
        if radius != -1:
            before = tn.copy()
            m.radius_correction(radius, level_from, tn, td, dist)
            corrected += int((tn != before).sum())
        m.write_global(float(radius), False, tn[primary], td[primary], dist[primary])
        assert m.columns() == r.columns()
        for c in m.columns():
            assert np.array_equal(m.attr(c), r.attr(c)), (trial, radius, c)
    assert corrected > 0 or case != "office16"  # the correction is exercised


@pytest.mark.ref
@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("kind", ["prep", "prep_pb", "vga3", "sd", "vga_l2"])
def test_loaded_map_equals_the_references_view(files, case, kind):
    if not po.have_ref():
        pytest.skip("compiled reference not present")
    d, _ = files
    src = os.path.join(d, f"{case}__{kind}.graph")
    m = capi.GraphFile(src).map()
    r = po.RefMap(graph_file=src)
    assert (m.cols, m.rows, m.spacing, m.bl_x, m.bl_y, m.n) == (r.cols, r.rows, r.spacing, r.bl_x, r.bl_y, r.n)
    for a, b in zip(m.flat_rows(), r.edges()):
        assert np.array_equal(a, b)
    cnt, dist = m.bins()
    rc, rd, rg = r.bins()
    assert np.array_equal(cnt, rc) and np.array_equal(dist, rd) and np.array_equal(m.grid_connections(), rg)
    assert m.columns() == r.columns()
    for c in m.columns():
        assert np.array_equal(m.attr(c), r.attr(c)), c
    st = np.zeros(m.cols * m.rows, np.uint16)
    po.rlib().dmxref_state(r.h, st.ctypes.data)
    assert np.array_equal(m.state(), st)


def test_damaged_files_are_rejected_not_crashed(tmp_path):
    """Random corruption (byte flips, overwritten counts, truncation) of a reference-written file: the reader either
    loads it or reports it, in a subprocess so that a crash or a runaway allocation would show as a failure."""
    import subprocess
    import sys
    from conftest import ROOT
    fx = golden("graphfiles")
    src = tmp_path / "good.graph"
    src.write_bytes(fx["oblique12__vga_l"].tobytes())
    script = f"""
import sys, resource
sys.path.insert(0, {ROOT!r})
resource.setrlimit(resource.RLIMIT_AS, (4 << 30, 4 << 30))
import numpy as np
from depthmapx_b200 import capi
good = open({str(src)!r}, 'rb').read()
rng = np.random.default_rng(7)
loaded = rejected = 0
for trial in range(300):
    b = bytearray(good)
    kind = trial % 3
    if kind == 0:
        for _ in range(int(rng.integers(1, 6))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
    elif kind == 1:
        p = int(rng.integers(0, len(b) - 4))
        b[p:p + 4] = int(rng.integers(0, 2 ** 32)).to_bytes(4, 'little')
    else:
        b = b[:int(rng.integers(8, len(b)))]
    open({str(tmp_path / 'bad.graph')!r}, 'wb').write(bytes(b))
    try:
        g = capi.GraphFile({str(tmp_path / 'bad.graph')!r})
        if g.num_maps:
            m = g.map(0)
            m.flat_rows(); m.bins(); m.columns()
            g.save({str(tmp_path / 'out.graph')!r})
        loaded += 1
    except (RuntimeError, MemoryError):
        rejected += 1
print('OK', loaded, rejected)
"""
    r = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout[-500:] + r.stderr[-2000:]
    loaded, rejected = [int(x) for x in r.stdout.split()[1:3]]
    assert loaded > 0 and rejected > 0
