"""GPU: the .graph pipelines of tests/test_graphfile.py with libvga_b200 doing the work (SURVEY §8 rows f1 + f2): the host
layer reads the reference CLI's files, builds / uploads the graph on the B200, runs the analyses and writes .graph files
that must be byte-identical to what the unmodified reference CLI wrote (tests/golden/graphfiles.npz).

The codec and the attribute stages are pinned on CPU (tests/test_graphfile.py); what is new on the GPU side is
PointMap::nodes() (Node encoding from the device rows), PointMap::ensureGraph() (upload of a loaded adjacency with bins)
and the host-level VGAVisualGlobalDepth.  Green on a B200 since round 2 (all cases; the first-run xfail guard of round 1
is gone); the CPU suite runs the same module against the SIMT emulation (tests/test_emulated_kernels.py)."""
import os

import numpy as np
import pytest

from conftest import golden
from depthmapx_b200 import capi

# a kernel that never returns would hang the GPU box: give up on the whole run instead (these modules run last)
pytestmark = [pytest.mark.timeout(900, method="thread"), pytest.mark.gpu]

CASES = ["oblique12", "oblique10s07", "office16"]


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


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("boundary", [False, True])
def test_visprep_from_drawing(files, tmp_path, case, boundary):
    """-m VISPREP -pg -pp -pm [-pb]: plan.graph -> grid, fill, makegraph on the GPU -> .graph"""
    d, args = files
    _spec, grid, seed, _sdp = args[case][:4]
    g = capi.GraphFile(os.path.join(d, f"{case}__plan.graph"))
    m = g.new_map(float(grid))
    assert m.fill(*[float(x) for x in seed.split(",")])
    assert g.make_graph(boundary)
    out = str(tmp_path / "o.graph")
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__prep_pb.graph" if boundary else f"{case}__prep.graph"))


@pytest.mark.parametrize("case", CASES)
def test_vga_on_loaded_graph(files, tmp_path, case):
    """-m VGA -vm visibility -vg -vl -vr n, then -vg -vr 3 on the result, each from a file loaded by the host layer"""
    d, _ = files
    out = str(tmp_path / "o.graph")
    g = capi.GraphFile(os.path.join(d, f"{case}__prep.graph"))
    m = g.map()
    assert m.vga_local() and m.vga_global(-1.0)
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga.graph"))
    g = capi.GraphFile(os.path.join(d, f"{case}__vga.graph"))
    assert g.map().vga_global(3.0)
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga3.graph"))


@pytest.mark.parametrize("case", CASES)
def test_metric_and_angular_on_loaded_graph(files, tmp_path, case):
    """Row f4 through the host layer (dmx::VGAMetric / dmx::VGAAngular): -m VGA -vm metric -vr n | -vr <4 cells>, -vm angular
    (and a second angular run on its own output: getOrInsertColumn keeps the columns), and both on the map with merge links --
    files byte-identical to the unmodified reference CLI's (tests/golden/graphfiles_metric.npz)."""
    d, _ = files
    fx = golden("graphfiles_metric")
    out = str(tmp_path / "o.graph")

    def check(src, key, fn):
        if isinstance(src, bytes):
            open(str(tmp_path / "in.graph"), "wb").write(src)
            src = str(tmp_path / "in.graph")
        g = capi.GraphFile(src)
        assert fn(g.map())
        g.save(out)
        assert data(out) == fx[f"{case}__{key}"].tobytes(), key

    prep, prep_l = os.path.join(d, f"{case}__prep.graph"), os.path.join(d, f"{case}__prep_l.graph")
    check(prep, "metric", lambda m: m.vga_metric(-1.0))
    check(prep, "metric_r", lambda m: m.vga_metric(float(str(fx[f"{case}__radius"]))))
    check(prep, "angular", lambda m: m.vga_angular(-1.0))
    check(fx[f"{case}__angular"].tobytes(), "angular2", lambda m: m.vga_angular(-1.0))
    check(prep_l, "metric_l", lambda m: m.vga_metric(-1.0))
    check(prep_l, "angular_l", lambda m: m.vga_angular(-1.0))


@pytest.mark.parametrize("case", CASES)
def test_step_depth_on_loaded_graph(files, tmp_path, case):
    """-m STEPDEPTH -sdp x,y -sdt visual"""
    d, args = files
    sdp = [float(x) for x in args[case][3].split(",")]
    g = capi.GraphFile(os.path.join(d, f"{case}__prep.graph"))
    assert g.map().step_depth([sdp])
    out = str(tmp_path / "o.graph")
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__sd.graph"))


@pytest.mark.parametrize("case", CASES)
def test_merge_links_on_the_gpu(files, tmp_path, case):
    """VGA global (radius n and 2) + local and step depth on a map with merge links."""
    d, args = files
    out = str(tmp_path / "o.graph")
    g = capi.GraphFile(os.path.join(d, f"{case}__prep_l.graph"))
    m = g.map()
    assert m.vga_local() and m.vga_global(-1.0)
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga_l.graph"))
    g = capi.GraphFile(os.path.join(d, f"{case}__prep_l.graph"))
    assert g.map().vga_global(2.0)
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__vga_l2.graph"))
    g = capi.GraphFile(os.path.join(d, f"{case}__prep_l.graph"))
    assert g.map().step_depth([[float(x) for x in args[case][3].split(",")]])
    g.save(out)
    assert data(out) == data(os.path.join(d, f"{case}__sd_l.graph"))


def test_loaded_and_built_graphs_agree(files):
    """The adjacency uploaded from a file and the one built on the GPU from the same plan give the same integers."""
    d, args = files
    _spec, grid, seed, _sdp = args["office16"][:4]
    g1 = capi.GraphFile(os.path.join(d, "office16__plan.graph"))
    m1 = g1.new_map(float(grid))
    m1.fill(*[float(x) for x in seed.split(",")])
    g1.make_graph(False)
    g2 = capi.GraphFile(os.path.join(d, "office16__prep.graph"))
    m2 = g2.map()
    for a, b in zip(m1.flat_rows(), m2.flat_rows()):
        assert np.array_equal(a, b)
    for m in (m1, m2):
        m.vga_local()
        m.vga_global(-1.0)
    assert m1.columns() == m2.columns()
    for c in m1.columns():
        assert np.array_equal(m1.attr(c), m2.attr(c)), c


@pytest.mark.parametrize("case", CASES)
def test_cli_stepdepth_shim(files, tmp_path, case):
    """The real depthmapXcli with integration/vga_depth_gpu.cpp behind VGAVisualGlobalDepth::run:
    -m STEPDEPTH -sdt visual must write the reference's bytes."""
    import subprocess
    from conftest import ROOT
    gpu_cli = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_gpu")
    if not os.path.exists(gpu_cli):
        pytest.skip("integration binaries not built (make -C integration)")
    d, args = files
    out = str(tmp_path / "sd.graph")
    r = subprocess.run([gpu_cli, "-m", "STEPDEPTH", "-f", os.path.join(d, f"{case}__prep.graph"), "-o", out, "-sdp", args[case][3],
                        "-sdt", "visual"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert data(out) == data(os.path.join(d, f"{case}__sd.graph"))


@pytest.mark.parametrize("case", CASES)
def test_cli_shims_on_linked_graph(files, tmp_path, case):
    """The real depthmapXcli with the GPU shims on a map with merge links (-m LINK): VGA global radius n + local, global
    radius 2 (at-the-radius correction) and visual step depth must write the reference's bytes."""
    import subprocess
    from conftest import ROOT
    gpu_cli = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_gpu")
    if not os.path.exists(gpu_cli):
        pytest.skip("integration binaries not built (make -C integration)")
    d, args = files
    src = os.path.join(d, f"{case}__prep_l.graph")
    out = str(tmp_path / "o.graph")
    for extra, want in ((["-m", "VGA", "-vm", "visibility", "-vg", "-vl", "-vr", "n"], "vga_l"),
                        (["-m", "VGA", "-vm", "visibility", "-vg", "-vr", "2"], "vga_l2"),
                        (["-m", "STEPDEPTH", "-sdp", args[case][3], "-sdt", "visual"], "sd_l")):
        r = subprocess.run([gpu_cli, "-f", src, "-o", out] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        assert data(out) == data(os.path.join(d, f"{case}__{want}.graph")), want


@pytest.mark.parametrize("case", ["oblique12", "office16"])
def test_context_filled_map_host_layer_and_cli(files, tmp_path, case):
    """Semi-filled (FILLED | CONTEXTFILLED) maps: cells that are not "even" are skipped as sources and, under a radius,
    counted but not expanded (vgavisualglobal.cpp:75, 108-110; vgavisuallocal.cpp:43; vgavisualglobaldepth.cpp:53).  The CLI
    cannot semi-fill (GUI only), so the filled map is written by the host layer and then taken through makegraph, VGA
    (radius n and 3, local) and step depth by the unmodified reference CLI, by the CLI with the GPU shims and by the host
    layer: all three must write the same bytes."""
    import subprocess
    from conftest import ROOT
    ref_cli = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_ref")
    gpu_cli = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_gpu")
    if not (os.path.exists(ref_cli) and os.path.exists(gpu_cli)):
        pytest.skip("integration binaries not built (make -C integration)")
    d, args = files
    _spec, grid, seed, sdp = args[case][:4]
    t = str(tmp_path)

    def run(binary, a):
        r = subprocess.run([binary] + a, cwd=t, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, " ".join(a) + "\n" + r.stdout + r.stderr

    g = capi.GraphFile(os.path.join(d, f"{case}__plan.graph"))
    m = g.new_map(float(grid))
    assert m.fill(*[float(x) for x in seed.split(",")], fill_type=1)
    assert m.context_skip().sum() > 0
    g.save(os.path.join(t, "ctxfill.graph"))
    for tag, binary in (("ref", ref_cli), ("gpu", gpu_cli)):
        run(binary, ["-m", "VISPREP", "-f", "ctxfill.graph", "-o", f"prep_{tag}.graph", "-pm"])
        run(binary, ["-m", "VGA", "-f", f"prep_{tag}.graph", "-o", f"vga_{tag}.graph", "-vm", "visibility", "-vg", "-vl", "-vr", "n"])
        run(binary, ["-m", "VGA", "-f", f"vga_{tag}.graph", "-o", f"vga3_{tag}.graph", "-vm", "visibility", "-vg", "-vr", "3"])
        run(binary, ["-m", "STEPDEPTH", "-f", f"prep_{tag}.graph", "-o", f"sd_{tag}.graph", "-sdp", sdp, "-sdt", "visual"])
    for name in ("prep", "vga", "vga3", "sd"):
        assert data(os.path.join(t, f"{name}_ref.graph")) == data(os.path.join(t, f"{name}_gpu.graph")), f"CLI shim: {name}"
    # host layer
    g.make_graph()
    g.save(os.path.join(t, "prep_host.graph"))
    assert data(os.path.join(t, "prep_host.graph")) == data(os.path.join(t, "prep_ref.graph"))
    assert m.vga_local() and m.vga_global(-1.0)
    g.save(os.path.join(t, "vga_host.graph"))
    assert data(os.path.join(t, "vga_host.graph")) == data(os.path.join(t, "vga_ref.graph"))
    g3 = capi.GraphFile(os.path.join(t, "vga_ref.graph"))
    assert g3.map().vga_global(3.0)
    g3.save(os.path.join(t, "vga3_host.graph"))
    assert data(os.path.join(t, "vga3_host.graph")) == data(os.path.join(t, "vga3_ref.graph"))
    g4 = capi.GraphFile(os.path.join(t, "prep_ref.graph"))
    assert g4.map().step_depth([[float(x) for x in sdp.split(",")]])
    g4.save(os.path.join(t, "sd_host.graph"))
    assert data(os.path.join(t, "sd_host.graph")) == data(os.path.join(t, "sd_ref.graph"))
