"""Drop-in check on the GPU box: the REAL depthmapXcli with libvga_b200.so behind PointMap::sparkGraph2,
VGAVisualGlobal::run and VGAVisualLocal::run (oracle/_ref/depthmapXcli_gpu, built by integration/Makefile from
the unmodified reference sources + the shim translation units in integration/) must write byte-identical
.graph files to the unmodified reference CLI (oracle/_ref/depthmapXcli_ref) for
    -m VISPREP -pg .. -pp .. -pm      and      -m VGA -vm visibility -vg -vl -vr {n,3}
-- the reference's own regression method (RegressionTest/depthmaprunner.py:19-79: byte-diff of the outputs)."""
import os
import subprocess

import pytest

from conftest import ROOT
from depthmapx_b200 import capi, plans

pytestmark = pytest.mark.gpu

REF = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_ref")
GPU = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_gpu")


def run(binary, args, cwd):
    r = subprocess.run([binary] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, f"{os.path.basename(binary)} {' '.join(args)}\n{r.stdout}\n{r.stderr}"
    return r.stdout


def same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("name,grid,seed", [("oblique:20:20:11", "1", "1,1"), ("office:24:24:2", "1", "1,1"),
                                            ("oblique:16:16:5:0.7", "0.7", "0.7,0.7")])
def test_cli_outputs_byte_identical(tmp_path, name, grid, seed):
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("integration binaries not built (make -C integration)")
    if capi.device_count() < 1:
        pytest.fail("needs a CUDA device")
    plan = plans.by_name(name)
    d = str(tmp_path)
    open(os.path.join(d, "walls.csv"), "w").write(plan.csv())
    run(REF, ["-m", "IMPORT", "-f", "walls.csv", "-o", "plan.graph", "-it", "drawing"], d)
    for tag, binary in (("ref", REF), ("gpu", GPU)):
        run(binary, ["-m", "VISPREP", "-f", "plan.graph", "-o", f"prep_{tag}.graph", "-pg", grid, "-pp", seed, "-pm"], d)
    assert same(os.path.join(d, "prep_ref.graph"), os.path.join(d, "prep_gpu.graph")), "makegraph .graph differs"
    for radius in ("n", "3"):
        for tag, binary in (("ref", REF), ("gpu", GPU)):
            run(binary, ["-m", "VGA", "-f", "prep_ref.graph", "-o", f"vga_{radius}_{tag}.graph", "-vm", "visibility",
                         "-vg", "-vl", "-vr", radius], d)
        assert same(os.path.join(d, f"vga_{radius}_ref.graph"), os.path.join(d, f"vga_{radius}_gpu.graph")), \
            f"VGA -vr {radius} .graph differs"
    # row f4: metric (radius n and 6 cells) and angular VGA
    for tag2, args in (("metric_n", ["-vm", "metric", "-vr", "n"]), ("metric_r", ["-vm", "metric", "-vr", str(6 * float(grid))]),
                       ("angular", ["-vm", "angular"])):
        for tag, binary in (("ref", REF), ("gpu", GPU)):
            run(binary, ["-m", "VGA", "-f", "prep_ref.graph", "-o", f"{tag2}_{tag}.graph"] + args, d)
        assert same(os.path.join(d, f"{tag2}_ref.graph"), os.path.join(d, f"{tag2}_gpu.graph")), f"VGA {tag2} .graph differs"
    # boundary graph (-pb) and restricted visibility (-pr) variants of VISPREP
    for extra, tag2 in ((["-pb"], "pb"), (["-pr", "7.5"], "pr")):
        for tag, binary in (("ref", REF), ("gpu", GPU)):
            run(binary, ["-m", "VISPREP", "-f", "plan.graph", "-o", f"prep_{tag2}_{tag}.graph", "-pg", grid, "-pp", seed,
                         "-pm"] + extra, d)
        assert same(os.path.join(d, f"prep_{tag2}_ref.graph"), os.path.join(d, f"prep_{tag2}_gpu.graph")), tag2


# ---- the reference's own regression cases for this path (RegressionTest/regressionconfig.json) ---------------------
# Inputs: tests/golden/regression/ (copied from the reference's testdata by tests/golden/make_regression_fixtures.py).
# Expected output = what the unmodified reference CLI writes in the same run (byte-diff, same day: the file holds a date).

REGRESSION = os.path.join(ROOT, "tests", "golden", "regression")
REGRESSION_CASES = {
    "pointmap_create_fill_make_one_operation": ("gallery_empty.graph", ["-m", "VISPREP", "-pg", "0.04", "-pp", "1.32,7.24",
                                                                        "-pp", "4.88,5.24", "-pm"]),
    "dense_pointmap_create_fill_make": ("rect1x1.graph", ["-m", "VISPREP", "-pg", "0.02", "-pp", "0.5,0.5"]),
    "dense_pointmap_create_fill_make_graph": ("rect1x1.graph", ["-m", "VISPREP", "-pg", "0.02", "-pp", "0.5,0.5", "-pm"]),
    "visibility_global_n": ("gallery_connected.graph", ["-m", "VGA", "-vm", "visibility", "-vg", "-vr", "n"]),
    "visibility_global_3": ("gallery_connected.graph", ["-m", "VGA", "-vm", "visibility", "-vg", "-vr", "3"]),
    "visibility_local": ("gallery_connected.graph", ["-m", "VGA", "-vm", "visibility", "-vl"]),
    "visibility_global_local_simple": ("gallery_connected.graph", ["-m", "VGA", "-vm", "visibility", "-vg", "-vl", "-vr", "n", "-s"]),
    "vga_visual_step_depth": ("gallery_connected.graph", ["-m", "STEPDEPTH", "-sdp", "3,5", "-sdt", "visual"]),
    # row f4: metric / angular VGA (both inputs carry merge links)
    "vga_metric": ("turns_connected.graph", ["-m", "VGA", "-vm", "metric", "-vr", "n"]),
    "vga_angular": ("turns_connected.graph", ["-m", "VGA", "-vm", "angular"]),
    "vga_metric_only_map": ("gallery_connected.graph", ["-m", "VGA", "-vm", "metric", "-vr", "n"]),
    "vga_angular_only_map": ("gallery_connected.graph", ["-m", "VGA", "-vm", "angular"]),
    "vga_metric_radius": ("turns_connected.graph", ["-m", "VGA", "-vm", "metric", "-vr", "0.6"]),
}


def regression_input(name, d):
    import gzip
    import shutil
    dst = os.path.join(d, name)
    if os.path.exists(os.path.join(REGRESSION, name)):
        shutil.copyfile(os.path.join(REGRESSION, name), dst)
    else:
        with gzip.open(os.path.join(REGRESSION, name + ".gz"), "rb") as f, open(dst, "wb") as o:
            shutil.copyfileobj(f, o)
    return dst


@pytest.mark.parametrize("case", sorted(REGRESSION_CASES))
def test_reference_regression_case(tmp_path, case):
    if not (os.path.exists(REF) and os.path.exists(GPU)):
        pytest.skip("integration binaries not built (make -C integration)")
    if capi.device_count() < 1:
        pytest.fail("needs a CUDA device")
    infile, args = REGRESSION_CASES[case]
    d = str(tmp_path)
    regression_input(infile, d)
    for tag, binary in (("ref", REF), ("gpu", GPU)):
        run(binary, args + ["-f", infile, "-o", f"out_{tag}.graph"], d)
    assert same(os.path.join(d, "out_ref.graph"), os.path.join(d, "out_gpu.graph")), f"{case}: .graph differs"
