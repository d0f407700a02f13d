"""CPU tests of the host layer and of the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden
from depthmapx_b200 import capi, plans


@pytest.mark.parametrize("name", GOLDEN)
def test_host_prep_matches_reference(name):
    """setGrid + blockLines + makePoints of dmx::PointMap == the reference's (golden fixture)."""
    fx = golden(name)
    m = capi.HostMap(fx["walls"], float(fx["spacing"]))
    assert (m.cols, m.rows) == (int(fx["cols"]), int(fx["rows"]))
    assert m.bl_x == float(fx["bl_x"]) and m.bl_y == float(fx["bl_y"])
    for s in fx["seeds"]:
        assert m.fill(float(s[0]), float(s[1]))
    f = m.flat()
    assert np.array_equal(f.state, fx["state"])
    assert np.array_equal(f.line_off, fx["line_off"])
    assert np.array_equal(f.lines, fx["lines"].reshape(-1, 5))


def test_setgrid_quirks():
    """cols/rows/origin rules of PointMap::setGrid (salaTest/testpointmap.cpp:94-311 style cases)."""
    # region (0,0)-(1.5,1.5), spacing 0.5: offset fmod(0,0.5)=0 -> +0.5 ... centre grid on multiples of 0.5
    m = capi.HostMap([[0, 0, 0, 1.5], [0, 1.5, 1.5, 1.5], [1.5, 1.5, 1.5, 0], [1.5, 0, 0, 0]], 0.5)
    assert (m.cols, m.rows) == (4, 4)
    assert (m.bl_x, m.bl_y) == (0.0, 0.0)
    m = capi.HostMap([[0.5, 0.5, 100.5, 0.5], [100.5, 0.5, 100.5, 100.5]], 1.0)
    assert (m.cols, m.rows) == (102, 102)
    assert (m.bl_x, m.bl_y) == (0.0, 0.0)


def test_fill_rejects_outside_and_refill():
    p = plans.room(10, 10, 0)
    m = capi.HostMap(p.walls, 1.0)
    assert not m.fill(-50.0, -50.0)
    assert m.fill(1.0, 1.0)
    assert not m.fill(1.0, 1.0)  # already filled
    assert m.n == 100


def test_plans_deterministic_and_sized():
    a, b = plans.by_name("C2"), plans.by_name("C2")
    assert a.walls == b.walls
    f = capi.prepare(plans.office(64, 64, 1))
    assert f.n_filled == 64 * 64
    for w in plans.by_name("C5").walls:
        for v in w:
            assert (v * 2) % 2 == 1  # half-integer coordinates: no wall through a cell centre


def test_abi_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "vga_b200.h")).read()
    declared = set(re.findall(r"\b(vga_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"vga_progress_fn", "vga_cancel_fn"}
    lib = C.CDLL(os.path.join(capi.LIBDIR, "libvga_b200.so"))
    assert declared, "no declarations parsed"
    for s in sorted(declared):
        assert hasattr(lib, s), s
    assert declared == set(capi.ABI_SYMBOLS)
    hh = open(os.path.join(ROOT, "include", "vga_host.h")).read()
    hdecl = set(re.findall(r"\b(dmxh_[a-z_0-9]+)\s*\(", hh))
    hl = C.CDLL(os.path.join(capi.LIBDIR, "libvga_host.so"))
    for s in sorted(hdecl):
        assert hasattr(hl, s), s
    assert hdecl == set(capi.HOST_SYMBOLS)


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every compute entry point must fail loudly."""
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.VgaError) as e:
        capi.Context(0)
    assert e.value.code == -2
    m = capi.HostMap(plans.room(12, 12, 0).walls, 1.0)
    m.fill(1.0, 1.0)
    with pytest.raises(RuntimeError):
        m.make_graph()


def test_product_does_not_reference_oracle():
    """The product package must never import, link or execute anything under oracle/."""
    bad = []
    for dp, _, files in os.walk(os.path.join(ROOT, "depthmapx_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|pyoracle|libvgaoracle|libdmxref|vga_oracle\.h", txt):
                    bad.append(fn)
    assert not bad, bad


def test_formula_stage_matches_golden():
    """vga_global_attributes / vga_local_attributes are host code: check them on the CPU against the
    reference's columns, feeding integers recomputed from the fixture's adjacency."""
    from oracle import pyoracle as po
    fx = golden("office24")
    g = po.Grid(int(fx["cols"]), int(fx["rows"]), float(fx["spacing"]), float(fx["bl_x"]), float(fx["bl_y"]),
                fx["state"], fx["line_off"], fx["lines"])
    og = po.OracleGraph(g, edges=(fx["rowptr"], fx["ref"]))
    cols = [str(c) for c in fx["columns"]]
    for radius in (-1, 3):
        tn, td, dist, nl = og.global_ints(radius)
        out = capi.global_attributes(tn, td, dist)
        sfx = "" if radius == -1 else f" R{radius}"
        for k, v in out.items():
            assert np.array_equal(v, fx[f"attr_{cols.index(k + sfx)}"]), k
    out = capi.local_attributes(*og.local_ints())
    for k, v in out.items():
        assert np.array_equal(v, fx[f"attr_{cols.index(k)}"]), k
