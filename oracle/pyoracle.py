"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for oracle/_ref/libvgaoracle.so (our C restatement)
and oracle/_ref/libdmxref.so (the unmodified reference + oracle/ref_harness.cpp).

May be imported only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Never from depthmapx_b200/.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REFDIR = os.path.join(HERE, "_ref")
REFERENCE_ROOT = "/root/reference"

c_dp = C.POINTER(C.c_double)


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def build(ref: bool = True) -> None:
    """Compile the C restatement and, when /root/reference is present, the reference library."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref and os.path.isdir(REFERENCE_ROOT):
        subprocess.check_call(["make", "-s", "-j8", "-C", HERE, "ref"])


def have_ref() -> bool:
    return os.path.exists(os.path.join(REFDIR, "libdmxref.so"))


def have_oracle() -> bool:
    return os.path.exists(os.path.join(REFDIR, "libvgaoracle.so"))


# ----------------------------------------------------------------------------------------- oracle

class VgaoGrid(C.Structure):
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("spacing", C.c_double), ("bl_x", C.c_double),
                ("bl_y", C.c_double), ("maxdist", C.c_double), ("state", C.POINTER(C.c_uint16)),
                ("line_off", C.POINTER(C.c_uint32)), ("lines", c_dp)]


@dataclass
class Grid:
    """Flat grid inputs of the hot path (the vga_grid of include/vga_b200.h)."""
    cols: int
    rows: int
    spacing: float
    bl_x: float
    bl_y: float
    state: np.ndarray      # uint16 [cols*rows], x-major
    line_off: np.ndarray   # uint32 [cols*rows+1]
    lines: np.ndarray      # float64 [nseg,5]
    maxdist: float = -1.0

    def c(self) -> VgaoGrid:
        self.state = np.ascontiguousarray(self.state, np.uint16)
        self.line_off = np.ascontiguousarray(self.line_off, np.uint32)
        self.lines = np.ascontiguousarray(self.lines, np.float64).reshape(-1, 5)
        if self.lines.shape[0] == 0:
            self._dummy = np.zeros((1, 5))
            lp = _ptr(self._dummy, C.c_double)
        else:
            lp = _ptr(self.lines, C.c_double)
        return VgaoGrid(self.cols, self.rows, self.spacing, self.bl_x, self.bl_y, self.maxdist,
                        _ptr(self.state, C.c_uint16), _ptr(self.line_off, C.c_uint32), lp)

    @property
    def n_filled(self) -> int:
        return int(((self.state & 2) != 0).sum())


_olib = None


def olib():
    global _olib
    if _olib is None:
        L = C.CDLL(os.path.join(REFDIR, "libvgaoracle.so"))
        L.vgao_makegraph.restype = C.c_void_p
        L.vgao_makegraph.argtypes = [C.POINTER(VgaoGrid)]
        L.vgao_makegraph_range.restype = C.c_void_p
        L.vgao_makegraph_range.argtypes = [C.POINTER(VgaoGrid), C.c_int64, C.c_int64]
        L.vgao_graph_from_edges.restype = C.c_void_p
        L.vgao_graph_from_edges.argtypes = [C.POINTER(VgaoGrid), C.c_void_p, C.c_void_p]
        L.vgao_graph_free.argtypes = [C.c_void_p]
        for f in ("vgao_num_cells", "vgao_num_acc", "vgao_num_iter"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.vgao_cell_refs.argtypes = [C.c_void_p, C.c_void_p]
        L.vgao_acc_rows.argtypes = [C.c_void_p] * 4
        L.vgao_iter_rows.argtypes = [C.c_void_p] * 4
        L.vgao_node_attrs.argtypes = [C.c_void_p] * 7
        L.vgao_global.restype = C.c_int
        L.vgao_global.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_int32, C.c_void_p]
        L.vgao_global_formulas.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p] + \
                                          [C.c_void_p] * 7
        L.vgao_local.restype = C.c_int
        L.vgao_local.argtypes = [C.c_void_p, C.c_int64, C.c_int64] + [C.c_void_p] * 4
        L.vgao_local_formulas.argtypes = [C.c_int64] + [C.c_void_p] * 7
        L.vgao_step_depth.restype = C.c_int
        L.vgao_step_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.vgao_metric.restype = C.c_int
        L.vgao_metric.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int64, C.c_int64] + [C.c_void_p] * 4
        L.vgao_angular.restype = C.c_int
        L.vgao_angular.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_int64, C.c_int64] + [C.c_void_p] * 3
        L.vgao_global_csr.restype = C.c_int
        L.vgao_global_csr.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_int32]
        L.vgao_local_csr.restype = C.c_int
        L.vgao_local_csr.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vgao_sieve_kat.restype = C.c_int
        L.vgao_sieve_kat.argtypes = [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _olib = L
    return _olib


def _p(a):
    return a.ctypes.data if a is not None else None


GLOBAL_COLS = ["Visual Node Count", "Visual Mean Depth", "Visual Integration [HH]", "Visual Integration [P-value]",
               "Visual Integration [Tekl]", "Visual Entropy", "Visual Relativised Entropy"]
LOCAL_COLS = ["Visual Clustering Coefficient", "Visual Control", "Visual Controllability"]


class OracleGraph:
    def __init__(self, grid: Grid, src_range=None, edges=None):
        self.grid = grid
        self._cg = grid.c()
        L = olib()
        if edges is not None:
            rowptr, ref = edges
            self._rp = np.ascontiguousarray(rowptr, np.uint64)
            self._rf = np.ascontiguousarray(ref, np.int32)
            self.h = L.vgao_graph_from_edges(C.byref(self._cg), _p(self._rp), _p(self._rf))
        elif src_range is None:
            self.h = L.vgao_makegraph(C.byref(self._cg))
        else:
            self.h = L.vgao_makegraph_range(C.byref(self._cg), src_range[0], src_range[1])
        self.n = L.vgao_num_cells(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            olib().vgao_graph_free(self.h)
            self.h = None

    def cell_refs(self):
        r = np.zeros(self.n, np.int32)
        olib().vgao_cell_refs(self.h, _p(r))
        return r

    def _rows(self, which):
        L = olib()
        ne = L.vgao_num_acc(self.h) if which == "acc" else L.vgao_num_iter(self.h)
        rowptr = np.zeros(self.n + 1, np.uint64)
        ref = np.zeros(max(ne, 1), np.int32)
        b = np.zeros(max(ne, 1), np.uint8)
        (L.vgao_acc_rows if which == "acc" else L.vgao_iter_rows)(self.h, _p(rowptr), _p(ref), _p(b))
        return rowptr, ref[:ne], b[:ne]

    def acc_rows(self):
        return self._rows("acc")

    def iter_rows(self):
        return self._rows("iter")

    def node_attrs(self):
        n = self.n
        out = dict(connectivity=np.zeros(n, np.float32), first_moment=np.zeros(n, np.float32),
                   second_moment=np.zeros(n, np.float32), far=np.zeros((n, 32), np.float32),
                   bin_count=np.zeros((n, 32), np.uint16), gridconn=np.zeros(n, np.uint8))
        olib().vgao_node_attrs(self.h, _p(out["connectivity"]), _p(out["first_moment"]), _p(out["second_moment"]),
                               _p(out["far"]), _p(out["bin_count"]), _p(out["gridconn"]))
        return out

    def global_ints(self, radius=-1, src=None, maxl=64):
        b, e = (0, self.n) if src is None else src
        k = e - b
        tn = np.zeros(k, np.int32)
        td = np.zeros(k, np.int64)
        dist = np.zeros((k, maxl), np.int32)
        nl = np.zeros(k, np.int32)
        rc = olib().vgao_global(self.h, radius, b, e, _p(tn), _p(td), _p(dist), maxl, _p(nl))
        if rc != 0:
            raise RuntimeError("oracle BFS: maxl too small")
        return tn, td, dist, nl

    def step_depth(self, sources):
        src = np.ascontiguousarray(sources, np.int32)
        d = np.zeros(self.n, np.int32)
        olib().vgao_step_depth(self.h, _p(src), len(src), _p(d))
        return d

    def metric(self, spacing, radius=-1.0, src=None, partner=None):
        """(Metric Mean Shortest-Path Angle, Metric Mean Shortest-Path Distance, Metric Mean Straight-Line Distance,
        Metric Node Count) as float32 columns of the sources src = (begin, end)."""
        b, e = (0, self.n) if src is None else src
        out = [np.zeros(e - b, np.float32) for _ in range(4)]
        mp = None if partner is None else np.ascontiguousarray(partner, np.int32)
        olib().vgao_metric(self.h, None if mp is None else _p(mp), float(spacing), float(radius), b, e, *[_p(a) for a in out])
        return tuple(out)

    def angular(self, radius=-1.0, src=None, partner=None):
        """(Angular Mean Depth, Angular Total Depth, Angular Node Count) as float32 columns."""
        b, e = (0, self.n) if src is None else src
        out = [np.zeros(e - b, np.float32) for _ in range(3)]
        mp = None if partner is None else np.ascontiguousarray(partner, np.int32)
        olib().vgao_angular(self.h, None if mp is None else _p(mp), float(radius), b, e, *[_p(a) for a in out])
        return tuple(out)

    def local_ints(self, src=None):
        b, e = (0, self.n) if src is None else src
        k = e - b
        cl = np.zeros(k, np.int64)
        kk = np.zeros(k, np.int32)
        tot = np.zeros(k, np.int32)
        ctl = np.zeros(k, np.float32)
        olib().vgao_local(self.h, b, e, _p(cl), _p(kk), _p(tot), _p(ctl))
        return cl, kk, tot, ctl


def _threads():
    return max(1, min(32, os.cpu_count() or 1))


def global_csr(n, rowptr, col, sources, radius=-1, maxl=64, shift=0):
    """vgao_global over an ordinal CSR (rowptr uint64 [n+1], col uint32 [E], entry = col >> shift) for the listed
    sources; the sources are split over host threads (the C call releases the GIL).  No copy of col is made."""
    from concurrent.futures import ThreadPoolExecutor
    rowptr = np.ascontiguousarray(rowptr, np.uint64)
    assert col.dtype == np.uint32 and col.flags.c_contiguous
    src = np.ascontiguousarray(sources, np.int64)
    k = len(src)
    tn = np.zeros(k, np.int32)
    td = np.zeros(k, np.int64)
    dist = np.zeros((k, maxl), np.int32)
    L = olib()

    def part(lo, hi):
        return L.vgao_global_csr(n, _p(rowptr), _p(col), shift, radius, src[lo:hi].ctypes.data, hi - lo,
                                 tn[lo:hi].ctypes.data, td[lo:hi].ctypes.data, dist[lo:hi].ctypes.data, maxl)
    nt = min(_threads(), k) or 1
    cuts = [k * i // nt for i in range(nt + 1)]
    with ThreadPoolExecutor(nt) as ex:
        rcs = list(ex.map(lambda ab: part(*ab), zip(cuts[:-1], cuts[1:])))
    if any(rcs):
        raise RuntimeError("oracle BFS: maxl too small")
    return tn, td, dist


def local_csr(n, nv, rowptr, col, refs, cells, shift=0):
    """vgao_local over an ordinal CSR for the listed cells (rows sorted or not); refs = packed PixelRef of all nv
    vertices (cells then ghosts).  Split over host threads."""
    from concurrent.futures import ThreadPoolExecutor
    rowptr = np.ascontiguousarray(rowptr, np.uint64)
    assert col.dtype == np.uint32 and col.flags.c_contiguous
    refs = np.ascontiguousarray(refs, np.int32)
    cs = np.ascontiguousarray(cells, np.int64)
    k = len(cs)
    cl = np.zeros(k, np.int64)
    kk = np.zeros(k, np.int32)
    tot = np.zeros(k, np.int32)
    ctl = np.zeros(k, np.float32)
    L = olib()

    def part(lo, hi):
        return L.vgao_local_csr(n, nv, _p(rowptr), _p(col), shift, _p(refs), cs[lo:hi].ctypes.data, hi - lo,
                                cl[lo:hi].ctypes.data, kk[lo:hi].ctypes.data, tot[lo:hi].ctypes.data, ctl[lo:hi].ctypes.data)
    nt = min(_threads(), k) or 1
    cuts = [k * i // nt for i in range(nt + 1)]
    with ThreadPoolExecutor(nt) as ex:
        list(ex.map(lambda ab: part(*ab), zip(cuts[:-1], cuts[1:])))
    return cl, kk, tot, ctl


def global_formulas(tn, td, dist, nl):
    n = len(tn)
    dist = np.ascontiguousarray(dist, np.int32)
    outs = [np.zeros(n, np.float32) for _ in range(7)]
    olib().vgao_global_formulas(n, _p(np.ascontiguousarray(tn, np.int32)), _p(np.ascontiguousarray(td, np.int64)),
                                _p(dist), dist.shape[1], _p(np.ascontiguousarray(nl, np.int32)), *[_p(o) for o in outs])
    return dict(zip(GLOBAL_COLS, outs))


def local_formulas(cl, kk, tot, ctl):
    n = len(cl)
    outs = [np.zeros(n, np.float32) for _ in range(3)]
    olib().vgao_local_formulas(n, _p(np.ascontiguousarray(cl, np.int64)), _p(np.ascontiguousarray(kk, np.int32)),
                               _p(np.ascontiguousarray(tot, np.int32)), _p(np.ascontiguousarray(ctl, np.float32)),
                               *[_p(o) for o in outs])
    return dict(zip(LOCAL_COLS, outs))


def sieve_kat(cx, cy, q, segs):
    segs = np.ascontiguousarray(segs, np.float64).reshape(-1, 4)
    gaps = np.zeros((16, 2))
    n = olib().vgao_sieve_kat(cx, cy, q, _p(segs), segs.shape[0], _p(gaps), 16)
    return gaps[:n]


# ----------------------------------------------------------------------------------------- reference

_rlib = None


def rlib():
    global _rlib
    if _rlib is None:
        L = C.CDLL(os.path.join(REFDIR, "libdmxref.so"))
        L.dmxref_create.restype = C.c_void_p
        L.dmxref_create.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.dmxref_destroy.argtypes = [C.c_void_p]
        L.dmxref_grid.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.dmxref_fill.restype = C.c_int
        L.dmxref_fill.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.dmxref_block_lines.argtypes = [C.c_void_p]
        L.dmxref_filled_count.argtypes = [C.c_void_p]
        L.dmxref_state.argtypes = [C.c_void_p, C.c_void_p]
        L.dmxref_cell_lines.restype = C.c_int64
        L.dmxref_cell_lines.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.dmxref_makegraph.restype = C.c_double
        L.dmxref_makegraph.argtypes = [C.c_void_p, C.c_int, C.c_double]
        L.dmxref_edges.restype = C.c_int64
        L.dmxref_edges.argtypes = [C.c_void_p] * 4
        L.dmxref_bins.argtypes = [C.c_void_p] * 4
        L.dmxref_attr.restype = C.c_int
        L.dmxref_attr.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.dmxref_columns.restype = C.c_int
        L.dmxref_columns.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        L.dmxref_vga_global.restype = C.c_double
        L.dmxref_vga_global.argtypes = [C.c_void_p, C.c_double, C.c_int]
        L.dmxref_vga_local.restype = C.c_double
        L.dmxref_vga_local.argtypes = [C.c_void_p, C.c_int]
        L.dmxref_vga_metric.restype = C.c_double
        L.dmxref_vga_metric.argtypes = [C.c_void_p, C.c_double]
        L.dmxref_vga_angular.restype = C.c_double
        L.dmxref_vga_angular.argtypes = [C.c_void_p, C.c_double]
        L.dmxref_step_depth.restype = C.c_double
        L.dmxref_step_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.dmxref_sample_makegraph.restype = C.c_double
        L.dmxref_sample_makegraph.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_void_p]
        L.dmxref_sample_global.restype = C.c_double
        L.dmxref_sample_global.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.dmxref_merge.argtypes = [C.c_void_p] + [C.c_double] * 4
        L.dmxref_fill_type.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int]
        L.dmxref_graph_open.restype = C.c_void_p
        L.dmxref_graph_open.argtypes = [C.c_char_p]
        L.dmxref_graph_save.argtypes = [C.c_void_p, C.c_char_p]
        L.dmxref_graph_rewrite.argtypes = [C.c_char_p, C.c_char_p]
        _rlib = L
    return _rlib


def ref_graph_rewrite(src, dst) -> bool:
    """MetaGraph::readFromFile + MetaGraph::write by the reference, nothing in between."""
    return bool(rlib().dmxref_graph_rewrite(os.fsencode(src), os.fsencode(dst)))


class RefMap:
    """The reference's PointMap driven through oracle/ref_harness.cpp."""

    def __init__(self, walls=None, spacing=1.0, graph_file=None):
        """walls + spacing: a new map; graph_file: the displayed point map of a .graph read by MetaGraph."""
        if graph_file is not None:
            self.h = rlib().dmxref_graph_open(os.fsencode(graph_file))
            if not self.h:
                raise RuntimeError(f"the reference could not read {graph_file}")
        else:
            w = np.ascontiguousarray(walls, np.float64).reshape(-1, 4)
            self.h = rlib().dmxref_create(_p(w), w.shape[0], spacing)
        c, r = C.c_int(), C.c_int()
        s, bx, by = C.c_double(), C.c_double(), C.c_double()
        rlib().dmxref_grid(self.h, C.byref(c), C.byref(r), C.byref(s), C.byref(bx), C.byref(by))
        self.cols, self.rows, self.spacing, self.bl_x, self.bl_y = c.value, r.value, s.value, bx.value, by.value

    def __del__(self):
        if getattr(self, "h", None):
            rlib().dmxref_destroy(self.h)
            self.h = None

    def fill(self, x, y, fill_type=0):
        if fill_type:
            return bool(rlib().dmxref_fill_type(self.h, x, y, fill_type))
        return bool(rlib().dmxref_fill(self.h, x, y))

    def save(self, path):
        return bool(rlib().dmxref_graph_save(self.h, os.fsencode(path)))

    def merge(self, ax, ay, bx, by):
        return bool(rlib().dmxref_merge(self.h, ax, ay, bx, by))

    def block_lines(self):
        rlib().dmxref_block_lines(self.h)

    @property
    def n(self):
        return rlib().dmxref_filled_count(self.h)

    def grid(self, maxdist=-1.0) -> Grid:
        """Flat hot-path inputs as the reference prepared them (call after fill, before makegraph)."""
        cells = self.cols * self.rows
        state = np.zeros(cells, np.uint16)
        rlib().dmxref_state(self.h, _p(state))
        off = np.zeros(cells + 1, np.uint32)
        nseg = rlib().dmxref_cell_lines(self.h, _p(off), None)
        lines = np.zeros((max(nseg, 1), 5))
        rlib().dmxref_cell_lines(self.h, _p(off), _p(lines))
        return Grid(self.cols, self.rows, self.spacing, self.bl_x, self.bl_y, state, off, lines[:nseg], maxdist)

    def makegraph(self, boundary=False, maxdist=-1.0):
        t = rlib().dmxref_makegraph(self.h, int(boundary), maxdist)
        if t < 0:
            raise RuntimeError("reference sparkGraph2 failed")
        return t

    def edges(self):
        n = self.n
        ne = rlib().dmxref_edges(self.h, None, None, None)
        rowptr = np.zeros(n + 1, np.uint64)
        ref = np.zeros(max(ne, 1), np.int32)
        b = np.zeros(max(ne, 1), np.uint8)
        rlib().dmxref_edges(self.h, _p(rowptr), _p(ref), _p(b))
        return rowptr, ref[:ne], b[:ne]

    def bins(self):
        n = self.n
        cnt = np.zeros((n, 32), np.uint16)
        dist = np.zeros((n, 32), np.float32)
        gc = np.zeros(n, np.uint8)
        rlib().dmxref_bins(self.h, _p(cnt), _p(dist), _p(gc))
        return cnt, dist, gc

    def attr(self, name):
        out = np.zeros(self.n, np.float32)
        if not rlib().dmxref_attr(self.h, name.encode(), _p(out)):
            raise KeyError(name)
        return out

    def columns(self):
        buf = C.create_string_buffer(8192)
        rlib().dmxref_columns(self.h, buf, 8192)
        return [s for s in buf.value.decode().split("\n") if s]

    def vga_global(self, radius=-1.0, simple=False):
        return rlib().dmxref_vga_global(self.h, radius, int(simple))

    def vga_local(self, simple=False):
        return rlib().dmxref_vga_local(self.h, int(simple))

    def vga_metric(self, radius=-1.0):
        return rlib().dmxref_vga_metric(self.h, radius)

    def vga_angular(self, radius=-1.0):
        return rlib().dmxref_vga_angular(self.h, radius)

    def step_depth(self, sources):
        src = np.ascontiguousarray(sources, np.int32)
        t = rlib().dmxref_step_depth(self.h, _p(src), len(src))
        if t < 0:
            raise RuntimeError("reference VGAVisualGlobalDepth failed")
        return self.attr("Visual Step Depth")

    def sample_makegraph(self, src, maxdist=-1.0):
        src = np.ascontiguousarray(src, np.int32)
        e = C.c_int64()
        t = rlib().dmxref_sample_makegraph(self.h, _p(src), len(src), maxdist, C.byref(e))
        return t, e.value

    def sample_global(self, src, radius=-1):
        src = np.ascontiguousarray(src, np.int32)
        tn = np.zeros(len(src), np.int32)
        td = np.zeros(len(src), np.int64)
        t = rlib().dmxref_sample_global(self.h, _p(src), len(src), radius, _p(tn), _p(td))
        return t, tn, td
