"""ctypes bindings of the product libraries: the C ABI (include/vga_b200.h) and the host layer
(include/vga_host.h).  Importing this module never touches oracle/ -- there is no CPU compute
path; every compute call fails loudly without a CUDA device."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBDIR = os.path.join(HERE, "lib")


class VgaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vga error {code}: {msg}")
        self.code = code


class VgaGrid(C.Structure):
    _fields_ = [("cols", C.c_int32), ("rows", C.c_int32), ("spacing", C.c_double), ("bl_x", C.c_double),
                ("bl_y", C.c_double), ("maxdist", C.c_double), ("state", C.c_void_p), ("line_off", C.c_void_p),
                ("lines", C.c_void_p)]


class VgaTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("main_kernel_ms", C.c_double), ("launches", C.c_int64), ("main_launches", C.c_int64),
                ("algo_bytes", C.c_double), ("algo_bytes_csr", C.c_double), ("prep_ms", C.c_double),
                ("batch_words", C.c_int64)]


ABI_SYMBOLS = [
    "vga_last_error", "vga_version", "vga_device_count", "vga_ctx_create", "vga_ctx_destroy", "vga_ctx_set_callbacks",
    "vga_ctx_set_option", "vga_ctx_timing", "vga_ctx_sync", "vga_graph_build", "vga_grid_upload", "vga_dgrid_free",
    "vga_graph_build_resident", "vga_graph_from_csr", "vga_graph_free", "vga_graph_num_cells", "vga_graph_num_ghosts",
    "vga_graph_num_edges", "vga_graph_src_begin", "vga_graph_src_end", "vga_graph_csr", "vga_graph_cell_refs", "vga_graph_set_cell_refs",
    "vga_graph_node_stats", "vga_graph_set_noexpand", "vga_global", "vga_global_attributes", "vga_local", "vga_local_attributes", "vga_step_depth",
    "vga_graph_device_rows", "vga_graph_from_device_rows", "vga_global_sources", "vga_graph_batch_order",
    "vga_graph_device_runs", "vga_graph_from_device_runs", "vga_graph_runs_alloc", "vga_graph_runs_commit",
    "vga_graph_device_degrees", "vga_graph_list_sizes", "vga_metric", "vga_angular",
]
HOST_SYMBOLS = [
    "dmxh_last_error", "dmxh_map_create", "dmxh_map_destroy", "dmxh_map_grid", "dmxh_map_block_lines", "dmxh_map_fill",
    "dmxh_map_filled_count", "dmxh_map_flat", "dmxh_map_make_graph", "dmxh_map_vga_global", "dmxh_map_vga_local", "dmxh_map_vga_metric", "dmxh_map_vga_angular",
    "dmxh_map_columns", "dmxh_map_attr", "dmxh_map_grid_connections", "dmxh_map_graph", "dmxh_release_context",
    "dmxh_map_state", "dmxh_map_step_depth", "dmxh_map_select", "dmxh_map_selection", "dmxh_map_flat_rows", "dmxh_map_bins",
    "dmxh_map_encode_nodes", "dmxh_map_begin_graph", "dmxh_map_finish_graph", "dmxh_map_write_global", "dmxh_map_write_local",
    "dmxh_map_write_step_depth", "dmxh_graph_open", "dmxh_graph_close", "dmxh_graph_save", "dmxh_graph_num_maps",
    "dmxh_graph_displayed_map", "dmxh_graph_map", "dmxh_graph_walls", "dmxh_graph_new_map", "dmxh_graph_make_graph",
    "dmxh_graph_made", "dmxh_map_num_rows", "dmxh_map_fill_type", "dmxh_map_context_skip", "dmxh_map_merge", "dmxh_map_contracted_rows", "dmxh_map_radius_correction",
]

LEVEL_PREPARE_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int64, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32))
LEVEL_RUN_FN = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_int64), C.c_int64, C.POINTER(C.c_int32))

_abi = None
_host = None
vp = C.c_void_p
i64 = C.c_int64


def abi():
    global _abi
    if _abi is None:
        path = os.path.join(LIBDIR, "libvga_b200.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -m depthmapx_b200.build` (there is no fallback)")
        L = C.CDLL(path, mode=C.RTLD_GLOBAL)
        L.vga_last_error.restype = C.c_char_p
        L.vga_version.restype = C.c_char_p
        L.vga_ctx_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.vga_ctx_destroy.argtypes = [vp]
        L.vga_ctx_set_option.argtypes = [vp, C.c_char_p, i64]
        L.vga_ctx_timing.argtypes = [vp, C.POINTER(VgaTiming)]
        L.vga_ctx_sync.argtypes = [vp]
        L.vga_graph_build.argtypes = [vp, C.POINTER(VgaGrid), i64, i64, C.POINTER(vp)]
        L.vga_grid_upload.argtypes = [vp, C.POINTER(VgaGrid), C.POINTER(vp)]
        L.vga_dgrid_free.argtypes = [vp]
        L.vga_graph_build_resident.argtypes = [vp, vp, i64, i64, C.POINTER(vp)]
        L.vga_graph_from_csr.argtypes = [vp, i64, i64, vp, vp, vp, C.POINTER(vp)]
        L.vga_graph_free.argtypes = [vp]
        for f in ("vga_graph_num_cells", "vga_graph_num_ghosts", "vga_graph_num_edges", "vga_graph_src_begin",
                  "vga_graph_src_end"):
            getattr(L, f).restype = i64
            getattr(L, f).argtypes = [vp]
        L.vga_graph_csr.argtypes = [vp] * 5
        L.vga_graph_cell_refs.argtypes = [vp, vp]
        L.vga_graph_set_cell_refs.argtypes = [vp, vp, i64]
        L.vga_graph_node_stats.argtypes = [vp] * 7
        L.vga_graph_set_noexpand.argtypes = [vp, vp]
        L.vga_global.argtypes = [vp, vp, C.c_int, i64, i64, vp, vp, vp, C.c_int32, C.POINTER(C.c_int32)]
        L.vga_global_attributes.argtypes = [i64, vp, vp, vp, C.c_int32] + [vp] * 7
        L.vga_local.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp]
        L.vga_step_depth.argtypes = [vp, vp, vp, i64, vp]
        L.vga_metric.argtypes = [vp, vp, vp, vp, C.c_double, C.c_double, vp, i64, vp, vp, vp, vp, C.POINTER(i64)]
        L.vga_angular.argtypes = [vp, vp, vp, vp, C.c_double, vp, i64, vp, vp, vp, C.POINTER(i64)]
        L.vga_local_attributes.argtypes = [i64] + [vp] * 7
        L.vga_graph_device_rows.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
        L.vga_graph_from_device_rows.argtypes = [vp, i64, i64, vp, vp, i64, C.POINTER(vp)]
        L.vga_global_sources.argtypes = [vp, vp, C.c_int, vp, i64, vp, vp, vp, C.c_int32, C.POINTER(C.c_int32)]
        L.vga_graph_batch_order.argtypes = [vp, vp, vp]
        L.vga_graph_device_runs.argtypes = [vp, vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
        L.vga_graph_from_device_runs.argtypes = [vp, i64, i64, vp, vp, i64, vp, C.POINTER(vp)]
        L.vga_graph_runs_alloc.argtypes = [vp, i64, i64, i64, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
        L.vga_graph_runs_commit.argtypes = [vp]
        L.vga_graph_device_degrees.argtypes = [vp, vp, C.POINTER(vp)]
        L.vga_graph_list_sizes.argtypes = [vp, vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
        _abi = L
    return _abi


def host():
    global _host
    if _host is None:
        abi()
        path = os.path.join(LIBDIR, "libvga_host.so")
        if not os.path.exists(path):
            raise ImportError(f"{path} is missing: run `python -m depthmapx_b200.build`")
        H = C.CDLL(path)
        H.dmxh_last_error.restype = C.c_char_p
        H.dmxh_map_create.restype = vp
        H.dmxh_map_create.argtypes = [vp, C.c_int, C.c_double]
        H.dmxh_map_destroy.argtypes = [vp]
        H.dmxh_map_grid.argtypes = [vp] * 6
        H.dmxh_map_block_lines.argtypes = [vp]
        H.dmxh_map_fill.argtypes = [vp, C.c_double, C.c_double]
        H.dmxh_map_filled_count.argtypes = [vp]
        H.dmxh_map_flat.argtypes = [vp] * 6
        H.dmxh_map_make_graph.argtypes = [vp, C.c_int, C.c_double]
        H.dmxh_map_vga_global.argtypes = [vp, C.c_double, C.c_int]
        H.dmxh_map_vga_local.argtypes = [vp, C.c_int]
        H.dmxh_map_vga_metric.argtypes = [vp, C.c_double]
        H.dmxh_map_vga_angular.argtypes = [vp, C.c_double]
        H.dmxh_map_columns.argtypes = [vp, C.c_char_p, C.c_int]
        H.dmxh_map_attr.argtypes = [vp, C.c_char_p, vp]
        H.dmxh_map_grid_connections.argtypes = [vp, vp]
        H.dmxh_map_graph.restype = vp
        H.dmxh_map_graph.argtypes = [vp]
        H.dmxh_map_state.argtypes = [vp, vp]
        H.dmxh_map_step_depth.argtypes = [vp, vp, C.c_int]
        H.dmxh_map_select.argtypes = [vp, vp, C.c_int]
        H.dmxh_map_selection.restype = i64
        H.dmxh_map_selection.argtypes = [vp, vp]
        H.dmxh_map_flat_rows.argtypes = [vp] * 6
        H.dmxh_map_bins.argtypes = [vp] * 3
        H.dmxh_map_encode_nodes.argtypes = [vp] * 6
        H.dmxh_map_begin_graph.argtypes = [vp, C.c_int]
        H.dmxh_map_finish_graph.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        H.dmxh_map_write_global.argtypes = [vp, C.c_double, C.c_int, vp, vp, vp, C.c_int32]
        H.dmxh_map_write_local.argtypes = [vp, C.c_int, vp, vp, vp, vp]
        H.dmxh_map_write_step_depth.argtypes = [vp, vp]
        H.dmxh_graph_open.restype = vp
        H.dmxh_graph_open.argtypes = [C.c_char_p]
        H.dmxh_graph_close.argtypes = [vp]
        H.dmxh_graph_save.argtypes = [vp, C.c_char_p]
        H.dmxh_graph_num_maps.argtypes = [vp]
        H.dmxh_graph_displayed_map.argtypes = [vp]
        H.dmxh_graph_map.restype = vp
        H.dmxh_graph_map.argtypes = [vp, C.c_int]
        H.dmxh_graph_walls.restype = i64
        H.dmxh_graph_walls.argtypes = [vp, vp]
        H.dmxh_graph_new_map.restype = vp
        H.dmxh_graph_new_map.argtypes = [vp, C.c_double]
        H.dmxh_graph_make_graph.argtypes = [vp, C.c_int, C.c_double]
        H.dmxh_graph_made.argtypes = [vp]
        H.dmxh_map_merge.argtypes = [vp] + [C.c_double] * 4
        H.dmxh_map_fill_type.argtypes = [vp, C.c_double, C.c_double, C.c_int]
        H.dmxh_map_num_rows.restype = i64
        H.dmxh_map_num_rows.argtypes = [vp]
        H.dmxh_map_context_skip.argtypes = [vp, vp]
        H.dmxh_map_contracted_rows.argtypes = [vp] * 6
        H.dmxh_map_radius_correction.argtypes = [vp, C.c_int, LEVEL_PREPARE_FN, LEVEL_RUN_FN, vp, vp, vp, vp, C.c_int32]
        _host = H
    return _host


def check(rc):
    if rc != 0:
        raise VgaError(rc, abi().vga_last_error().decode())


def _p(a):
    return a.ctypes.data if a is not None else None


def device_count() -> int:
    return abi().vga_device_count()


class FlatGrid:
    """The vga_grid arrays (see include/vga_b200.h)."""

    def __init__(self, cols, rows, spacing, bl_x, bl_y, state, line_off, lines, maxdist=-1.0):
        self.cols, self.rows, self.spacing, self.bl_x, self.bl_y, self.maxdist = cols, rows, spacing, bl_x, bl_y, maxdist
        self.state = np.ascontiguousarray(state, np.uint16)
        self.line_off = np.ascontiguousarray(line_off, np.uint32)
        self.lines = np.ascontiguousarray(lines, np.float64).reshape(-1, 5)

    def c(self) -> VgaGrid:
        return VgaGrid(self.cols, self.rows, self.spacing, self.bl_x, self.bl_y, self.maxdist, _p(self.state),
                       _p(self.line_off), _p(self.lines) if self.lines.size else None)

    @property
    def n_filled(self) -> int:
        return int(((self.state & 2) != 0).sum())

    def input_bytes(self) -> int:
        return self.state.nbytes + self.line_off.nbytes + self.lines.nbytes


class Context:
    def __init__(self, device: int = 0):
        import weakref
        self.h = vp()
        self._children = weakref.WeakSet()  # graphs / device grids that must die before the context
        check(abi().vga_ctx_create(device, C.byref(self.h)))

    def close(self):
        if self.h:
            for ch in list(self._children):
                ch.free()
            abi().vga_ctx_destroy(self.h)
            self.h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, key: str, value: int):
        check(abi().vga_ctx_set_option(self.h, key.encode(), int(value)))

    def timing(self) -> dict:
        t = VgaTiming()
        check(abi().vga_ctx_timing(self.h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in VgaTiming._fields_}

    def sync(self):
        check(abi().vga_ctx_sync(self.h))

    def upload(self, grid: FlatGrid) -> "DeviceGrid":
        return DeviceGrid(self, grid)

    def build(self, grid, src=(0, -1)) -> "Graph":
        g = vp()
        if isinstance(grid, DeviceGrid):
            check(abi().vga_graph_build_resident(self.h, grid.h, src[0], src[1], C.byref(g)))
        else:
            cg = grid.c()
            check(abi().vga_graph_build(self.h, C.byref(cg), src[0], src[1], C.byref(g)))
        return Graph(self, g)

    def graph_from_csr(self, n_cells, n_ghosts, rowptr, col, bin=None) -> "Graph":
        rowptr = np.ascontiguousarray(rowptr, np.uint64)
        col = np.ascontiguousarray(col, np.uint32)
        b = np.ascontiguousarray(bin, np.uint8) if bin is not None else None
        g = vp()
        check(abi().vga_graph_from_csr(self.h, n_cells, n_ghosts, _p(rowptr), _p(col), _p(b), C.byref(g)))
        return Graph(self, g)

    def graph_from_device_runs(self, n_cells, n_ghosts, d_runptr: int, d_runs: int, n_runs: int, d_degree: int = 0) -> "Graph":
        """Adopt device-resident run-length rows of all cells (BFS-only graph); pointers are CUDA device addresses."""
        g = vp()
        check(abi().vga_graph_from_device_runs(self.h, n_cells, n_ghosts, d_runptr, d_runs, n_runs, d_degree or None, C.byref(g)))
        return Graph(self, g)

    def graph_runs_alloc(self, n_cells, n_ghosts, n_runs):
        """A BFS-only graph whose run-length rows the caller fills on the device: returns (graph, runptr address,
        runs address, degree address); call graph.runs_commit() when the buffers are complete."""
        g, rp, runs, deg = vp(), vp(), vp(), vp()
        check(abi().vga_graph_runs_alloc(self.h, n_cells, n_ghosts, n_runs, C.byref(g), C.byref(rp), C.byref(runs), C.byref(deg)))
        return Graph(self, g), rp.value, runs.value, deg.value

    def graph_from_device_rows(self, n_cells, n_ghosts, d_rowptr: int, d_adj: int, n_entries: int) -> "Graph":
        g = vp()
        check(abi().vga_graph_from_device_rows(self.h, n_cells, n_ghosts, d_rowptr, d_adj, n_entries, C.byref(g)))
        return Graph(self, g)


class DeviceGrid:
    def __init__(self, ctx: Context, grid: FlatGrid):
        self.ctx = ctx
        self.h = vp()
        cg = grid.c()
        check(abi().vga_grid_upload(ctx.h, C.byref(cg), C.byref(self.h)))
        ctx._children.add(self)

    def free(self):
        if getattr(self, "h", None):
            abi().vga_dgrid_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Graph:
    def __init__(self, ctx: Context, handle, owned=True):
        self.ctx, self.h, self.owned = ctx, handle, owned
        ctx._children.add(self)
        L = abi()
        self.n = L.vga_graph_num_cells(handle)
        self.ghosts = L.vga_graph_num_ghosts(handle)
        self.entries = L.vga_graph_num_edges(handle)
        self.src_begin = L.vga_graph_src_begin(handle)
        self.src_end = L.vga_graph_src_end(handle)

    def free(self):
        if self.h and self.owned:
            abi().vga_graph_free(self.h)
        self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def csr(self, bins=True):
        """(rowptr, col, bin, accepted); bins=False skips the two byte arrays (full-size graphs)."""
        rows = self.src_end - self.src_begin
        rowptr = np.zeros(rows + 1, np.uint64)
        col = np.empty(max(self.entries, 1), np.uint32)
        b = np.empty(max(self.entries, 1), np.uint8) if bins else None
        acc = np.empty(max(self.entries, 1), np.uint8) if bins else None
        check(abi().vga_graph_csr(self.h, _p(rowptr), _p(col), _p(b), _p(acc)))
        if not bins:
            return rowptr, col[:self.entries], None, None
        return rowptr, col[:self.entries], b[:self.entries], acc[:self.entries]

    def cell_refs(self):
        r = np.zeros(self.n + self.ghosts, np.int32)
        check(abi().vga_graph_cell_refs(self.h, _p(r)))
        return r

    def set_cell_refs(self, refs):
        r = np.ascontiguousarray(refs, np.int32)
        check(abi().vga_graph_set_cell_refs(self.h, _p(r), len(r)))

    def set_noexpand(self, flags):
        """uint8 [N]: cells that a radius-limited vga_global / vga_step_depth counts but does not expand; None clears."""
        if flags is None or len(flags) == 0:
            check(abi().vga_graph_set_noexpand(self.h, None))
        else:
            f = np.ascontiguousarray(flags, np.uint8)
            assert len(f) == self.n
            check(abi().vga_graph_set_noexpand(self.h, _p(f)))

    def node_stats(self):
        rows = self.src_end - self.src_begin
        out = dict(connectivity=np.zeros(rows, np.int32), sum_d=np.zeros(rows), sum_d2=np.zeros(rows),
                   far=np.zeros((rows, 32), np.float32), bin_count=np.zeros((rows, 32), np.int32),
                   gridconn=np.zeros(rows, np.uint8))
        check(abi().vga_graph_node_stats(self.h, _p(out["connectivity"]), _p(out["sum_d"]), _p(out["sum_d2"]),
                                         _p(out["far"]), _p(out["bin_count"]), _p(out["gridconn"])))
        return out

    def device_rows(self):
        rp, adj, ne = vp(), vp(), i64()
        check(abi().vga_graph_device_rows(self.h, C.byref(rp), C.byref(adj), C.byref(ne)))
        return rp.value, adj.value, ne.value

    def device_runs(self):
        """(device address of runptr, of the (first, length) run pairs, number of runs) of the rows this graph holds."""
        rp, runs, nr = vp(), vp(), i64()
        check(abi().vga_graph_device_runs(self.ctx.h, self.h, C.byref(rp), C.byref(runs), C.byref(nr)))
        return rp.value, runs.value, nr.value

    def runs_commit(self):
        check(abi().vga_graph_runs_commit(self.h))

    def device_degrees(self):
        d = vp()
        check(abi().vga_graph_device_degrees(self.ctx.h, self.h, C.byref(d)))
        return d.value

    def list_sizes(self):
        """dict(out_runs, out_nodes, in_runs, in_nodes): sizes of the row lists the BFS reads."""
        a, b, c, d = i64(), i64(), i64(), i64()
        check(abi().vga_graph_list_sizes(self.ctx.h, self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return {"out_runs": a.value, "out_nodes": b.value, "in_runs": c.value, "in_nodes": d.value}

    def batch_order(self):
        """Ordinals of all N sources in the order vga_global batches them (spatially compact groups)."""
        o = np.zeros(self.n, np.int32)
        check(abi().vga_graph_batch_order(self.ctx.h, self.h, _p(o)))
        return o

    def global_ints(self, radius=-1, src=None, maxl=32, sources=None):
        """src = (begin, end) range of ordinals, or sources = explicit list of ordinals (outputs in list order)."""
        lst = None
        if sources is not None:
            lst = np.ascontiguousarray(sources, np.int64)
            b, e = 0, len(lst)
        else:
            b, e = (0, self.n) if src is None else src
        k = e - b
        while True:
            tn = np.zeros(k, np.int32)
            td = np.zeros(k, np.int64)
            dist = np.zeros((k, maxl), np.int32)
            used = C.c_int32(0)
            if lst is not None:
                rc = abi().vga_global_sources(self.ctx.h, self.h, radius, _p(lst), k, _p(tn), _p(td), _p(dist), maxl, C.byref(used))
            else:
                rc = abi().vga_global(self.ctx.h, self.h, radius, b, e, _p(tn), _p(td), _p(dist), maxl, C.byref(used))
            if rc == -5 and used.value > maxl:
                maxl = used.value
                continue
            check(rc)
            return tn, td, dist, used.value

    def step_depth(self, sources):
        """Visual step depth from a set of cells (x-major ordinals): int32 [N], -1 = not reached."""
        src = np.ascontiguousarray(sources, np.int64)
        d = np.zeros(self.n, np.int32)
        check(abi().vga_step_depth(self.ctx.h, self.h, _p(src), len(src), _p(d)))
        return d

    def metric(self, blocked_adjacent, spacing, radius=-1.0, sources=None, partner=None):
        """VGAMetric::run for the listed source ordinals (None = all): float32 columns (Metric Mean Shortest-Path Angle,
        Metric Mean Shortest-Path Distance, Metric Mean Straight-Line Distance, Metric Node Count) + the number of
        turn-angle evaluations whose float rounding is not guaranteed (see include/vga_b200.h)."""
        ba = np.ascontiguousarray(blocked_adjacent, np.uint8)
        src = None if sources is None else np.ascontiguousarray(sources, np.int64)
        k = self.n if src is None else len(src)
        out = [np.zeros(k, np.float32) for _ in range(4)]
        unsafe = i64()
        mp = None if partner is None else np.ascontiguousarray(partner, np.int32)
        check(abi().vga_metric(self.ctx.h, self.h, _p(ba), _p(mp), float(spacing), float(radius), None if src is None else _p(src), k,
                               *[_p(a) for a in out], C.byref(unsafe)))
        return tuple(out) + (unsafe.value,)

    def angular(self, blocked_adjacent, radius=-1.0, sources=None, partner=None):
        """VGAAngular::run: (Angular Mean Depth, Angular Total Depth, Angular Node Count, unsafe angle evaluations)."""
        ba = np.ascontiguousarray(blocked_adjacent, np.uint8)
        src = None if sources is None else np.ascontiguousarray(sources, np.int64)
        k = self.n if src is None else len(src)
        out = [np.zeros(k, np.float32) for _ in range(3)]
        unsafe = i64()
        mp = None if partner is None else np.ascontiguousarray(partner, np.int32)
        check(abi().vga_angular(self.ctx.h, self.h, _p(ba), _p(mp), float(radius), None if src is None else _p(src), k,
                                *[_p(a) for a in out], C.byref(unsafe)))
        return tuple(out) + (unsafe.value,)

    def local_ints(self, src=None):
        b, e = (0, self.n) if src is None else src
        k = e - b
        cl = np.zeros(k, np.int64)
        kk = np.zeros(k, np.int32)
        tot = np.zeros(k, np.int32)
        ctl = np.zeros(k, np.float32)
        check(abi().vga_local(self.ctx.h, self.h, b, e, _p(cl), _p(kk), _p(tot), _p(ctl)))
        return cl, kk, tot, ctl


GLOBAL_COLS = ["Visual Node Count", "Visual Mean Depth", "Visual Integration [HH]", "Visual Integration [P-value]",
               "Visual Integration [Tekl]", "Visual Entropy", "Visual Relativised Entropy"]
LOCAL_COLS = ["Visual Clustering Coefficient", "Visual Control", "Visual Controllability"]


def global_attributes(tn, td, dist):
    n = len(tn)
    dist = np.ascontiguousarray(dist, np.int32)
    outs = [np.zeros(n, np.float32) for _ in range(7)]
    check(abi().vga_global_attributes(n, _p(np.ascontiguousarray(tn, np.int32)), _p(np.ascontiguousarray(td, np.int64)),
                                      _p(dist), dist.shape[1], *[_p(o) for o in outs]))
    return dict(zip(GLOBAL_COLS, outs))


def local_attributes(cl, kk, tot, ctl):
    n = len(cl)
    outs = [np.zeros(n, np.float32) for _ in range(3)]
    check(abi().vga_local_attributes(n, _p(np.ascontiguousarray(cl, np.int64)), _p(np.ascontiguousarray(kk, np.int32)),
                                     _p(np.ascontiguousarray(tot, np.int32)), _p(np.ascontiguousarray(ctl, np.float32)),
                                     *[_p(o) for o in outs]))
    return dict(zip(LOCAL_COLS, outs))


class HostMap:
    """dmx::PointMap through the flat C view (mirrors PointMap's setGrid/makePoints/sparkGraph2 and the
    two VGA modules' run())."""

    def __init__(self, walls=None, spacing=1.0, handle=None, owner=None):
        """walls + spacing: a new map (PointMap(region, drawing) + setGrid); handle: a map owned by a GraphFile."""
        self._owner = owner  # keeps the GraphFile alive
        if handle is not None:
            self.h = handle
        else:
            w = np.ascontiguousarray(walls, np.float64).reshape(-1, 4)
            self.h = host().dmxh_map_create(_p(w), w.shape[0], spacing)
        self._refresh()

    def _refresh(self):
        c, r = C.c_int32(), C.c_int32()
        s, bx, by = C.c_double(), C.c_double(), C.c_double()
        host().dmxh_map_grid(self.h, C.addressof(c), C.addressof(r), C.addressof(s), C.addressof(bx), C.addressof(by))
        self.cols, self.rows, self.spacing, self.bl_x, self.bl_y = c.value, r.value, s.value, bx.value, by.value

    def __del__(self):
        if getattr(self, "h", None) and self._owner is None:
            host().dmxh_map_destroy(self.h)
        self.h = None

    def _ret(self, rc):
        if rc < 0:
            raise RuntimeError(host().dmxh_last_error().decode())
        return bool(rc)

    def block_lines(self):
        return self._ret(host().dmxh_map_block_lines(self.h))

    def fill(self, x, y, fill_type=0):
        """makePoints; fill_type 1 = semi-fill (context fill)."""
        if fill_type:
            return self._ret(host().dmxh_map_fill_type(self.h, x, y, fill_type))
        return self._ret(host().dmxh_map_fill(self.h, x, y))

    def context_skip(self):
        """uint8 [N] flags of the cells the analyses skip as sources (empty when there are none)."""
        out = np.zeros(max(self.n, 1), np.uint8)
        if not host().dmxh_map_context_skip(self.h, _p(out)):
            return np.zeros(0, np.uint8)
        return out[:self.n]

    @property
    def n(self):
        return host().dmxh_map_filled_count(self.h)

    def flat(self, maxdist=-1.0) -> FlatGrid:
        cells, nseg = i64(), i64()
        host().dmxh_map_flat(self.h, C.addressof(cells), C.addressof(nseg), None, None, None)
        state = np.zeros(cells.value, np.uint16)
        off = np.zeros(cells.value + 1, np.uint32)
        lines = np.zeros((max(nseg.value, 1), 5))
        host().dmxh_map_flat(self.h, None, None, _p(state), _p(off), _p(lines))
        return FlatGrid(self.cols, self.rows, self.spacing, self.bl_x, self.bl_y, state, off, lines[:nseg.value], maxdist)

    def make_graph(self, boundary=False, maxdist=-1.0):
        return self._ret(host().dmxh_map_make_graph(self.h, int(boundary), maxdist))

    def vga_global(self, radius=-1.0, simple=False):
        return self._ret(host().dmxh_map_vga_global(self.h, radius, int(simple)))

    def vga_local(self, simple=False):
        return self._ret(host().dmxh_map_vga_local(self.h, int(simple)))

    def vga_metric(self, radius=-1.0):
        return self._ret(host().dmxh_map_vga_metric(self.h, float(radius)))

    def vga_angular(self, radius=-1.0):
        return self._ret(host().dmxh_map_vga_angular(self.h, float(radius)))

    def columns(self):
        buf = C.create_string_buffer(8192)
        host().dmxh_map_columns(self.h, buf, 8192)
        return [s for s in buf.value.decode(errors="replace").split("\n") if s]

    def attr(self, name):
        rows = int(host().dmxh_map_num_rows(self.h))  # cells that had a Node made (= filled cells of a made graph)
        out = np.zeros(max(rows, 1), np.float32)[:rows]
        if not host().dmxh_map_attr(self.h, name.encode(), _p(out)):
            raise KeyError(name)
        return out

    def grid_connections(self):
        out = np.zeros(self.n, np.uint8)
        host().dmxh_map_grid_connections(self.h, _p(out))
        return out

    def state(self):
        out = np.zeros(self.cols * self.rows, np.uint16)
        host().dmxh_map_state(self.h, _p(out))
        return out

    def select(self, points):
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 2)
        return self._ret(host().dmxh_map_select(self.h, _p(pts), pts.shape[0]))

    def selection(self):
        n = host().dmxh_map_selection(self.h, None)
        out = np.zeros(n, np.int32)
        host().dmxh_map_selection(self.h, _p(out))
        return out

    def step_depth(self, points):
        """VGAVisualGlobalDepth::run from the cells containing `points` (GPU)."""
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 2)
        return self._ret(host().dmxh_map_step_depth(self.h, _p(pts), pts.shape[0]))

    def flat_rows(self):
        """(rowptr, ref, bin): the run-length adjacency in Node::first/next order."""
        n, e = i64(), i64()
        self._ret(host().dmxh_map_flat_rows(self.h, C.addressof(n), C.addressof(e), None, None, None))
        rowptr = np.zeros(n.value + 1, np.uint64)
        ref = np.zeros(max(e.value, 1), np.int32)
        b = np.zeros(max(e.value, 1), np.uint8)
        self._ret(host().dmxh_map_flat_rows(self.h, None, None, _p(rowptr), _p(ref), _p(b)))
        return rowptr, ref[:e.value], b[:e.value]

    def bins(self):
        cnt = np.zeros((self.n, 32), np.uint16)
        dist = np.zeros((self.n, 32), np.float32)
        self._ret(host().dmxh_map_bins(self.h, _p(cnt), _p(dist)))
        return cnt, dist

    def encode_nodes(self, rowptr, ref, b, accepted=None, far=None):
        rowptr = np.ascontiguousarray(rowptr, np.uint64)
        ref = np.ascontiguousarray(ref, np.int32)
        b = np.ascontiguousarray(b, np.uint8)
        acc = None if accepted is None else np.ascontiguousarray(accepted, np.uint8)
        f = None if far is None else np.ascontiguousarray(far, np.float32)
        return self._ret(host().dmxh_map_encode_nodes(self.h, _p(rowptr), _p(ref), _p(b), None if acc is None else _p(acc),
                                                      None if f is None else _p(f)))

    def merge(self, ax, ay, bx, by):
        """-m LINK -lnk ax,ay,bx,by"""
        return self._ret(host().dmxh_map_merge(self.h, ax, ay, bx, by))

    def contracted_rows(self):
        """(rowptr, col, primary): the adjacency the BFS analyses run on when cells are merged."""
        n, e = i64(), i64()
        self._ret(host().dmxh_map_contracted_rows(self.h, C.addressof(n), C.addressof(e), None, None, None))
        rowptr = np.zeros(n.value + 1, np.uint64)
        col = np.zeros(max(e.value, 1), np.uint32)
        primary = np.zeros(n.value, np.int32)
        self._ret(host().dmxh_map_contracted_rows(self.h, None, None, _p(rowptr), _p(col), _p(primary)))
        return rowptr, col[:e.value], primary

    def radius_correction(self, radius, level_to, total_nodes, total_depth, dist):
        """dmx::PointMap::radiusCorrection with a caller-supplied BFS: level_to(t_rowptr, t_col, seeds) -> int32 [n]
        (level of every vertex in a BFS from `seeds` over the given CSR, -1 = unreached).  Arrays are updated in place."""
        st = {}

        def prep(_u, n, rp, col):
            st["rp"] = np.ctypeslib.as_array(rp, (n + 1,)).copy()
            st["col"] = np.ctypeslib.as_array(col, (max(int(st["rp"][-1]), 1),)).copy()[:int(st["rp"][-1])]
            st["n"] = n

        def run(_u, seeds, ns, level):
            s = np.ctypeslib.as_array(seeds, (ns,)).copy() if ns else np.zeros(0, np.int64)
            out = np.ctypeslib.as_array(level, (st["n"],))
            out[:] = np.asarray(level_to(st["rp"], st["col"], s), np.int32)

        a, b = LEVEL_PREPARE_FN(prep), LEVEL_RUN_FN(run)
        assert total_nodes.dtype == np.int32 and total_depth.dtype == np.int64 and dist.dtype == np.int32
        assert total_nodes.flags.c_contiguous and total_depth.flags.c_contiguous and dist.flags.c_contiguous
        return self._ret(host().dmxh_map_radius_correction(self.h, int(radius), a, b, None, _p(total_nodes), _p(total_depth),
                                                           _p(dist), dist.shape[1]))

    def begin_graph(self, boundary=False):
        return self._ret(host().dmxh_map_begin_graph(self.h, int(boundary)))

    def finish_graph(self, boundary, connectivity, sum_d, sum_d2, gridconn):
        a = np.ascontiguousarray(connectivity, np.int32)
        b = np.ascontiguousarray(sum_d, np.float64)
        c = np.ascontiguousarray(sum_d2, np.float64)
        d = np.ascontiguousarray(gridconn, np.uint8)
        return self._ret(host().dmxh_map_finish_graph(self.h, int(boundary), _p(a), _p(b), _p(c), _p(d)))

    def write_global(self, radius, simple, total_nodes, total_depth, dist):
        tn = np.ascontiguousarray(total_nodes, np.int32)
        td = np.ascontiguousarray(total_depth, np.int64)
        d = np.ascontiguousarray(dist, np.int32)
        return self._ret(host().dmxh_map_write_global(self.h, radius, int(simple), _p(tn), _p(td), _p(d), d.shape[1]))

    def write_local(self, simple, cluster, k, total, control):
        a = np.ascontiguousarray(cluster, np.int64)
        b = np.ascontiguousarray(k, np.int32)
        c = np.ascontiguousarray(total, np.int32)
        d = np.ascontiguousarray(control, np.float32)
        return self._ret(host().dmxh_map_write_local(self.h, int(simple), _p(a), _p(b), _p(c), _p(d)))

    def write_step_depth(self, depth):
        d = np.ascontiguousarray(depth, np.int32)
        return self._ret(host().dmxh_map_write_step_depth(self.h, _p(d)))


class GraphFile:
    """dmx::GraphFile: a .graph container whose PointMap section is decoded / encoded by the host layer
    (MetaGraph::readFromStream / write, salalib/mgraph.cpp:2492-2763)."""

    def __init__(self, path):
        self.h = host().dmxh_graph_open(os.fsencode(path))
        if not self.h:
            raise RuntimeError(host().dmxh_last_error().decode())

    def __del__(self):
        if getattr(self, "h", None):
            host().dmxh_graph_close(self.h)
            self.h = None

    @property
    def num_maps(self):
        return host().dmxh_graph_num_maps(self.h)

    @property
    def displayed_map(self):
        return host().dmxh_graph_displayed_map(self.h)

    def map(self, i=None) -> HostMap:
        if i is None:
            i = self.displayed_map
        return HostMap(handle=host().dmxh_graph_map(self.h, i), owner=self)

    def walls(self):
        n = host().dmxh_graph_walls(self.h, None)
        out = np.zeros((max(n, 1), 4))
        host().dmxh_graph_walls(self.h, _p(out))
        return out[:n]

    def new_map(self, spacing) -> HostMap:
        h = host().dmxh_graph_new_map(self.h, spacing)
        if not h:
            raise RuntimeError(host().dmxh_last_error().decode())
        return HostMap(handle=h, owner=self)

    def make_graph(self, boundary=False, maxdist=-1.0):
        rc = host().dmxh_graph_make_graph(self.h, int(boundary), maxdist)
        if rc < 0:
            raise RuntimeError(host().dmxh_last_error().decode())
        return bool(rc)

    def graph_made(self):
        host().dmxh_graph_made(self.h)

    def save(self, path):
        if host().dmxh_graph_save(self.h, os.fsencode(path)) < 0:
            raise RuntimeError(host().dmxh_last_error().decode())


def blocked_adjacent(flat) -> np.ndarray:
    """uint8 [N] per filled cell (x-major): the cell is BLOCKED (Point::m_state & 0x0004) or one of its eight neighbours
    inside the grid is (Point::blocked || PointMap::blockedAdjacent, salalib/pointdata.cpp:1016-1066) -- the cells
    VGAMetric / VGAAngular expand (salalib/ngraph.cpp:71, 82)."""
    st = np.asarray(flat.state).reshape(flat.cols, flat.rows)
    pad = np.zeros((flat.cols + 2, flat.rows + 2), bool)
    pad[1:-1, 1:-1] = (st & 0x0004) != 0
    near = np.zeros((flat.cols, flat.rows), bool)
    for dx in (0, 1, 2):
        for dy in (0, 1, 2):
            near |= pad[dx:dx + flat.cols, dy:dy + flat.rows]
    return near.reshape(-1)[((st & 0x0002) != 0).reshape(-1)].astype(np.uint8)


def prepare(plan, maxdist=-1.0) -> FlatGrid:
    """Plan -> flat hot-path inputs via the host layer (setGrid, blockLines, fill)."""
    m = HostMap(plan.walls, plan.spacing)
    for s in plan.seeds:
        m.fill(*s)
    return m.flat(maxdist)
