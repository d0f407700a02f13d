"""TEST INFRASTRUCTURE ONLY -- the reference's CPU implementation on all host cores, for bench.py's reference arm
and cpu_baseline leg (never imported by depthmapx_b200/).

The reference (oracle/_ref/libdmxref.so = unmodified sources + oracle/ref_harness.cpp) is single-threaded and not
re-entrant, so parallelism comes from PROCESSES:

  1. the graph the BFS sample needs (every Node of the plan -- minutes of sparkPixel2 for the 10^6-cell workload) is
     made once by P forked builders, each running the body of sparkGraph2's source loop for an interleaved share of the
     sources and handing its Nodes over through the reference's own Node::write / Node::read; the parts are cached in
     a scratch directory ($VGA_REFCACHE or <tmp>/vga_refcache, 1.7 GB for the 10^6-cell plan), so later invocations on the same
     box load them;
  2. the parent loads all parts and forks P workers that share the graph copy-on-write;
  3. a timed step hands every worker a share of the step's sampled sources: sparkPixel2(make = 1) for each (the
     construction half of the metric) and the per-source body of VGAVisualGlobal::run around the reference's
     extractUnseen (the BFS half); the step's time is the wall time until the last worker is done.
"""
from __future__ import annotations

import hashlib
import multiprocessing as mp
import os
import tempfile
import time

import numpy as np

from oracle import pyoracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cache_dir(plan, spacing):
    h = hashlib.sha1()
    h.update(np.ascontiguousarray(plan.walls, np.float64).tobytes())
    h.update(repr((spacing, plan.seeds)).encode())
    base = os.environ.get("VGA_REFCACHE", os.path.join(tempfile.gettempdir(), "vga_refcache"))
    return os.path.join(base, h.hexdigest()[:16])


def _new_map(plan):
    m = po.RefMap(plan.walls, plan.spacing)
    for s in plan.seeds:
        m.fill(*s)
    return m


def _builder(plan, share, path, q):
    try:
        m = _new_map(plan)
        src = np.ascontiguousarray(share, np.int32)
        t = po.rlib().dmxref_build_nodes_to_file(m.h, src.ctypes.data, len(src), -1.0, os.fsencode(path + ".tmp"))
        if t < 0:
            raise RuntimeError("dmxref_build_nodes_to_file failed")
        os.replace(path + ".tmp", path)
        q.put(("ok", t))
    except Exception as e:  # pragma: no cover
        q.put(("error", repr(e)))


def _worker(m, conn):
    """Serves (sources, radius, want_makegraph) requests on the inherited map."""
    while True:
        msg = conn.recv()
        if msg is None:
            break
        src, radius, do_mk = msg
        t_mk = t_bfs = 0.0
        edges = 0
        if len(src):
            if do_mk:
                t_mk, edges = m.sample_makegraph(src)
            t_bfs, tn, td = m.sample_global(src, radius)
        conn.send((t_mk, t_bfs, edges))


class RefPool:
    def __init__(self, plan, procs=None, log=None):
        self.plan = plan
        self.procs = max(1, procs or os.cpu_count() or 1)
        self.log = log or (lambda *a: None)
        L = po.rlib()
        import ctypes as C
        L.dmxref_build_nodes_to_file.restype = C.c_double
        L.dmxref_build_nodes_to_file.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_char_p]
        L.dmxref_load_nodes.restype = C.c_int64
        L.dmxref_load_nodes.argtypes = [C.c_void_p, C.c_char_p]
        L.dmxref_node_count.restype = C.c_int64
        L.dmxref_node_count.argtypes = [C.c_void_p]
        t0 = time.time()
        self.map = _new_map(plan)
        self.n = self.map.n
        self.log(f"reference map: {self.n} filled cells ({time.time() - t0:.1f} s)")
        self.build_s = self._ensure_graph()
        ctx = mp.get_context("fork")
        self.workers = []
        for _ in range(self.procs):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(self.map, b), daemon=True)
            p.start()
            b.close()
            self.workers.append((p, a))

    def _ensure_graph(self):
        """Every Node of the plan in self.map: from the cache, or made by self.procs builder processes."""
        d = _cache_dir(self.plan, self.plan.spacing)
        os.makedirs(d, exist_ok=True)
        P = self.procs
        meta = os.path.join(d, "parts.txt")
        parts = []
        if os.path.exists(meta):
            parts = [os.path.join(d, x) for x in open(meta).read().split()]
            if not parts or not all(os.path.exists(x) for x in parts):
                parts = []
        spent = 0.0
        if not parts:
            t0 = time.time()
            ctx = mp.get_context("fork")
            q = ctx.Queue()
            parts = [os.path.join(d, f"part_{c:02d}_of_{P:02d}.bin") for c in range(P)]
            procs = [ctx.Process(target=_builder, args=(self.plan, np.arange(c, self.n, P, dtype=np.int32), parts[c], q))
                     for c in range(P)]
            for p in procs:
                p.start()
            res = [q.get() for _ in procs]
            for p in procs:
                p.join()
            bad = [r for r in res if r[0] != "ok"]
            if bad:
                raise RuntimeError(f"reference graph build failed: {bad[0][1]}")
            spent = time.time() - t0
            open(meta, "w").write("\n".join(os.path.basename(x) for x in parts))
            self.log(f"reference graph made by {P} processes in {spent:.1f} s (sum of sparkPixel2 time "
                     f"{sum(r[1] for r in res):.1f} s)")
        t0 = time.time()
        total = 0
        for x in parts:
            k = po.rlib().dmxref_load_nodes(self.map.h, os.fsencode(x))
            if k < 0:
                raise RuntimeError(f"could not load {x}")
            total += k
        if total != self.n or po.rlib().dmxref_node_count(self.map.h) != self.n:
            raise RuntimeError(f"reference graph incomplete: {total} of {self.n} nodes")
        self.log(f"reference graph loaded ({time.time() - t0:.1f} s)")
        return spent

    def step(self, sources, radius=-1, makegraph=True):
        """One bounded sample: the sources are dealt to the workers; returns (wall seconds, sum of construction seconds,
        sum of BFS seconds, max construction seconds, max BFS seconds)."""
        src = np.ascontiguousarray(sources, np.int32)
        shares = [src[c::self.procs] for c in range(self.procs)]
        t0 = time.perf_counter()
        for (p, conn), sh in zip(self.workers, shares):
            conn.send((sh, radius, makegraph))
        res = [conn.recv() for (p, conn) in self.workers]
        wall = time.perf_counter() - t0
        return wall, sum(r[0] for r in res), sum(r[1] for r in res), max(r[0] for r in res), max(r[1] for r in res)

    def close(self):
        for p, conn in self.workers:
            try:
                conn.send(None)
            except Exception:
                pass
        for p, conn in self.workers:
            p.join(timeout=5)
        self.workers = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
