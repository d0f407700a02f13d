#!/usr/bin/env python
"""bench.py -- throughput of the visibility-graph hot path (makegraph + VGA visibility global).

Workload (default): BASELINE.json configs[4], the configuration the metric is quoted on -- the 1024x1024 synthetic urban
grid, 1,048,576 open cells, 5.48e9 adjacency entries (C5).  Other configs: --workload C1|C2|C4|kind:W:H:seed.

One "step" = one pass of the hot path over the plan:
    sparkGraph2 for ALL N cells (every rank its share of the source rows)
    -> [N > 1: one exchange of the run-length rows over NCCL: every rank then holds the whole graph]
    -> all-sources BFS (VGAVisualGlobal) from a fixed, evenly spread subset of S sources made of whole spatial batches
       (every `stride`-th group of 512 sources of the library's batch order; S = N / stride; stride 1 = every source)
    -> per-source integers on the host of rank 0.
A full C5 pass (S = N) takes ~10 s on one GPU, the driver times 2 x 25 steps, hence the subset; the BFS cost per source
does not depend on which sources are chosen (every BFS covers the whole graph), and the subset keeps the library's
spatially compact batches intact.  It is identical for every N (strong scaling).

  metric       cells/s through makegraph + global BFS = 1 / (t_graph / N + t_bfs / S), t_graph = construction + exchange +
               derivation of the BFS row lists (once per graph), t_bfs = the BFS over the S sources; medians over the timed
               steps.  With S = N this is N / step time.  `full_job_ms_extrapolated` = t_graph + t_bfs * N / S.
  ms_per_step  median measured step (device time between CUDA events with the device idle at both ends, max over ranks)
  value        inputs (flat grid) already resident in HBM when the timed region starts
  e2e          the same step through the C ABI with HOST buffers (pinned): H2D of the grid and D2H of the results inside
  roofline     dominant kernel group = the BFS level kernels, CUDA-event timed inside the library on the stream they are
               launched on; algorithmic bytes per DESIGN.md with rows in the format the kernels read (pyramid node lists);
               the CSR-entry model of SURVEY.md 8d is reported beside it
  cpu_baseline the reference's own CPU implementation (oracle/_ref/libdmxref.so = unmodified reference sources) on a
               bounded sample, on all host cores (one process per core: the reference is single-threaded)

--impl reference times the reference CPU implementation on the same workload / metric (bounded sample per step).
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "VGA makegraph + global source-BFS cells/sec"
UNIT = "cells/s"
GROUP = 512  # sources per subset granule = the largest batch (8 words) of the library


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append([s.strip() for s in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            if self.proc:
                self.proc.terminate()
        except Exception:
            pass
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = max(mx, float(s[1]))
                for i, nme in enumerate(names):
                    if s[3 + i].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        # median under load: samples of an idle GPU (set-up phases) would pull it down
        busy = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def stats(xs):
    xs = [float(x) for x in xs]
    return {"min": min(xs), "median": float(np.median(xs)), "max": max(xs)}


def default_stride(workload):
    return 4 if workload == "C5" else 1


def workload_config(args, plan):
    return {"workload": f"{args.workload}: {plan.name} synthetic plan, spacing {plan.spacing}, makegraph over all cells + "
                        f"VGA visibility global radius {'n' if args.radius == -1 else args.radius}",
            "walls": len(plan.walls)}


# ------------------------------------------------------------------------------------------------ reference arm (CPU)

def reference_arm(args, plan):
    from oracle import pyoracle as po
    from oracle import refpool
    if not po.have_ref():
        po.build(ref=True)
    procs = os.cpu_count() or 1
    pool = refpool.RefPool(plan, procs, log=lambda *a: print("[reference]", *a, file=sys.stderr, flush=True))
    n = pool.n
    k = procs  # one source per process and step: the smallest sample that uses every core
    times, mk, bfs = [], [], []
    for it in range(args.warmup + args.steps):
        src = sorted(random.Random(100 + it).sample(range(n), min(k, n)))
        wall, mk_sum, bfs_sum, mk_max, bfs_max = pool.step(src, args.radius)
        if it >= args.warmup:
            times.append(wall)
            mk.append(mk_sum / len(src))
            bfs.append(bfs_sum / len(src))
    pool.close()
    ms = 1e3 * float(np.median(times))
    val = k / (ms / 1e3)
    sample = (f"each step = {k} sampled sources of the same plan, one per process on {procs} processes: the reference's "
              f"sparkPixel2(make=1) and the per-source body of VGAVisualGlobal::run (extractUnseen) on the complete graph "
              f"(made beforehand by the same processes, untimed: {pool.build_s:.0f} s, or loaded from the scratch cache); N={n}")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64 (sieve) + int (BFS)", "data": "synthetic",
            "config": dict(workload_config(args, plan), cells=n),
            "step_ms": stats([t * 1e3 for t in times]),
            "per_source_s": {"makegraph": float(np.median(mk)), "global_bfs": float(np.median(bfs))},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_sample(args, plan, log):
    """One bounded sample of the reference on all host cores (rank 0, N = 1)."""
    from oracle import pyoracle as po
    from oracle import refpool
    if not po.have_ref():
        return None
    procs = os.cpu_count() or 1
    pool = refpool.RefPool(plan, procs, log=log)
    n = pool.n
    src = sorted(random.Random(1).sample(range(n), min(procs * args.cpu_sources_per_core, n)))
    wall, mk_sum, bfs_sum, mk_max, bfs_max = pool.step(src, args.radius)
    pool.close()
    return {"value": len(src) / wall, "unit": UNIT, "cores": procs, "kind": "reference",
            "sample": f"{len(src)} sampled sources on {procs} processes ({wall:.1f} s wall): sparkPixel2(make=1) "
                      f"({mk_sum / len(src) * 1e3:.2f} ms per source) + per-source body of VGAVisualGlobal::run around the "
                      f"reference's extractUnseen ({bfs_sum / len(src):.3f} s per source) on the complete reference graph",
            "makegraph_cells_per_s_per_core": len(src) / mk_sum if mk_sum > 0 else None,
            "global_bfs_cells_per_s_per_core": len(src) / bfs_sum if bfs_sum > 0 else None,
            "host_cores_available": os.cpu_count()}


# ------------------------------------------------------------------------------------------------ our arm (GPU)

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C5", help="C1..C5 or kind:W:H:seed (default C5, BASELINE.json configs[4])")
    ap.add_argument("--radius", type=int, default=-1)
    ap.add_argument("--bfs-stride", type=int, default=0,
                    help="BFS sources = every stride-th group of 512 sources of the batch order (0 = 4 for C5, 1 otherwise)")
    ap.add_argument("--local-cells", type=int, default=4096,
                    help="VGA local on this many cells per step (split over the ranks), reported separately (0 = off)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm")
    ap.add_argument("--cpu-sources-per-core", type=int, default=1)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="vga_ctx_set_option for our arm (e.g. --opt bfs_words=4): A/B runs; recorded in config")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from depthmapx_b200 import plans
    plan = plans.by_name(args.workload)

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, plan)

    import torch
    from depthmapx_b200 import capi, multi
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- libvga_b200 has no CPU path")
    dist = None
    saved_stdout = None
    if world > 1:
        # NCCL prints a version banner on stdout: point fd 1 at stderr while the job runs and give it back for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    def log(*a):
        if rank == 0:
            print("[bench]", *a, file=sys.stderr, flush=True)

    t_setup = time.time()
    flat = capi.prepare(plan)  # host pre-steps (setGrid, blockLines, fill): not part of the hot path
    ctx = capi.Context(local_rank)
    for kv in args.opt:
        key, value = kv.split("=")
        ctx.set_option(key, int(value))
    dgrid = ctx.upload(flat)
    n = flat.n_filled
    # makegraph shards: contiguous source ranges balanced by estimated work (open area), SURVEY.md 8e
    if world > 1:
        parts = multi.partition_by_work(multi.estimate_source_work(flat.state, flat.cols, flat.rows), world)
    else:
        parts = [(0, n)]
    lo, hi = parts[rank]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB)
    log(f"{args.workload}: N={n}, host pre-steps {time.time() - t_setup:.1f} s")

    def barrier():
        if dist is not None:
            dist.barrier()
        ctx.sync()
        torch.cuda.synchronize(dev)

    def replicate(g, st):
        """N > 1: the run-length rows of every rank's shard, broadcast straight into the library's final allocation."""
        return multi.replicate_graph(ctx, g, n, hi - lo, dist, rank, world, dev, st)

    # ---- the fixed BFS source subset (set-up, untimed): whole groups of the library's batch order, evenly spread
    stride = args.bfs_stride if args.bfs_stride > 0 else default_stride(args.workload)
    g0 = ctx.build(dgrid, (lo, hi))
    full0 = replicate(g0, {}) if world > 1 else g0
    order = full0.batch_order()
    list_sizes = full0.list_sizes()
    edges_total = g0.entries
    if world > 1:
        full0.free()
    g0.free()
    ngroups = (n + GROUP - 1) // GROUP
    picked = [gi for gi in range(ngroups) if gi % stride == 0]
    my_groups = picked[(len(picked) * rank) // world:(len(picked) * (rank + 1)) // world]
    my_sources = np.concatenate([order[gi * GROUP:(gi + 1) * GROUP] for gi in my_groups]).astype(np.int64) if my_groups \
        else np.zeros(0, np.int64)
    counts = []
    for r in range(world):
        gs = picked[(len(picked) * r) // world:(len(picked) * (r + 1)) // world]
        counts.append(int(sum(min(GROUP, n - gi * GROUP) for gi in gs)))
    S = int(sum(counts))
    local_cells = args.local_cells
    l0 = max(0, n // 2 - local_cells // 2)
    l1 = min(n, l0 + local_cells)
    # every rank its share of the local-measure cells (the replicated run-length graph serves vga_local too)
    local_range = (l0 + ((l1 - l0) * rank) // world, l0 + ((l1 - l0) * (rank + 1)) // world)
    log(f"BFS subset: {S} of {n} sources ({len(picked)} groups of {GROUP}, stride {stride}); set-up {time.time() - t_setup:.1f} s")

    def step(grid_obj, st):
        """one pass of the hot path; grid_obj is a DeviceGrid (resident) or a FlatGrid (host buffers)"""
        t0 = time.perf_counter()
        g = ctx.build(grid_obj, (lo, hi))
        tb = ctx.timing()
        t1 = time.perf_counter()
        full = replicate(g, st) if world > 1 else g
        t2 = time.perf_counter()
        tn, td, hist, used = full.global_ints(args.radius, sources=my_sources)
        tg = ctx.timing()
        t3 = time.perf_counter()
        tl = None
        if local_cells > 0 and local_range[1] > local_range[0]:
            full.local_ints(local_range)
            tl = ctx.timing()
        t4 = time.perf_counter()
        if world > 1:
            # result gather to rank 0: source ordinal, tn, td, level histogram (padded to 62 levels)
            L = 62
            pack = np.zeros((len(my_sources), L + 3), np.int64)
            pack[:, 0] = my_sources
            pack[:, 1] = tn
            pack[:, 2] = td
            pack[:, 3:3 + min(L, hist.shape[1])] = hist[:, :L]
            res = multi.gather_results(torch.from_numpy(pack).to(dev), counts, dist, rank, world)
            torch.cuda.synchronize(dev)
            if rank == 0:
                st["_res"] = res  # the checksum is computed outside the timed region
            full.free()
        else:
            st["_local"] = (tn, td, hist)
        edges = g.entries
        g.free()
        t5 = time.perf_counter()
        st.update(build_ms=(t1 - t0) * 1e3, exchange_ms=(t2 - t1) * 1e3, bfs_ms=(t3 - t2) * 1e3, local_ms=(t4 - t3) * 1e3,
                  tail_ms=(t5 - t4) * 1e3, build_timing=tb, bfs_timing=tg, local_timing=tl, edges=edges, levels=used,
                  d2h_bytes=tn.nbytes + td.nbytes + hist.nbytes)

    def timed(grid_obj, steps, warmup, tag):
        per, acc = [], []
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        for it in range(warmup + steps):
            flush.fill_(it & 0xff)  # evict L2 between iterations
            barrier()
            st = {}
            ev0.record()
            step(grid_obj, st)
            barrier()
            ev1.record()
            ev1.synchronize()
            dt = ev0.elapsed_time(ev1)  # device idle at both marks: the span of the step on the device's clock
            if "_res" in st:
                r = st.pop("_res").cpu().numpy()
                st["checksum"] = {"sources": int(r.shape[0]), "sum_nodes": int(r[:, 1].sum()), "sum_depth": int(r[:, 2].sum()),
                                  "sum_hist_weighted": int((r[:, 3:] * (np.arange(r.shape[1] - 3) + 1)[None, :]).sum())}
            elif "_local" in st:
                tn_, td_, h_ = st.pop("_local")
                st["checksum"] = {"sources": int(len(tn_)), "sum_nodes": int(tn_.astype(np.int64).sum()), "sum_depth": int(td_.sum()),
                                  "sum_hist_weighted": int((h_.astype(np.int64) * (np.arange(h_.shape[1]) + 1)[None, :]).sum())}
            if os.environ.get("VGA_BENCH_DEBUG"):
                print(f"[rank {rank}] {tag} it={it} step={dt:.1f} build={st['build_ms']:.1f} exchange={st['exchange_ms']:.1f} "
                      f"bfs={st['bfs_ms']:.1f} (prep {st['bfs_timing']['prep_ms']:.1f}, level kernels "
                      f"{st['bfs_timing']['main_kernel_ms']:.1f}) local={st['local_ms']:.1f}", file=sys.stderr, flush=True)
            if it >= warmup:
                per.append(dt)
                acc.append(st)
        return per, acc

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    warm = max(args.warmup, 3)  # the contract's minimum
    per_res, st_res = timed(dgrid, args.steps, warm, "resident")
    pinned_inputs = False
    if args.no_e2e:
        per_e2e, st_e2e = per_res, st_res
    else:
        # host-buffer arm: the flat grid lives in pinned host memory (falls back to pageable if pinning is refused)
        flat_e2e, keep = flat, []
        try:
            def pin(a):
                t = torch.from_numpy(a.view(np.uint8).reshape(-1)).pin_memory()
                keep.append(t)
                return t.numpy().view(a.dtype).reshape(a.shape)
            flat_e2e = capi.FlatGrid(flat.cols, flat.rows, flat.spacing, flat.bl_x, flat.bl_y, pin(flat.state), pin(flat.line_off),
                                     pin(flat.lines) if flat.lines.size else flat.lines, flat.maxdist)
            pinned_inputs = all(t.is_pinned() for t in keep)
        except Exception:
            flat_e2e = flat
        per_e2e, st_e2e = timed(flat_e2e, args.steps, warm, "host-buffers")
    clocks = sampler.finish() if sampler else None

    def reduce(x, op):
        if dist is None:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    def rmax(x):
        return reduce(x, dist.ReduceOp.MAX if dist else None)

    def rsum(x):
        return reduce(x, dist.ReduceOp.SUM if dist else None)

    def med(sts, f):
        return float(np.median([f(s) for s in sts]))

    def summarize(per, sts):
        """medians over the timed steps, max over ranks (every collective happens on every rank)"""
        o = {}
        o["step_ms"] = rmax(float(np.median(per)))
        o["step_min"] = rmax(min(per))
        o["step_max"] = rmax(max(per))
        o["build_ms"] = rmax(med(sts, lambda s: s["build_ms"]))
        o["exchange_ms"] = rmax(med(sts, lambda s: s["exchange_ms"]))
        o["bfs_ms"] = rmax(med(sts, lambda s: s["bfs_ms"]))
        o["prep_ms"] = rmax(med(sts, lambda s: s["bfs_timing"]["prep_ms"]))
        o["local_ms"] = rmax(med(sts, lambda s: s["local_ms"]))
        o["bfs_main_ms"] = rmax(med(sts, lambda s: s["bfs_timing"]["main_kernel_ms"]))
        o["sieve_main_ms"] = rmax(med(sts, lambda s: s["build_timing"]["main_kernel_ms"]))
        o["build_kernel_ms"] = rmax(med(sts, lambda s: s["build_timing"]["kernel_ms"]))
        o["h2d_ms"] = rmax(med(sts, lambda s: s["build_timing"]["h2d_ms"]))
        o["d2h_ms"] = rmax(med(sts, lambda s: s["bfs_timing"]["d2h_ms"]))
        o["algo"] = rsum(med(sts, lambda s: s["bfs_timing"]["algo_bytes"]))
        o["algo_csr"] = rsum(med(sts, lambda s: s["bfs_timing"]["algo_bytes_csr"]))
        o["launches"] = rsum(float(np.sum([s["build_timing"]["launches"] + s["bfs_timing"]["launches"] +
                                           (s["local_timing"]["launches"] if s["local_timing"] else 0) for s in sts])))
        o["main_launches"] = med(sts, lambda s: s["bfs_timing"]["main_launches"])
        o["edges"] = rsum(float(sts[0]["edges"]))
        o["d2h_bytes"] = rsum(float(sts[0]["d2h_bytes"]))
        o["stage_ranges"] = {k: stats([s[k] for s in sts]) for k in ("build_ms", "exchange_ms", "bfs_ms", "local_ms")}
        o["stage_ranges"]["step_ms"] = stats(per)
        return o

    R = summarize(per_res, st_res)
    E = R if args.no_e2e else summarize(per_e2e, st_e2e)
    peak, peak_src = load_peaks()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    def cells_per_s(o):
        t_graph = o["build_ms"] + o["exchange_ms"] + o["prep_ms"]
        t_bfs = max(o["bfs_ms"] - o["prep_ms"], 1e-9)
        per_cell_ms = t_graph / n + t_bfs / max(S, 1)
        return 1e3 / per_cell_ms, t_graph + t_bfs * n / max(S, 1)

    val, full_ms = cells_per_s(R)
    val_e2e, full_ms_e2e = cells_per_s(E)
    achieved = R["algo"] / (R["bfs_main_ms"] * 1e-3) / 1e9 / world  # per GPU
    achieved_csr = R["algo_csr"] / (R["bfs_main_ms"] * 1e-3) / 1e9 / world
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r2_bfs_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            rec = tj.get(args.workload)
            if rec and world == 1:
                # measured dram bytes per source of the level kernels (ncu --set full capture) x the sources of one step
                traffic = float(rec["dram_bytes_per_source"]) * S
        except Exception:
            traffic = None
    cfg = dict(workload_config(args, plan), cells=n, edges=int(edges_total) if world == 1 else int(R["edges"]),
               bfs_sources=S, bfs_source_stride=stride,
               bfs_subset=f"every {stride}th group of {GROUP} sources of the library's spatial batch order (fixed, same for every N)",
               value_definition="1 / (t_graph/N + t_bfs/S): t_graph = makegraph + exchange + BFS row lists, t_bfs = BFS over the S sources; medians",
               row_lists=list_sizes, levels=int(st_res[0]["levels"]), bfs_batch_sources=64 * int(st_res[0]["bfs_timing"].get("batch_words", 0)), l2="flushed between iterations (256 MB write)",
               parallelism=f"makegraph rows and BFS sources sharded x{world}; graph replicated as run-length rows")
    if args.opt:
        cfg["options"] = dict(kv.split("=") for kv in args.opt)
    line = {
        "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": R["step_ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64 (sieve) + u64 bit-masks (BFS)", "data": "synthetic", "config": cfg,
        "full_job_ms_extrapolated": full_ms,
        "stages": {"makegraph_ms": R["build_ms"], "makegraph_cells_per_s": n / (R["build_ms"] * 1e-3),
                   "makegraph_edges_per_s": (edges_total if world == 1 else R["edges"]) / (R["build_ms"] * 1e-3),
                   "exchange_ms": R["exchange_ms"], "bfs_row_lists_ms": R["prep_ms"],
                   "global_bfs_ms": R["bfs_ms"], "global_bfs_cells_per_s": S / (max(R["bfs_ms"] - R["prep_ms"], 1e-9) * 1e-3),
                   "sieve_kernels_ms": R["sieve_main_ms"], "makegraph_kernels_ms": R["build_kernel_ms"],
                   "bfs_level_kernels_ms": R["bfs_main_ms"],
                   "local_ms": R["local_ms"] if local_cells > 0 else None,
                   "local_cells": (l1 - l0) if local_cells > 0 else 0,
                   "local_cells_per_s": ((l1 - l0) / (R["local_ms"] * 1e-3)) if local_cells > 0 and R["local_ms"] > 0 else None,
                   "ranges": R["stage_ranges"]},
        "e2e": {"value": val_e2e, "unit": UNIT, "ms_per_step": E["step_ms"], "full_job_ms_extrapolated": full_ms_e2e,
                "h2d_bytes_per_step": int(flat.input_bytes()), "d2h_bytes_per_step": int(E["d2h_bytes"]),
                "makegraph_ms": E["build_ms"], "exchange_ms": E["exchange_ms"], "global_bfs_ms": E["bfs_ms"],
                "h2d_ms": E["h2d_ms"], "d2h_ms": E["d2h_ms"], "ranges": E["stage_ranges"],
                "note": "vga_graph_build(host vga_grid) + vga_global_sources(host outputs); inputs in "
                        + ("pinned" if pinned_inputs else "pageable") + " host memory"},
        "gpu_launches": int(R["launches"]),
        "roofline": {"bound": "hbm", "kernel": "BFS level kernels k_push_delta/k_pyr_down/k_pyr_build/k_pull_nodes_coop/k_update/k_decide",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": traffic,
                     "byte_model": "rows as pyramid node lists (4 B per node id of the expanding vertices) + frontier / visited / next vectors",
                     "algorithmic_bytes_per_step": R["algo"], "kernel_ms_per_step": R["bfs_main_ms"],
                     "csr_model": {"algorithmic_bytes_per_step": R["algo_csr"], "achieved": achieved_csr, "frac": achieved_csr / peak,
                                   "note": "SURVEY.md 8d as written (4-byte CSR entries): what round 1's kernels streamed; kept for comparison"},
                     "launches_per_step": R["main_launches"]},
        "clocks": clocks,
        # size-independent result check: equal for every N on the same workload (sums over the subset's sources of
        # Node Count, total depth and the level histogram weighted by level+1)
        "result_checksum": st_res[-1].get("checksum"),
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            ref = cpu_baseline_sample(args, plan, log)
        except Exception as e:  # the checker library is optional on the GPU box
            ref = None
            line["cpu_baseline_error"] = repr(e)
        if ref:
            line["cpu_baseline"] = ref
    if dist is not None:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if dist is not None:
        os.dup2(2, 1)
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
