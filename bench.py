#!/usr/bin/env python
"""bench.py -- throughput of the visibility-graph hot path (makegraph + VGA visibility global).

One "step" = one full pass of the hot path over the workload's floor plan:
    sparkGraph2 (all N cells)  ->  all-sources BFS (N sources, radius n)  -> per-source integers on the host.
metric = cells/s through makegraph + VGA global (BASELINE.json), with the two stages also reported
separately (makegraph cells/s, global source-BFS cells/s).

  value        inputs (flat grid) already resident in HBM when the timed region starts
  e2e          the same step through the C ABI with HOST buffers: H2D of the grid and D2H of the
               results inside the timed region
  roofline     dominant kernel group = the BFS level kernels (push / pull / update), CUDA-event timed
               inside the library on the stream they are launched on; algorithmic bytes per DESIGN.md
  cpu_baseline the reference's own CPU implementation (oracle/_ref/libdmxref.so = unmodified reference
               sources) on a bounded sample, 1 core (the reference is single-threaded)

N > 1 (torchrun): makegraph rows are sharded by source range, the shards are all-gathered once over
NCCL, BFS sources are partitioned over the replicated adjacency, results are gathered to rank 0.
Strong scaling: the workload is fixed.

--impl reference times the reference CPU implementation on the same workload/metric (bounded sample).
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def sample_sources(n, k, seed=1):
    return sorted(random.Random(seed).sample(range(n), min(k, n)))


def reference_sample(plan, k_mk, k_bfs, radius, full_makegraph):
    """Time the reference CPU implementation.  Returns dict with per-cell seconds of both stages."""
    from oracle import pyoracle as po
    if not po.have_ref():
        return None
    a = po.RefMap(plan.walls, plan.spacing)
    for s in plan.seeds:
        a.fill(*s)
    n = a.n
    out = {"n": n, "kind": "reference"}
    if full_makegraph:
        t = a.makegraph()
        out["mk_s_per_cell"] = t / n
        out["mk_sample"] = f"full sparkGraph2 over all {n} cells ({t:.2f} s)"
        b = a
    else:
        src = sample_sources(n, k_mk, 2)
        t, edges = a.sample_makegraph(src)
        out["mk_s_per_cell"] = t / len(src)
        out["mk_sample"] = f"{len(src)} sampled sources through sparkPixel2(make=1) ({t:.2f} s)"
        b = po.RefMap(plan.walls, plan.spacing)
        for s in plan.seeds:
            b.fill(*s)
        b.makegraph()
    src = sample_sources(n, k_bfs, 1)
    t, tn, td = b.sample_global(src, radius)
    out["bfs_s_per_cell"] = t / len(src)
    out["bfs_sample"] = f"{len(src)} sampled sources, per-source body of VGAVisualGlobal::run around the reference's extractUnseen ({t:.2f} s)"
    out["map"] = b
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", help="C1..C5 or kind:W:H:seed (default C2, BASELINE.json configs[1])")
    ap.add_argument("--radius", type=int, default=-1)
    ap.add_argument("--vga-local", dest="local", action="store_true", help="also run VGA local in the step (not part of the headline metric)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer arm (big workloads run by hand)")
    ap.add_argument("--cpu-bfs-sources", type=int, default=48)
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE",
                    help="vga_ctx_set_option for our arm (e.g. --opt bfs_pull=1): A/B runs of opt-in kernels; recorded in config")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    from depthmapx_b200 import plans
    plan = plans.by_name(args.workload)
    workload_desc = {"workload": f"{args.workload}: {plan.name} synthetic plan, spacing {plan.spacing}, "
                                 f"makegraph + VGA visibility global radius {'n' if args.radius == -1 else args.radius}",
                     "walls": len(plan.walls)}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle import pyoracle as po
        if not po.have_ref():
            po.build(ref=True)
        k = 24
        a = po.RefMap(plan.walls, plan.spacing)
        for s in plan.seeds:
            a.fill(*s)
        n = a.n
        b = po.RefMap(plan.walls, plan.spacing)
        for s in plan.seeds:
            b.fill(*s)
        b.makegraph()  # untimed set-up: the BFS sample needs every Node
        times = []
        for it in range(args.warmup + args.steps):
            src = sample_sources(n, k, 100 + it)
            t1, _ = a.sample_makegraph(src)
            t2, _, _ = b.sample_global(src, args.radius)
            if it >= args.warmup:
                times.append(t1 + t2)
        ms = 1e3 * sum(times) / len(times)
        val = k / (ms / 1e3)
        sample = (f"each step = {k} sampled sources of the same plan through the reference's sparkPixel2(make=1) "
                  f"and the per-source body of VGAVisualGlobal::run (extractUnseen); N={n}")
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64+u64", "data": "synthetic",
                "config": dict(workload_desc, cells=n),
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "reference", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    from depthmapx_b200 import capi, multi
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device -- libvga_b200 has no CPU path")
    dist = None
    if world > 1:
        # NCCL prints a version banner on stdout: point fd 1 at stderr while the job runs and give it back
        # for the one JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    flat = capi.prepare(plan)  # host pre-steps (setGrid, blockLines, fill): not part of the hot path
    ctx = capi.Context(local_rank)
    for kv in args.opt:
        key, value = kv.split("=")
        ctx.set_option(key, int(value))
    if args.opt:
        workload_desc["options"] = dict(kv.split("=") for kv in args.opt)
    dgrid = ctx.upload(flat)
    n = flat.n_filled
    lo, hi = multi.partition(n, world)[rank]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > L2 (126 MB)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ctx.sync()

    def step(grid_obj, stats):
        """one pass of the hot path; grid_obj is a DeviceGrid (resident) or a FlatGrid (host buffers)"""
        t0 = time.perf_counter()
        g = ctx.build(grid_obj, (lo, hi))
        tb = ctx.timing()
        t1 = time.perf_counter()
        full = g
        if world > 1:
            rp_ptr, adj_ptr, ne = g.device_rows()
            rows = hi - lo
            rp_local = multi.wrap(rp_ptr, (rows + 1) * 8, torch.int64, dev)
            adj_local = multi.wrap(adj_ptr, ne * 4, torch.int32, dev)[:ne]
            rp_full, adj_full, base = multi.allgather_rows(rp_local, adj_local, dist, world)
            torch.cuda.synchronize(dev)
            full = ctx.graph_from_device_rows(n, g.ghosts, rp_full.data_ptr(), adj_full.data_ptr(), base)
            full.set_cell_refs(g.cell_refs())  # coordinates -> spatially coherent BFS batches
            del rp_full, adj_full, rp_local, adj_local
            if base * 4 > (4 << 30):
                torch.cuda.empty_cache()  # the library holds its own copy now; give multi-GB staging buffers back
            stats["gather_bytes"] = int(base) * 4
        t2 = time.perf_counter()
        tn, td, hist, used = full.global_ints(args.radius, (lo, hi))
        tg = ctx.timing()
        t3 = time.perf_counter()
        local_ms = 0.0
        if args.local:
            tl0 = time.perf_counter()
            full.local_ints((lo, hi))
            local_ms = (time.perf_counter() - tl0) * 1e3
        if world > 1:
            # result gather to rank 0: tn, td, level histogram (padded to 64 levels)
            L = 64
            pack = np.zeros((hi - lo, L + 2), np.int64)
            pack[:, 0] = tn
            pack[:, 1] = td
            pack[:, 2:2 + min(L, hist.shape[1])] = hist[:, :L]
            mine = torch.from_numpy(pack).to(dev)
            cnts = [e - s for s, e in multi.partition(n, world)]
            res = multi.gather_results(mine, cnts, dist, rank, world)
            torch.cuda.synchronize(dev)
            if rank == 0:
                stats["_gathered"] = res  # checksum is computed outside the timed region
            full.free()
        else:
            stats["_local"] = (tn, td, hist)
        g.free()
        t4 = time.perf_counter()
        stats.update(build_ms=(t1 - t0) * 1e3, gather_ms=(t2 - t1) * 1e3, bfs_ms=(t3 - t2) * 1e3, total_ms=(t4 - t0) * 1e3,
                     build_timing=tb, bfs_timing=tg, edges=g.entries, levels=used, local_ms=local_ms,
                     d2h_bytes=tn.nbytes + td.nbytes + hist.nbytes)
        return tn, td, hist

    def timed(grid_obj, steps, warmup):
        per, acc = [], []
        for it in range(warmup + steps):
            flush.fill_(it & 0xff)  # evict L2 between iterations
            barrier()
            st = {}
            t0 = time.perf_counter()
            step(grid_obj, st)
            barrier()
            dt = (time.perf_counter() - t0) * 1e3
            if os.environ.get("VGA_BENCH_DEBUG"):
                print(f"[rank {rank}] it={it} total={dt:.1f} build={st['build_ms']:.1f} gather={st['gather_ms']:.1f} "
                      f"bfs={st['bfs_ms']:.1f} (kernels {st['bfs_timing']['kernel_ms']:.1f}, level kernels "
                      f"{st['bfs_timing']['main_kernel_ms']:.1f}) tail={st['total_ms'] - st['build_ms'] - st['gather_ms'] - st['bfs_ms']:.1f}",
                      file=sys.stderr, flush=True)
            # result checksum of this iteration (untimed); only scalars are kept
            if "_gathered" in st:
                r = st.pop("_gathered").cpu().numpy()
                st["checksum"] = {"sum_nodes": int(r[:, 0].sum()), "sum_depth": int(r[:, 1].sum()),
                                  "sum_hist_weighted": int((r[:, 2:] * (np.arange(r.shape[1] - 2) + 1)[None, :]).sum())}
            elif "_local" in st:
                tn_, td_, h_ = st.pop("_local")
                st["checksum"] = {"sum_nodes": int(tn_.astype(np.int64).sum()), "sum_depth": int(td_.sum()),
                                  "sum_hist_weighted": int((h_.astype(np.int64) * (np.arange(h_.shape[1]) + 1)[None, :]).sum())}
            if it >= warmup:
                per.append(dt)
                acc.append(st)
        return per, acc

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # the contract's >= 3 warm-ups hold for the default workload; the minutes-long big workloads (C4/C5,
    # run by hand with --workload) may use fewer and can skip the host-buffer arm
    warm = max(args.warmup, 3) if args.workload == "C2" else args.warmup
    per_res, st_res = timed(dgrid, args.steps, warm)
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
        per_e2e, st_e2e = timed(flat_e2e, max(2, min(args.steps, 3)) if args.workload == "C2" else 1, 1 if args.workload == "C2" else 0)
    clocks = sampler.finish() if sampler else None

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ms_res = reduce_max(sum(per_res) / len(per_res))
    ms_e2e = reduce_max(sum(per_e2e) / len(per_e2e))
    build_ms = reduce_max(float(np.mean([s["build_ms"] for s in st_res])))
    bfs_ms = reduce_max(float(np.mean([s["bfs_ms"] for s in st_res])))
    gather_ms = reduce_max(float(np.mean([s["gather_ms"] for s in st_res])))
    # dominant kernel group: BFS level kernels (CUDA events inside the library, on its stream)
    bfs_main_ms = reduce_max(float(np.mean([s["bfs_timing"]["main_kernel_ms"] for s in st_res])))
    bfs_algo = reduce_sum(float(np.mean([s["bfs_timing"]["algo_bytes"] for s in st_res])))
    bfs_algo_runs = reduce_sum(float(np.mean([s["bfs_timing"].get("algo_bytes_runs", 0.0) for s in st_res])))
    sieve_main_ms = reduce_max(float(np.mean([s["build_timing"]["main_kernel_ms"] for s in st_res])))
    launches = reduce_sum(float(np.sum([s["build_timing"]["launches"] + s["bfs_timing"]["launches"] for s in st_res])))
    main_launches = float(np.mean([s["bfs_timing"]["main_launches"] for s in st_res]))
    edges = reduce_sum(float(st_res[0]["edges"]))
    # every collective must happen BEFORE the non-zero ranks leave (a lone all_reduce on rank 0 would
    # block until the NCCL watchdog aborts the process)
    local_ms = reduce_max(float(np.mean([s["local_ms"] for s in st_res]))) if args.local else None
    peak, peak_src = load_peaks()
    achieved = bfs_algo / (bfs_main_ms * 1e-3) / 1e9 / world  # per GPU
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC, "value": n / (ms_res * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 (sieve) + u64 bit-masks (BFS)", "data": "synthetic",
        "config": dict(workload_desc, cells=n, edges=int(edges), levels=int(st_res[0]["levels"]),
                       l2="flushed between iterations (256 MB write)", parallelism=f"source-sharded x{world}"),
        "stages": {"makegraph_ms": build_ms, "makegraph_cells_per_s": n / (build_ms * 1e-3),
                   "makegraph_edges_per_s": edges / (build_ms * 1e-3), "global_bfs_ms": bfs_ms,
                   "global_bfs_cells_per_s": n / (bfs_ms * 1e-3), "allgather_ms": gather_ms,
                   "sieve_kernels_ms": sieve_main_ms, "bfs_level_kernels_ms": bfs_main_ms,
                   "local_ms": local_ms},
        "e2e": {"value": n / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(flat.input_bytes()), "d2h_bytes_per_step": int(st_e2e[0]["d2h_bytes"]),
                "makegraph_ms": float(np.mean([s["build_ms"] for s in st_e2e])),
                "global_bfs_ms": float(np.mean([s["bfs_ms"] for s in st_e2e])),
                "h2d_ms": float(np.mean([s["build_timing"]["h2d_ms"] for s in st_e2e])),
                "d2h_ms": float(np.mean([s["bfs_timing"]["d2h_ms"] for s in st_e2e])),
                "note": "vga_graph_build(host vga_grid) + vga_global(host outputs); inputs in "
                        + ("pinned" if pinned_inputs else "pageable") + " host memory"},
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "BFS level kernels k_push/k_pull/k_update/k_decide",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": None,
                     "algorithmic_bytes_per_step": bfs_algo, "kernel_ms_per_step": bfs_main_ms,
                     # with run-length rows (--opt bfs_push=1): the same model at 8 bytes per run instead of 4 per entry
                     "algorithmic_bytes_per_step_run_length": bfs_algo_runs if bfs_algo_runs > 0 else None,
                     "launches_per_step": main_launches},
        "clocks": clocks,
        # size-independent result check: equal for every N on the same workload (sums over all sources of
        # Node Count, total depth and the level histogram weighted by level+1)
        "result_checksum": st_res[-1].get("checksum"),
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            ref = reference_sample(plan, 256, args.cpu_bfs_sources, args.radius, full_makegraph=(edges < 1.5e8))
        except Exception as e:  # the checker library is optional on the GPU box
            ref = None
            line["cpu_baseline_error"] = str(e)
        if ref:
            per_cell = ref["mk_s_per_cell"] + ref["bfs_s_per_cell"]
            line["cpu_baseline"] = {"value": 1.0 / per_cell, "unit": UNIT, "cores": 1, "kind": "reference",
                                    "sample": f"makegraph: {ref['mk_sample']}; global: {ref['bfs_sample']}",
                                    "makegraph_cells_per_s": 1.0 / ref["mk_s_per_cell"],
                                    "global_bfs_cells_per_s": 1.0 / ref["bfs_s_per_cell"],
                                    "host_cores_available": os.cpu_count()}
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
