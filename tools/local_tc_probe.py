"""Measurement aid (not product code): the tensor-core formulation of the VGA local measures that SURVEY.md 7-K4 and the
north star ask to evaluate -- cluster = rowsum(A o (A.A)), total = nnz_row(A.A) with A the 0/1 adjacency as a dense
matrix -- timed through the LIBRARY GEMMs (cuBLASLt tensor cores: int8 -> int32 via torch._int_mm, and bf16 -> fp32 via
torch.mm(out_dtype=float32); both exact for 0/1 operands and counts < 2^24) against the library's own kernels on the same
cells.  A hand-written tcgen05 kernel could at best approach the library GEMM rate on dense tiles, so this bounds what
a tensor-core path can win.

    python tools/local_tc_probe.py C1            # whole map, results compared exactly
    python tools/local_tc_probe.py C4 8192 65536 # rate only: 8192 cells against a slab of 65536 columns (A is N^2 elements)
"""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from depthmapx_b200 import capi, plans

name = sys.argv[1]
cells = int(sys.argv[2]) if len(sys.argv) > 2 else 0
slab = int(sys.argv[3]) if len(sys.argv) > 3 else 0
flat = capi.prepare(plans.by_name(name))
ctx = capi.Context(0)
g = ctx.build(flat)
n, U = g.n, g.n + g.ghosts
lo = 0 if cells <= 0 else max(0, n // 2 - cells // 2)
hi = n if cells <= 0 else min(n, lo + cells)
dev = torch.device("cuda", 0)

ours = {}
for mode, label in ((1, "bit_parallel_batches"), (3, "run_length_bitmaps")):
    ctx.set_option("local_mode", mode)
    g.local_ints((lo, min(hi, lo + 64)))  # warm-up (row lists, workspace)
    t0 = time.time()
    cl, kk, tot, ctl = g.local_ints((lo, hi))
    ours[label] = {"wall_s": time.time() - t0, "kernel_ms": ctx.timing()["main_kernel_ms"]}
rp, col, _, _ = g.csr(bins=False)
edges = int(g.entries)
ctx.close()  # give the device memory back before the dense matrices are built

Ucols = slab if slab > 0 else U
Upad = (Ucols + 63) // 64 * 64
npad = (n + 63) // 64 * 64
c0 = 0 if slab <= 0 else max(0, U // 2 - slab // 2)
rows = torch.from_numpy(np.repeat(np.arange(n, dtype=np.int64), np.diff(rp).astype(np.int64))).to(dev)
cols = torch.from_numpy(col.astype(np.int64)).to(dev)
out = {"plan": name, "cells": hi - lo, "n": n, "universe": U, "edges": edges, "ours": ours, "column_slab": slab or None}
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
blk = 4096
for dtype, label in ((torch.int8, "int8_int32"), (torch.bfloat16, "bf16_fp32")):
    try:
        # right operand: A[u, w] for the filled middle vertices u and the columns w of the slab
        Ar = torch.zeros((npad, Upad), dtype=dtype, device=dev)
        keep = (cols >= c0) & (cols < c0 + Ucols)
        Ar[rows[keep], cols[keep] - c0] = 1
        # left operand rows: A[v, u] for the probed cells v, u < n
        ml = (hi - lo + 63) // 64 * 64
        Al = torch.zeros((ml, npad), dtype=dtype, device=dev)
        keepl = (rows >= lo) & (rows < hi) & (cols < n)
        Al[rows[keepl] - lo, cols[keepl]] = 1
        torch.cuda.synchronize()
        gemm_ms = 0.0
        res_cl, res_tot = [], []
        t0 = time.time()
        for b0 in range(0, hi - lo, blk):
            b1 = min(hi - lo, b0 + blk)
            m = (b1 - b0 + 63) // 64 * 64
            left = Al[b0:b0 + m].contiguous()
            for rep in range(2):  # the second run is the timed one (the first includes cuBLASLt's heuristics)
                ev0.record()
                P = torch._int_mm(left, Ar) if dtype == torch.int8 else torch.mm(left, Ar, out_dtype=torch.float32)
                ev1.record()
                ev1.synchronize()
            gemm_ms += ev0.elapsed_time(ev1)
            if slab <= 0:
                mask = Ar[lo + b0:lo + b1] != 0
                res_cl.append(torch.where(mask, P[:b1 - b0], torch.zeros((), dtype=P.dtype, device=dev)).sum(1, dtype=torch.float64).cpu().numpy())
                res_tot.append((P[:b1 - b0] > 0).sum(1).cpu().numpy())
                del mask
            del P
        torch.cuda.synchronize()
        wall = time.time() - t0
        flops = 2.0 * (hi - lo) * npad * Upad
        rec = {"gemm_ms": gemm_ms, "tera_ops_per_s": flops / (gemm_ms * 1e-3) / 1e12, "wall_s_with_torch_epilogue": wall}
        if slab <= 0:
            rec["results_identical"] = bool(np.array_equal(np.concatenate(res_cl).astype(np.int64), cl) and
                                            np.array_equal(np.concatenate(res_tot).astype(np.int32), tot))
        else:
            rec["gemm_ms_extrapolated_to_all_columns"] = gemm_ms * (U / Ucols)
        out[label] = rec
        del Ar, Al
        torch.cuda.empty_cache()
    except Exception as e:  # e.g. out of memory for the wider dtype
        out[label] = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
out["note"] = ("dense operands (N x columns elements); GEMMs are library kernels (cuBLASLt via torch); the elementwise mask / count "
               "epilogue is torch and not part of gemm_ms")
print(json.dumps(out))
