"""Measurement aid (not product code): the tensor-core formulation of the VGA local measures that SURVEY.md 7-K4 and the
north star ask to evaluate -- cluster = rowsum(A o (A.A)), total = nnz_row(A.A) with A the 0/1 adjacency as int8 --
timed through the LIBRARY int8 GEMM (torch._int_mm -> cuBLASLt, tensor cores) against the library's run-length kernel
(k_local_runs) on the same cells, results compared exactly.  A hand-written tcgen05 kernel could at best approach the
library GEMM on these dense tiles, so this bounds what the tensor-core path can win.

    python tools/local_tc_probe.py C1            # whole map
    python tools/local_tc_probe.py C4 8192       # a slice of 8192 cells (A stays dense: N^2 bytes)
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
flat = capi.prepare(plans.by_name(name))
ctx = capi.Context(0)
g = ctx.build(flat)
n, U = g.n, g.n + g.ghosts
lo = 0 if cells <= 0 else max(0, n // 2 - cells // 2)
hi = n if cells <= 0 else min(n, lo + cells)
dev = torch.device("cuda", 0)

# ours
g.local_ints((lo, min(hi, lo + 64)))  # warm-up (run lists)
torch.cuda.synchronize()
t0 = time.time()
cl, kk, tot, ctl = g.local_ints((lo, hi))
t_ours = time.time() - t0
tm = ctx.timing()

# dense int8 adjacency on the device: A[u, w] = 1 iff w in row u (rows: filled cells, columns: the universe incl. ghosts)
rp, col, _, _ = g.csr(bins=False)
Upad = (U + 63) // 64 * 64
npad = (n + 63) // 64 * 64
A = torch.zeros((npad, Upad), dtype=torch.int8, device=dev)
rows = torch.from_numpy(np.repeat(np.arange(n, dtype=np.int64), np.diff(rp).astype(np.int64))).to(dev)
A[rows, torch.from_numpy(col.astype(np.int64)).to(dev)] = 1
del rows
# 2-paths v -> u -> w need the middle vertex u to be a filled cell: left operand = columns < n of A
Al = torch.zeros((npad, npad), dtype=torch.int8, device=dev)
Al[:, :n] = A[:, :n]
torch.cuda.synchronize()
blk = 8192
res_cl, res_tot = [], []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gemm_ms = 0.0
t0 = time.time()
for b0 in range(lo, hi, blk):
    b1 = min(hi, b0 + blk)
    m = (b1 - b0 + 31) // 32 * 32
    left = Al[b0:b0 + m] if b0 + m <= npad else torch.cat([Al[b0:], torch.zeros((b0 + m - npad, npad), dtype=torch.int8, device=dev)])
    ev0.record()
    P = torch._int_mm(left.contiguous(), A)  # [m, Upad] int32: number of 2-paths
    ev1.record()
    ev1.synchronize()
    gemm_ms += ev0.elapsed_time(ev1)
    mask = A[b0:b1].to(torch.int32)
    res_cl.append((P[:b1 - b0] * mask).sum(1, dtype=torch.int64).cpu().numpy())
    res_tot.append((P[:b1 - b0] > 0).sum(1).cpu().numpy())
    del P, mask
torch.cuda.synchronize()
t_tc = time.time() - t0
tc_cl = np.concatenate(res_cl)
tc_tot = np.concatenate(res_tot)
ok = bool(np.array_equal(tc_cl, cl) and np.array_equal(tc_tot.astype(np.int32), tot))
flops = 2.0 * (hi - lo) * npad * Upad
print(json.dumps({"plan": name, "cells": hi - lo, "n": n, "universe": U, "edges": int(g.entries),
                  "ours_run_length_kernel_s": t_ours, "ours_kernel_ms": tm["main_kernel_ms"],
                  "tensor_core_library_gemm_ms": gemm_ms, "tensor_core_total_s_with_epilogue": t_tc,
                  "dense_int8_tflops": flops / (gemm_ms * 1e-3) / 1e12, "results_identical": ok,
                  "note": "A dense int8 (N^2 bytes); GEMM = torch._int_mm (cuBLASLt int8 tensor cores), epilogue = torch elementwise"}))
