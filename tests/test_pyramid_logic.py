"""CPU: the index logic of the OR-pyramid used by the bottom-up BFS step with run-length in-rows (bfs_pull = 1,
depthmapx_b200/csrc/pyramid.cuh): level layout, three-levels-at-once build, range decomposition.  The header is plain
C++ apart from the __host__ __device__ annotation, so the very functions the kernels call are compiled with g++ here and
checked against brute force."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("pyr") / "libpyrchk.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", "-o", so,
                           os.path.join(ROOT, "tests", "native", "pyramid_check.cpp")])
    L = C.CDLL(so)
    L.pyrchk_layout.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    L.pyrchk_updates.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.pyrchk_queries.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
    return L


def test_layout(lib):
    for n in [1, 2, 3, 7, 8, 9, 255, 256, 257, 65536, 1048576, 1000003]:
        cnt = np.zeros(34, np.int64)
        off = np.zeros(34, np.int64)
        total = C.c_int64()
        levels = lib.pyrchk_layout(n, cnt.ctypes.data, off.ctypes.data, C.addressof(total))
        assert cnt[0] == n and cnt[levels - 1] == 1
        for k in range(1, levels):
            assert cnt[k] == (cnt[k - 1] + 1) // 2
            assert off[k] == cnt[1:k].sum()
        assert total.value == cnt[1:levels].sum() and total.value <= n + levels


@pytest.mark.parametrize("w", [1, 2, 4])
def test_range_or_equals_brute_force(lib, w):
    rng = np.random.default_rng(w)
    for n in [1, 2, 3, 5, 8, 13, 64, 100, 257, 1000, 4099]:
        # sparse random frontier words so that a wrong node shows
        fr = np.where(rng.random((n, w)) < 0.05, rng.integers(1, 2 ** 63, (n, w), dtype=np.uint64), np.uint64(0)).astype(np.uint64)
        if n <= 64:  # every interval
            a, ln = np.array([(i, l) for i in range(n) for l in range(1, n - i + 1)], np.uint32).T
        else:
            a = rng.integers(0, n, 3000).astype(np.uint32)
            ln = np.minimum(rng.integers(1, 200, 3000), n - a).astype(np.uint32)
            a = np.append(a, [0, 0, n - 1]).astype(np.uint32)
            ln = np.append(ln, [n, 1, 1]).astype(np.uint32)
        a, ln = np.ascontiguousarray(a), np.ascontiguousarray(ln)
        out = np.zeros((len(a), w), np.uint64)
        loads = np.zeros(len(a), np.int32)
        assert lib.pyrchk_queries(w, fr.ctypes.data, n, a.ctypes.data, ln.ctypes.data, len(a), out.ctypes.data, loads.ctypes.data) == 0
        for q in range(len(a)):
            want = np.bitwise_or.reduce(fr[a[q]:a[q] + ln[q]], axis=0)
            assert np.array_equal(out[q], want), (n, int(a[q]), int(ln[q]))
            assert loads[q] <= 2 * max(1, int(np.ceil(np.log2(ln[q] + 1)))) + 2


@pytest.mark.parametrize("w", [1, 2, 4])
def test_range_or_updates_equal_brute_force(lib, w):
    """Top-down direction: words ORed into the nodes of a range, then the down pass (pyr_down_group, top chunk first)."""
    rng = np.random.default_rng(10 + w)
    for n in [1, 2, 3, 5, 8, 9, 64, 65, 100, 511, 512, 513, 1000, 4099]:
        nu = 1 if n < 4 else 200
        a = rng.integers(0, n, nu).astype(np.uint32)
        ln = np.minimum(rng.integers(1, max(2, n), nu), n - a).astype(np.uint32)
        word = rng.integers(1, 2 ** 63, (nu, w), dtype=np.uint64) & rng.integers(1, 2 ** 63, (nu, w), dtype=np.uint64)
        pre = np.where(rng.random((n, w)) < 0.1, rng.integers(1, 2 ** 63, (n, w), dtype=np.uint64), np.uint64(0)).astype(np.uint64)
        leaves = pre.copy()
        assert lib.pyrchk_updates(w, n, a.ctypes.data, ln.ctypes.data, word.ctypes.data, nu, leaves.ctypes.data) == 0
        want = pre.copy()
        for q in range(nu):
            want[a[q]:a[q] + ln[q]] |= word[q]
        assert np.array_equal(leaves, want), n
