"""CPU: the CUDA translation units of libvga_b200 executed by a SIMT emulator (tests/emu/: kernel launches become
fiber-per-thread block executions with real warp collectives; see tests/emu/include/simt_emu.h), so that kernel logic and
the host-side launch sequences are exercised without a GPU.  It keeps every kernel under test in the GPU-less container (round 2's BFS rewrite around pyramid node
lists and the O(runs) transposition were developed against it before their first B200 run).

GPU test modules are re-run unchanged in a subprocess whose ctypes loader points at the emulation build
(VGA_EMU_LIBDIR, tests/conftest.py).  The emulator is single-threaded: it checks logic (indices, masks, collectives, launch order), not
races or performance.  Nothing in the product can load it."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu")


@pytest.fixture(scope="module")
def emu_dir():
    sys.path.insert(0, EMU)
    try:
        import build_emu
    finally:
        sys.path.pop(0)
    return build_emu.build()


def run_gpu_tests_emulated(emu_dir, args, timeout=1500):
    env = dict(os.environ, VGA_EMU_LIBDIR=emu_dir, LD_LIBRARY_PATH=emu_dir + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "--runxfail", "-x", "-q", "-p", "no:cacheprovider"] + args,
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-25:])
    assert r.returncode == 0, tail
    return tail


def test_emulator_semantics(emu_dir, tmp_path):
    """The emulator itself: shuffles, ballots, sub-warp masks, reductions, __syncthreads, shared memory, exited lanes."""
    src = tmp_path / "selftest.cpp"
    src.write_text(r'''
#include <cuda_runtime.h>
#include <vector>
__global__ void k(int n, const int *in, int *out, int *blocksum) {
    __shared__ int s[8];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (blockIdx.y == 1 && warp == 1) return;                     // a whole warp exits before the block barrier
    int v = i < n ? in[i] : 0;
    unsigned odd = __ballot_sync(0xffffffffu, v & 1);              // full-warp ballot
    int left = __shfl_up_sync(0xffffffffu, v, 1);                  // neighbour
    int bcast = __shfl_sync(0xffffffffu, v, 5);                    // broadcast of lane 5
    unsigned gmask = 0xffu << ((lane >> 3) * 8);                   // 8-lane groups
    unsigned gor = __reduce_or_sync(gmask, (unsigned)v);
    int sum = v;
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0) s[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (unsigned w = 0; w < blockDim.x / 32; w++) if (!(blockIdx.y == 1 && w == 1)) t += s[w];
        blocksum[blockIdx.y * gridDim.x + blockIdx.x] = t;
    }
    if (i < n && blockIdx.y == 0) out[i] = (int)__popc(odd) * 1000000 + (lane ? left : -1) * 0 + bcast * 1000 + (int)(gor & 0xff) + (lane ? left : 0) * 0;
}
extern "C" int selftest() {
    const int n = 256;
    std::vector<int> in(n), out(n, 0), bs(4, 0);
    for (int i = 0; i < n; i++) in[i] = (i * 7 + 3) % 251;
    simt::launch(dim3(2, 2), dim3(128), 0, [&] { k(n, in.data(), out.data(), bs.data()); });
    for (int b = 0; b < 2; b++) {
        int want0 = 0, want1 = 0;
        for (int t = 0; t < 128; t++) {
            want0 += in[b * 128 + t];
            if (t / 32 != 1) want1 += in[b * 128 + t];
        }
        if (bs[b] != want0 || bs[2 + b] != want1) return 1;
    }
    for (int i = 0; i < n; i++) {
        int w0 = i & ~31, g0 = i & ~7, odd = 0, gor = 0;
        for (int l = 0; l < 32; l++) odd += in[w0 + l] & 1;
        for (int l = 0; l < 8; l++) gor |= in[g0 + l];
        if (out[i] != odd * 1000000 + in[w0 + 5] * 1000 + (gor & 0xff)) return 2;
    }
    return 0;
}
''')
    so = str(tmp_path / "selftest.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I" + os.path.join(EMU, "include"), "-o", so, str(src)])
    import ctypes
    assert ctypes.CDLL(so).selftest() == 0


def test_default_paths_small_plans(emu_dir):
    """Validated defaults re-checked under emulation (guards against emulator drift): goldens of the reference, shards,
    ghost cells, empty inputs, host layer."""
    run_gpu_tests_emulated(emu_dir, ["tests/test_gpu_parity.py", "-k",
                                     "golden or ghost or empty or single or shard or host_layer or boundary or maxdist or several_chunks "
                                     "or (local_vs_oracle and oblique)"])


def test_step_depth_kernel(emu_dir):
    run_gpu_tests_emulated(emu_dir, ["tests/test_stepdepth_gpu.py"])


def test_metric_angular_kernel(emu_dir):
    """tests/test_gpu_metric.py on the smallest plan: the warp-per-source search with the 32-ary indexed heap (metric and
    angular, whole map, permuted source lists, one CTA of slots vs many, error paths) against the oracle."""
    run_gpu_tests_emulated(emu_dir, ["tests/test_gpu_metric.py", "-k", "whole_map or errors"])


def test_graphfile_pipelines_merges_and_cli_shim(emu_dir):
    """tests/test_zzz_graphfile_gpu.py: VISPREP / VGA / STEPDEPTH / LINK pipelines file to file and the real depthmapXcli
    with the step-depth shim, byte-compared with the reference CLI's files."""
    run_gpu_tests_emulated(emu_dir, ["tests/test_zzz_graphfile_gpu.py"])


def test_bfs_schedules_and_row_ordering(emu_dir):
    """tests/test_gpu_bfs_schedules.py on the small plans: the three directions, every word width, explicit source lists,
    the O(runs) transposition (pinned by the bottom-up-only schedule), both row-ordering kernels, the run-length round trip."""
    run_gpu_tests_emulated(emu_dir, ["tests/test_gpu_bfs_schedules.py", "-k", "oblique or holes or round_trip"])


def test_bfs_schedules_on_a_plan_with_rooms(emu_dir):
    """office:64 (4,096 cells, deeper level structure, several batches): hybrid with automatic word width, bottom-up only
    with two words, top-down only with four."""
    f = "tests/test_gpu_bfs_schedules.py::test_schedules_vs_oracle"
    run_gpu_tests_emulated(emu_dir, [f + "[2-0-2--1-office:64:64:1]", f + "[1-2-2--1-office:64:64:1]", f + "[0-4-2-2-office:64:64:1]"])


def test_real_cli_with_gpu_shims(emu_dir):
    """The real depthmapXcli with the shims of integration/ (one plan; the B200 run covers all three)."""
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_gpu")):
        pytest.skip("integration binaries not built")
    run_gpu_tests_emulated(emu_dir, ["tests/test_cli_dropin.py", "-k", "oblique:20"])


@pytest.mark.skipif(not os.environ.get("VGA_EMU_ASAN"), reason="slow: set VGA_EMU_ASAN=1 (AddressSanitizer build of the emulation)")
def test_kernels_under_address_sanitizer():
    """Device memory is the host heap in the emulation, so an AddressSanitizer build catches out-of-bounds accesses and
    reads of uninitialised-by-luck memory (ASan poisons fresh allocations) inside kernels.  Minutes, hence opt-in."""
    sys.path.insert(0, EMU)
    try:
        import build_emu
    finally:
        sys.path.pop(0)
    d = build_emu.build(sanitize=True)
    libs = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip() + ":" + \
        subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    env = dict(os.environ, VGA_EMU_LIBDIR=d, LD_LIBRARY_PATH=d, LD_PRELOAD=libs,
               ASAN_OPTIONS="detect_leaks=0:detect_stack_use_after_return=0")
    r = subprocess.run([sys.executable, "-m", "pytest", "-m", "gpu", "--runxfail", "-x", "-q", "-p", "no:cacheprovider",
                        "tests/test_gpu_bfs_schedules.py", "tests/test_gpu_parity.py", "tests/test_stepdepth_gpu.py",
                        "tests/test_zzz_graphfile_gpu.py", "-k",
                        "(oblique and not overflow) or golden or ghost or empty or several or step_depth or visprep or loaded or merge"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=7200)
    tail = "\n".join((r.stdout + r.stderr).splitlines()[-30:])
    assert r.returncode == 0 and "AddressSanitizer" not in r.stdout + r.stderr.replace("ASan doesn't fully support", ""), tail
