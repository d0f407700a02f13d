"""In-tree build of the two shared libraries (no JIT cache, the .so files travel with gpurun).

  lib/libvga_b200.so  CUDA kernels + C ABI (include/vga_b200.h), sm_100a only
  lib/libvga_host.so  C++ host mirror of the reference interface (include/vga_host.h)
"""
from __future__ import annotations

import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "lib")
CU = ["makegraph", "bfs", "local", "local_tc", "stepdepth", "metric", "cabi"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> None:
    os.makedirs(LIB, exist_ok=True)
    hdrs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hdrs.append(os.path.join(HERE, "..", "include", "vga_b200.h"))

    def compile_one(name):
        src = os.path.join(CSRC, name + ".cu")
        obj = os.path.join(LIB, name + ".o")
        if force or _newer(obj, [src] + hdrs):
            r = subprocess.run(["nvcc"] + NVCC_FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
            with open(os.path.join(LIB, name + ".log"), "w") as f:
                f.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {name}.cu:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(compile_one, CU))
    so = os.path.join(LIB, "libvga_b200.so")
    if force or _newer(so, objs):
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", so] + objs)
    hsrc = [os.path.join(HOST, "pointmap.cpp"), os.path.join(HOST, "graphio.cpp"), os.path.join(HOST, "capi.cpp")]
    hdeps = hsrc + [os.path.join(HOST, "pointmap.h"), os.path.join(HOST, "graphfile.h"), os.path.join(HOST, "merge_contract.h"), os.path.join(HOST, "geometry.h"),
                    os.path.join(HERE, "..", "include", "vga_host.h"), so]
    hso = os.path.join(LIB, "libvga_host.so")
    if force or _newer(hso, hdeps):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", hso] + hsrc +
                              ["-L" + LIB, "-lvga_b200", "-Wl,-rpath,$ORIGIN"])
    if verbose:
        print("built", so, hso)


if __name__ == "__main__":
    build(force=True, verbose=True)
