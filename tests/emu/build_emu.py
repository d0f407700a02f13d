"""TEST INFRASTRUCTURE ONLY: builds a CPU emulation of libvga_b200.so (and the host layer on top of it) for
`pytest -m "not gpu"` runs of kernel logic.  See tests/emu/include/simt_emu.h for the execution model.

    python tests/emu/build_emu.py [outdir]        # default tests/emu/_build (git-ignored)

Steps: every depthmapx_b200/csrc/*.cu is copied with two textual changes -- `kernel<<<grid, block, smem, stream>>>(args)`
becomes `simt::launch(grid, block, smem, [&] { kernel(args); })` and `extern __shared__ T name[];` becomes a pointer to
the launch's dynamic shared memory -- and compiled with g++ against tests/emu/include (stand-ins for <cuda_runtime.h> and
<cub/cub.cuh>).  Kernel bodies, device functions and the host-side launch sequences are compiled unchanged.
The product never loads the result: the tests point their ctypes loader at the scratch directory explicitly.
"""
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "depthmapx_b200", "csrc")
HOST = os.path.join(ROOT, "depthmapx_b200", "host")
CU = ["makegraph", "bfs", "local", "local_tc", "stepdepth", "metric", "cabi"]


def _match_back_angle(s, i):
    """s[i] == '>': index of the matching '<' scanning backwards."""
    depth = 0
    while i >= 0:
        if s[i] == '>':
            depth += 1
        elif s[i] == '<':
            depth -= 1
            if depth == 0:
                return i
        i -= 1
    raise ValueError("unbalanced template arguments before <<<")


def _match_paren(s, i):
    """s[i] == '(': index of the matching ')'."""
    depth = 0
    while i < len(s):
        if s[i] == '(':
            depth += 1
        elif s[i] == ')':
            depth -= 1
            if depth == 0:
                return i
        i += 1
    raise ValueError("unbalanced kernel arguments")


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def preprocess(src: str) -> str:
    out, pos, n = [], 0, 0
    while True:
        k = src.find("<<<", pos)
        if k < 0:
            break
        # kernel expression: identifier [<template args>] right before <<<
        j = k - 1
        if src[j] == '>':
            j = _match_back_angle(src, j) - 1
        while j >= 0 and (src[j].isalnum() or src[j] in "_:"):
            j -= 1
        start = j + 1
        kernel = src[start:k]
        e = src.find(">>>", k)
        cfg = _split_top(src[k + 3:e])
        assert len(cfg) == 4, f"launch configuration {cfg}"
        a0 = e + 3
        assert src[a0] == "(", src[a0 - 20:a0 + 20]
        a1 = _match_paren(src, a0)
        args = src[a0 + 1:a1]
        out.append(src[pos:start])
        out.append(f"simt::launch({cfg[0]}, {cfg[1]}, {cfg[2]}, [&] {{ {kernel}({args}); }})")
        pos = a1 + 1
        n += 1
    out.append(src[pos:])
    s = "".join(out)
    s = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([\w ]+?)\s+(\w+)\[\];", r"\1 *\2 = (\1 *)simt::dyn_smem();", s)
    return s


def build(outdir=None, force=False, sanitize=False):
    """sanitize=True: AddressSanitizer + UBSan build (device memory is the host heap, so out-of-bounds accesses of a kernel
    are caught); uses the ucontext fiber switch, which ASan understands."""
    outdir = outdir or os.path.join(HERE, "_build_asan" if sanitize else "_build")
    os.makedirs(outdir, exist_ok=True)
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HOST, f) for f in os.listdir(HOST)] + \
           [os.path.join(HERE, "include", "simt_emu.h"), os.path.abspath(__file__), os.path.join(ROOT, "include", "vga_b200.h")]
    so = os.path.join(outdir, "libvga_b200.so")
    hso = os.path.join(outdir, "libvga_host.so")
    newest = max(os.path.getmtime(p) for p in srcs)
    if not force and os.path.exists(so) and os.path.exists(hso) and min(os.path.getmtime(so), os.path.getmtime(hso)) > newest:
        return outdir
    san = ["-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-DSIMT_USE_UCONTEXT"] if sanitize else []
    flags = ["-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-I" + os.path.join(HERE, "include"), "-I" + CSRC] + san

    def one(name):
        cpp = os.path.join(outdir, name + ".emu.cpp")
        open(cpp, "w").write(preprocess(open(os.path.join(CSRC, name + ".cu")).read()))
        obj = os.path.join(outdir, name + ".o")
        r = subprocess.run(["g++"] + flags + ["-c", cpp, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"emulation build of {name}.cu failed:\n{r.stderr[:6000]}")
        return obj

    with ThreadPoolExecutor(max_workers=5) as ex:
        objs = list(ex.map(one, CU))
    subprocess.check_call(["g++", "-shared", "-o", so] + san + objs)
    hsrc = [os.path.join(HOST, f) for f in ("pointmap.cpp", "graphio.cpp", "capi.cpp")]
    subprocess.check_call(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", hso] + san + hsrc +
                          ["-L" + outdir, "-lvga_b200", "-Wl,-rpath,$ORIGIN"])
    return outdir


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a != "--asan"]
    print(build(args[0] if args else None, force=True, sanitize="--asan" in sys.argv))
