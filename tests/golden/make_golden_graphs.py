"""Generate tests/golden/graphfiles.npz: .graph files written by the UNMODIFIED reference CLI
(oracle/_ref/depthmapXcli_ref, built by integration/Makefile from /root/reference) for small plans.  Run in the
build container only:

    python tests/golden/make_golden_graphs.py

Per plan: plan (IMPORT of the wall CSV), fill (VISPREP -pg -pp), prep (… -pm), prep_pb (… -pm -pb), vga (VGA -vm
visibility -vg -vl -vr n on prep), vga3 (-vg -vr 3 on vga), sd (STEPDEPTH -sdp … -sdt visual on prep).  Each file is
stored as a uint8 array; tests/test_graphfile.py reads them with the host layer's own .graph codec.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from depthmapx_b200 import plans  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_ref")
CASES = {"oblique12": ("oblique:12:12:3", "1", "1,1", "3,3"), "oblique10s07": ("oblique:10:10:5:0.7", "0.7", "0.7,0.7", "2.1,2.1"),
         "office16": ("office:16:16:1", "1", "1,1", "2,9")}


def run(args, cwd):
    r = subprocess.run([REF] + args, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, (args, r.stdout, r.stderr)


def make_case(d, plan, grid, seed, sdp):
    open(os.path.join(d, "walls.csv"), "w").write(plan.csv())
    run(["-m", "IMPORT", "-f", "walls.csv", "-o", "plan.graph", "-it", "drawing"], d)
    run(["-m", "VISPREP", "-f", "plan.graph", "-o", "fill.graph", "-pg", grid, "-pp", seed], d)
    run(["-m", "VISPREP", "-f", "plan.graph", "-o", "prep.graph", "-pg", grid, "-pp", seed, "-pm"], d)
    run(["-m", "VISPREP", "-f", "plan.graph", "-o", "prep_pb.graph", "-pg", grid, "-pp", seed, "-pm", "-pb"], d)
    run(["-m", "VGA", "-f", "prep.graph", "-o", "vga.graph", "-vm", "visibility", "-vg", "-vl", "-vr", "n"], d)
    run(["-m", "VGA", "-f", "vga.graph", "-o", "vga3.graph", "-vm", "visibility", "-vg", "-vr", "3"], d)
    run(["-m", "STEPDEPTH", "-f", "prep.graph", "-o", "sd.graph", "-sdp", sdp, "-sdt", "visual"], d)
    return {k: np.frombuffer(open(os.path.join(d, k + ".graph"), "rb").read(), np.uint8)
            for k in ("plan", "fill", "prep", "prep_pb", "vga", "vga3", "sd")}


def main():
    out = {}
    for name, (spec, grid, seed, sdp) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            for k, v in make_case(d, plans.by_name(spec), grid, seed, sdp).items():
                out[f"{name}__{k}"] = v
        out[f"{name}__args"] = np.array([spec, grid, seed, sdp])
    path = os.path.join(ROOT, "tests", "golden", "graphfiles.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
