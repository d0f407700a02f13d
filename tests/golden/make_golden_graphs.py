"""Generate tests/golden/graphfiles.npz: .graph files written by the UNMODIFIED reference CLI
(oracle/_ref/depthmapXcli_ref, built by integration/Makefile from /root/reference) for small plans.  Run in the
build container only:

    python tests/golden/make_golden_graphs.py

Per plan: plan (IMPORT of the wall CSV), fill (VISPREP -pg -pp), prep (… -pm), prep_pb (… -pm -pb), vga (VGA -vm
visibility -vg -vl -vr n on prep), vga3 (-vg -vr 3 on vga), sd (STEPDEPTH -sdp … -sdt visual on prep); with merge links
(SURVEY §8 f3): prep_l (LINK -lnk … on prep), vga_l (-vg -vl -vr n), vga_l2 (-vg -vr 2), sd_l (STEPDEPTH), all on prep_l.  Each file is
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


def pick_links(path, seed, pairs=2):
    """Two pairs of distinct filled cells of the map in `path`, as -lnk arguments (cell centres)."""
    from depthmapx_b200 import capi
    m = capi.GraphFile(path).map()
    st = m.state().reshape(m.cols, m.rows)
    filled = np.argwhere((st & 2) != 0)
    rng = np.random.default_rng(seed)
    cells = filled[rng.choice(len(filled), 2 * pairs, replace=False)]
    pt = lambda c: f"{m.bl_x + c[0] * m.spacing:.10g},{m.bl_y + c[1] * m.spacing:.10g}"
    return [pt(cells[2 * i]) + "," + pt(cells[2 * i + 1]) for i in range(pairs)]


def run(args, cwd):
    r = subprocess.run([REF] + args, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, (args, r.stdout, r.stderr)


def make_case(d, plan, grid, seed, sdp, link_seed):
    open(os.path.join(d, "walls.csv"), "w").write(plan.csv())
    run(["-m", "IMPORT", "-f", "walls.csv", "-o", "plan.graph", "-it", "drawing"], d)
    run(["-m", "VISPREP", "-f", "plan.graph", "-o", "fill.graph", "-pg", grid, "-pp", seed], d)
    run(["-m", "VISPREP", "-f", "plan.graph", "-o", "prep.graph", "-pg", grid, "-pp", seed, "-pm"], d)
    run(["-m", "VISPREP", "-f", "plan.graph", "-o", "prep_pb.graph", "-pg", grid, "-pp", seed, "-pm", "-pb"], d)
    run(["-m", "VGA", "-f", "prep.graph", "-o", "vga.graph", "-vm", "visibility", "-vg", "-vl", "-vr", "n"], d)
    run(["-m", "VGA", "-f", "vga.graph", "-o", "vga3.graph", "-vm", "visibility", "-vg", "-vr", "3"], d)
    run(["-m", "STEPDEPTH", "-f", "prep.graph", "-o", "sd.graph", "-sdp", sdp, "-sdt", "visual"], d)
    links = pick_links(os.path.join(d, "prep.graph"), link_seed)
    lnk = []
    for l in links:
        lnk += ["-lnk", l]
    run(["-m", "LINK", "-f", "prep.graph", "-o", "prep_l.graph"] + lnk, d)
    run(["-m", "VGA", "-f", "prep_l.graph", "-o", "vga_l.graph", "-vm", "visibility", "-vg", "-vl", "-vr", "n"], d)
    run(["-m", "VGA", "-f", "prep_l.graph", "-o", "vga_l2.graph", "-vm", "visibility", "-vg", "-vr", "2"], d)
    run(["-m", "STEPDEPTH", "-f", "prep_l.graph", "-o", "sd_l.graph", "-sdp", sdp, "-sdt", "visual"], d)
    out = {k: np.frombuffer(open(os.path.join(d, k + ".graph"), "rb").read(), np.uint8)
           for k in ("plan", "fill", "prep", "prep_pb", "vga", "vga3", "sd", "prep_l", "vga_l", "vga_l2", "sd_l")}
    return out, links


def main():
    out = {}
    for name, (spec, grid, seed, sdp) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            made, links = make_case(d, plans.by_name(spec), grid, seed, sdp, len(out))
            for k, v in made.items():
                out[f"{name}__{k}"] = v
        out[f"{name}__args"] = np.array([spec, grid, seed, sdp] + links)
    path = os.path.join(ROOT, "tests", "golden", "graphfiles.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
