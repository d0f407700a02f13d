"""Generate tests/golden/graphfiles_metric.npz: .graph files written by the UNMODIFIED reference CLI
(oracle/_ref/depthmapXcli_ref) for the metric / angular analyses on the prep / prep_l (merge links)
files of tests/golden/graphfiles.npz: metric (VGA -vm metric -vr n), metric_r (-vr 4 cells), angular (VGA -vm angular),
angular2 (-vm angular run again on the angular file: getOrInsertColumn keeps the columns), metric_l / angular_l on the
linked maps.  Run in the build container only:

    python tests/golden/make_golden_graphs_metric.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_ref")
CASES = ["oblique12", "oblique10s07", "office16"]


def run(args, cwd):
    r = subprocess.run([REF] + args, cwd=cwd, capture_output=True, text=True)
    assert r.returncode == 0, (args, r.stdout, r.stderr)


def main():
    fx = np.load(os.path.join(ROOT, "tests", "golden", "graphfiles.npz"), allow_pickle=False)
    out = {}
    for case in CASES:
        grid = float(str(fx[case + "__args"][1]))
        with tempfile.TemporaryDirectory() as d:
            for k in ("prep", "prep_l"):
                open(os.path.join(d, k + ".graph"), "wb").write(fx[f"{case}__{k}"].tobytes())
            radius = "%g" % (4 * grid)
            run(["-m", "VGA", "-f", "prep.graph", "-o", "metric.graph", "-vm", "metric", "-vr", "n"], d)
            run(["-m", "VGA", "-f", "prep.graph", "-o", "metric_r.graph", "-vm", "metric", "-vr", radius], d)
            run(["-m", "VGA", "-f", "prep.graph", "-o", "angular.graph", "-vm", "angular"], d)
            run(["-m", "VGA", "-f", "angular.graph", "-o", "angular2.graph", "-vm", "angular"], d)
            run(["-m", "VGA", "-f", "prep_l.graph", "-o", "metric_l.graph", "-vm", "metric", "-vr", "n"], d)
            run(["-m", "VGA", "-f", "prep_l.graph", "-o", "angular_l.graph", "-vm", "angular"], d)
            for k in ("metric", "metric_r", "angular", "angular2", "metric_l", "angular_l"):
                out[f"{case}__{k}"] = np.frombuffer(open(os.path.join(d, k + ".graph"), "rb").read(), np.uint8)
            out[f"{case}__radius"] = np.array(radius)
    path = os.path.join(ROOT, "tests", "golden", "graphfiles_metric.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
