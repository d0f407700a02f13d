"""End-to-end CLI timing rows: the real depthmapXcli with the GPU shims (oracle/_ref/depthmapXcli_gpu) next
to the unmodified reference CLI (oracle/_ref/depthmapXcli_ref), both with `-t times.csv`, rows "Making graph"
(depthmapXcli/runmethods.cpp:334) and "Run VGA" (:263).  Runs on the GPU box; the outputs of the two CLIs are byte-compared.

    python tools/cli_timing.py [C1 [C2 ...]] > gpurun_out/cli_timing.json

The reference is single-threaded: its VGA global on C1 (10,000 cells) takes ~5 minutes, on C2 over an hour, and its local
measures are slower still (16 s on a 24 x 24 office), so the paired rows are VISPREP + `-vg` for C1, VISPREP only for C2
and everything for small plans; the GPU CLI additionally runs the unpaired modes."""
import csv
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depthmapx_b200 import plans  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_ref")
GPU = os.path.join(ROOT, "oracle", "_ref", "depthmapXcli_gpu")


def run(binary, args, cwd):
    t0 = time.time()
    r = subprocess.run([binary] + args, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{os.path.basename(binary)} {' '.join(args)}\n{r.stdout}\n{r.stderr}")
    return time.time() - t0


def times(path):
    out = {}
    with open(path) as f:
        for row in csv.reader(f):
            if len(row) >= 2:
                try:
                    out[row[0].strip('"')] = float(row[1])
                except ValueError:
                    pass
    return out


def same(a, b):
    return open(a, "rb").read() == open(b, "rb").read()


def one(name, paired, unpaired):
    """paired / unpaired: lists of VGA argument lists run by both CLIs / by the GPU CLI only."""
    plan = plans.by_name(name)
    rec = {"plan": plan.name, "rows": []}
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "walls.csv"), "w").write(plan.csv())
        run(REF, ["-m", "IMPORT", "-f", "walls.csv", "-o", "plan.graph", "-it", "drawing"], d)
        seed = f"{plan.seeds[0][0]},{plan.seeds[0][1]}"
        for tag, binary in (("ref", REF), ("gpu", GPU)):
            wall = run(binary, ["-m", "VISPREP", "-f", "plan.graph", "-o", f"prep_{tag}.graph", "-pg", str(plan.spacing), "-pp", seed,
                                "-pm", "-t", f"t_prep_{tag}.csv"], d)
            t = times(os.path.join(d, f"t_prep_{tag}.csv"))
            rec["rows"].append({"cli": tag, "mode": "VISPREP -pm", "Making graph": t.get("Making graph"), "process_wall_s": wall})
        rec["visprep_identical"] = same(os.path.join(d, "prep_ref.graph"), os.path.join(d, "prep_gpu.graph"))
        for i, vga in enumerate(paired + unpaired):
            both = i < len(paired)
            for tag, binary in ((("ref", REF), ("gpu", GPU)) if both else (("gpu", GPU),)):
                wall = run(binary, ["-m", "VGA", "-f", "prep_ref.graph", "-o", f"vga_{i}_{tag}.graph", "-vm", "visibility"] + vga +
                           ["-t", f"t_vga_{i}_{tag}.csv"], d)
                t = times(os.path.join(d, f"t_vga_{i}_{tag}.csv"))
                rec["rows"].append({"cli": tag, "mode": "VGA -vm visibility " + " ".join(vga), "Run VGA": t.get("Run VGA"),
                                    "process_wall_s": wall})
            if both:
                rec["identical: " + " ".join(vga)] = same(os.path.join(d, f"vga_{i}_ref.graph"), os.path.join(d, f"vga_{i}_gpu.graph"))
    return rec


if __name__ == "__main__":
    names = sys.argv[1:] or ["C1"]
    out = {"host_cores": os.cpu_count(), "reference_threads": 1, "plans": {}}
    G, L = ["-vg", "-vr", "n"], ["-vl"]
    for name in names:
        if name == "C1":
            out["plans"][name] = one(name, [G], [L, ["-vg", "-vl", "-vr", "n"], ["-vg", "-vr", "3"]])
        elif name in ("C2", "C3", "C4", "C5"):
            out["plans"][name] = one(name, [], [G, L])
        else:
            out["plans"][name] = one(name, [G, L, ["-vg", "-vl", "-vr", "3"]], [])
    print(json.dumps(out, indent=1))
