"""Per-kernel device time and DRAM bytes of an ncu launch list (CSV of `ncu --metrics gpu__time_duration.sum,
dram__bytes_read.sum,dram__bytes_write.sum --csv`), and the measured DRAM traffic per BFS source of the level kernels ->
profiles/r2_bfs_traffic.json (read by bench.py for `roofline.traffic`).

    python tools/bfs_traffic.py launches.csv C5 8192 [out.json ...]
"""
import collections
import csv
import json
import os
import re
import sys

LEVEL = ("k_push", "k_pyr_down", "k_pyr_build", "k_pull", "k_update", "k_decide")


def aggregate(path):
    rows = list(csv.reader(open(path, errors="replace")))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ix = {h: i for i, h in enumerate(rows[start])}
    t, cnt, rd, wr = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float), collections.defaultdict(float)
    tscale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}
    bscale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for r in rows[start + 1:]:
        if len(r) < len(ix):
            continue
        k = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("vga::<unnamed>::", "")
        m, u = r[ix["Metric Name"]], r[ix["Metric Unit"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        if m == "gpu__time_duration.sum":
            t[k] += v * tscale.get(u, 1e-6)
            cnt[k] += 1
        elif m == "dram__bytes_read.sum":
            rd[k] += v * bscale.get(u, 1.0)
        elif m == "dram__bytes_write.sum":
            wr[k] += v * bscale.get(u, 1.0)
    return t, cnt, rd, wr


def main():
    path, workload, nsrc = sys.argv[1], sys.argv[2], int(sys.argv[3])
    t, cnt, rd, wr = aggregate(path)
    tot = sum(t.values())
    for k, v in sorted(t.items(), key=lambda x: -x[1]):
        print(f"{k[:58]:58s} {cnt[k]:5d} launches {v:9.2f} ms {100 * v / tot:5.1f} %  read {rd[k] / 1e9:8.2f} GB  written {wr[k] / 1e9:8.2f} GB")
    lv = [k for k in t if k.startswith(LEVEL)]
    ms = sum(t[k] for k in lv)
    by = sum(rd[k] + wr[k] for k in lv)
    print(f"level kernels: {ms:.2f} ms of launch time, {by / 1e9:.2f} GB of DRAM traffic, {by / nsrc / 1e6:.3f} MB per source")
    rec = {workload: {"dram_bytes_per_source": by / nsrc, "sources_in_capture": nsrc,
                      "from": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum over every launch of "
                              f"vga_global (radius n, one call) on a {workload} slice of {nsrc} sources with the default kernels "
                              f"({os.path.basename(path)}), level kernels only: {by / 1e9:.1f} GB in {ms:.1f} ms of launch time",
                      "share_of_launch_time": {k: round(t[k] / ms, 4) for k in sorted(lv, key=lambda k: -t[k])}}}
    for out in sys.argv[4:]:
        old = {}
        if os.path.exists(out):
            try:
                old = json.load(open(out))
            except Exception:
                old = {}
        old.update(rec)
        json.dump(old, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
