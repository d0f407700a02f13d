"""Copies the INPUT files of the reference's own regression cases for this path into tests/golden/regression/
(the GPU box has no /root/reference).  Test data only -- no reference source.

Cases (RegressionTest/regressionconfig.json): pointmap_create_fill_make_one_operation (gallery_empty.graph),
dense_pointmap_create_fill_make (rect1x1.graph), visibility_global_n / visibility_global_3 / visibility_local /
vga_visual_step_depth (gallery_connected.graph).  The expected outputs are not stored: tests/test_cli_dropin.py runs
the unmodified reference CLI (oracle/_ref/depthmapXcli_ref) next to the CLI with the GPU shims and byte-diffs the
two .graph files, the reference's own regression method (RegressionTest/depthmaprunner.py:19-79).

    python tests/golden/make_regression_fixtures.py
"""
import gzip
import os
import shutil

SRC = "/root/reference/testdata"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "regression")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for name in ("gallery_empty.graph", "rect1x1.graph"):
        shutil.copyfile(os.path.join(SRC, name), os.path.join(DST, name))
    # 3.7 MB / 0.6 MB -> stored gzip-compressed with a fixed mtime so the files are reproducible
    # (turns_connected.graph: input of the vga_metric / vga_angular regression cases)
    for name in ("gallery_connected.graph", "turns_connected.graph"):
        with open(os.path.join(SRC, name), "rb") as f, \
                open(os.path.join(DST, name + ".gz"), "wb") as raw, \
                gzip.GzipFile(filename="", mode="wb", fileobj=raw, mtime=0) as g:
            g.write(f.read())
    for f in sorted(os.listdir(DST)):
        print(f, os.path.getsize(os.path.join(DST, f)))
