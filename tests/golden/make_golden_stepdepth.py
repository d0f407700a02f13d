"""Golden vectors for visual step depth, from the UNMODIFIED reference's VGAVisualGlobalDepth::run
(oracle/_ref/libdmxref.so).  Run in the build container:  python tests/golden/make_golden_stepdepth.py
Adds tests/golden/stepdepth.npz: for three of the fixtures in this directory, source selections and the
"Visual Step Depth" column the reference wrote."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = {}
for name, sels in (("oblique20", [[0], [5, 100, 200]]), ("office24", [[17], [0, 575]]), ("oblique16s07", [[3, 400]])):
    fx = np.load(os.path.join(HERE, name + ".npz"))
    rm = po.RefMap(fx["walls"], float(fx["spacing"]))
    for s in fx["seeds"]:
        assert rm.fill(float(s[0]), float(s[1]))
    rm.makegraph()
    for i, sel in enumerate(sels):
        out[f"{name}__{i}__src"] = np.array(sel, np.int64)
        out[f"{name}__{i}__depth"] = rm.step_depth(sel)
np.savez_compressed(os.path.join(HERE, "stepdepth.npz"), **out)
print(sorted(out))
