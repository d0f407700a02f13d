"""File-to-file example of the stand-alone host layer (no reference code involved): what
    depthmapXcli -m VISPREP -f plan.graph -o prep.graph -pg 1 -pp 1,1 -pm
    depthmapXcli -m VGA     -f prep.graph -o vga.graph  -vm visibility -vg -vl -vr n
    depthmapXcli -m STEPDEPTH -f prep.graph -o sd.graph -sdp 3,3 -sdt visual
do, with makegraph / BFS / local measures on the B200 (needs a CUDA device: there is no CPU path).

    python examples/graph_file_pipeline.py plan.graph out_dir [grid spacing] [seed x,y]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from depthmapx_b200 import capi  # noqa: E402


def main():
    src, out = sys.argv[1], sys.argv[2]
    spacing = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    seed = [float(v) for v in (sys.argv[4] if len(sys.argv) > 4 else "1,1").split(",")]
    os.makedirs(out, exist_ok=True)

    g = capi.GraphFile(src)                      # drawing layers -> wall segments
    m = g.new_map(spacing)                       # MetaGraph::addNewPointMap + setGrid
    assert m.fill(*seed), "seed outside the plan or on a wall"
    g.make_graph()                               # PointMap::sparkGraph2 on the GPU
    g.save(os.path.join(out, "prep.graph"))
    print(f"prep.graph: {m.n} cells, columns {m.columns()}")

    g = capi.GraphFile(os.path.join(out, "prep.graph"))   # adjacency is uploaded from the file's Nodes
    m = g.map()
    m.vga_local()
    m.vga_global(-1.0)
    g.save(os.path.join(out, "vga.graph"))
    print(f"vga.graph: mean Visual Integration [HH] = {m.attr('Visual Integration [HH]').mean():.4f}")

    g = capi.GraphFile(os.path.join(out, "prep.graph"))
    g.map().step_depth([seed])
    g.save(os.path.join(out, "sd.graph"))
    print("sd.graph written")


if __name__ == "__main__":
    main()
