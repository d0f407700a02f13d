// Shared helpers of the shim translation units that put libvga_b200.so behind the reference's
// own entry points.  These files are compiled INSIDE the reference tree's
// include path (they are members of reference classes); they hold no reference code.
#pragma once

#include <cstdlib>
#include <string>
#include <vector>

#include "genlib/exceptions.h"
#include "salalib/pointdata.h"
#include "vga_b200.h"
#include "merge_contract.h"  // depthmapx_b200/host: contraction of merged pairs, shared with the stand-alone host layer

namespace vga_shim {

inline vga_ctx *gpu() {  // one process drives one GPU
    static vga_ctx *c = nullptr;
    if (!c) {
        int dev = 0;
        if (const char *e = std::getenv("VGA_DEVICE")) dev = std::atoi(e);
        if (vga_ctx_create(dev, &c) != VGA_OK)
            throw depthmapX::RuntimeException(std::string("GPU path: ") + vga_last_error());  // no CPU fallback
    }
    return c;
}

struct CommState {
    Communicator *comm;
    time_t atime;
};
inline void progress_cb(void *u, int64_t done, int64_t) {
    CommState *s = static_cast<CommState *>(u);
    if (s->comm) s->comm->CommPostMessage(Communicator::CURRENT_RECORD, (int)done);
}
inline int cancel_cb(void *u) {
    CommState *s = static_cast<CommState *>(u);
    return (s->comm && qtimer(s->atime, 500) && s->comm->IsCancelled()) ? 1 : 0;
}

// x-major list of filled cells and the cell -> ordinal map (unfilled cells: N + rank among unfilled)
struct Ordinals {
    std::vector<PixelRef> cells;   // filled, x-major
    std::vector<int32_t> ord;      // [cols*rows]
    std::vector<uint8_t> skip;     // per filled cell: context-filled and not even (empty when there is none)
    int64_t n = 0, ghosts = 0;
};
inline Ordinals make_ordinals(PointMap &map) {
    Ordinals o;
    const size_t cols = map.getCols(), rows = map.getRows();
    o.ord.assign(cols * rows, -1);
    int64_t unfilled = 0;
    bool any_skip = false;
    for (size_t i = 0; i < cols; i++)
        for (size_t j = 0; j < rows; j++) {
            PixelRef p((short)i, (short)j);
            if (map.getPoint(p).filled()) {
                o.ord[i * rows + j] = (int32_t)o.cells.size();
                o.cells.push_back(p);
                // skipped as a source, counted but not expanded under a radius (vgavisualglobal.cpp:75, 108-110)
                const bool skip = map.getPoint(p).contextfilled() && !p.iseven();
                o.skip.push_back(skip ? 1 : 0);
                any_skip = any_skip || skip;
            } else {
                o.ord[i * rows + j] = -(int32_t)(1 + unfilled++);
            }
        }
    o.n = (int64_t)o.cells.size();
    o.ghosts = unfilled;
    if (!any_skip) o.skip.clear();
    return o;
}

inline void set_refs(vga_graph *g, PointMap &map, const Ordinals &o) {
    std::vector<int32_t> refs;
    refs.reserve((size_t)(o.n + o.ghosts));
    for (const PixelRef &p : o.cells) refs.push_back((int)p);
    const size_t cols = map.getCols(), rows = map.getRows();
    for (size_t i = 0; i < cols; i++)
        for (size_t j = 0; j < rows; j++)
            if (o.ord[i * rows + j] < 0) refs.push_back((int)PixelRef((short)i, (short)j));
    vga_graph_set_cell_refs(g, refs.data(), (int64_t)refs.size());
}

// adjacency of a made graph, flattened from the Nodes (Node::first/next iteration order), as ordinal CSR
inline void rows_from_nodes(PointMap &map, const Ordinals &o, std::vector<uint64_t> &rowptr, std::vector<uint32_t> &col) {
    rowptr.assign((size_t)o.n + 1, 0);
    col.clear();
    const size_t rows = map.getRows();
    for (int64_t v = 0; v < o.n; v++) {
        Point &pt = map.getPoint(o.cells[(size_t)v]);
        if (pt.hasNode()) {
            Node &node = pt.getNode();
            node.first();
            while (!node.is_tail()) {
                PixelRef w = node.cursor();
                int32_t id = o.ord[(size_t)w.x * rows + (size_t)w.y];
                col.push_back(id >= 0 ? (uint32_t)id : (uint32_t)(o.n + (-(id + 1))));
                node.next();
            }
        }
        rowptr[(size_t)v + 1] = col.size();
    }
}

// The adjacency the BFS analyses run on: merged pairs contracted (merge_contract.h).  primary is empty when nothing is
// merged; c is filled only then.
inline vga_graph *analysis_graph(PointMap &map, const Ordinals &o, dmx::Contracted &c, std::vector<int32_t> &primary) {
    std::vector<uint64_t> rowptr;
    std::vector<uint32_t> col;
    rows_from_nodes(map, o, rowptr, col);
    const size_t rows = map.getRows();
    std::vector<int32_t> partner((size_t)o.n, -1);
    bool any = false;
    for (int64_t v = 0; v < o.n; v++) {
        Point &pt = map.getPoint(o.cells[(size_t)v]);
        PixelRef m = pt.getMergePixel();
        if (m.empty()) continue;
        if (!map.includes(m) || !map.getPoint(m).filled() || map.getPoint(m).getMergePixel() != o.cells[(size_t)v])
            throw depthmapX::RuntimeException("GPU path: merge links must pair filled cells symmetrically");
        partner[(size_t)v] = o.ord[(size_t)m.x * rows + (size_t)m.y];
        any = true;
    }
    primary.clear();
    vga_graph *g = nullptr;
    int rc;
    if (any) {
        dmx::contract_rows(o.n, rowptr.data(), col.data(), partner.data(), c);
        c.ghosts = o.ghosts;
        primary = c.primary;
        rc = vga_graph_from_csr(gpu(), o.n, o.ghosts, c.rowptr.data(), c.col.data(), nullptr, &g);
    } else {
        rc = vga_graph_from_csr(gpu(), o.n, o.ghosts, rowptr.data(), col.data(), nullptr, &g);
    }
    if (rc != VGA_OK) throw depthmapX::RuntimeException(std::string("GPU path: ") + vga_last_error());
    set_refs(g, map, o);
    if (!o.skip.empty()) {
        if (any) throw depthmapX::RuntimeException("GPU path: context-filled cells together with merge links are not supported");
        if (vga_graph_set_noexpand(g, o.skip.data()) != VGA_OK)
            throw depthmapX::RuntimeException(std::string("GPU path: ") + vga_last_error());
    }
    return g;
}

// level of every source to a vertex set = vga_step_depth from the set over the transposed adjacency
struct GpuLevelTo : dmx::LevelTo {
    vga_graph *graph = nullptr;
    int64_t n = 0;
    ~GpuLevelTo() override {
        if (graph) vga_graph_free(graph);
    }
    void prepare(int64_t cells, const std::vector<uint64_t> &t_rowptr, const std::vector<uint32_t> &t_col) override {
        n = cells;
        if (vga_graph_from_csr(gpu(), n, 0, t_rowptr.data(), t_col.data(), nullptr, &graph) != VGA_OK)
            throw depthmapX::RuntimeException(std::string("GPU path: ") + vga_last_error());
    }
    void run(const std::vector<int64_t> &seeds, std::vector<int32_t> &level) override {
        level.assign((size_t)n, -1);
        if (vga_step_depth(gpu(), graph, seeds.data(), (int64_t)seeds.size(), level.data()) != VGA_OK)
            throw depthmapX::RuntimeException(std::string("GPU path: ") + vga_last_error());
    }
};

// adjacency of a made graph as it is (local measures ignore merge links)
inline vga_graph *graph_from_nodes(PointMap &map, const Ordinals &o) {
    std::vector<uint64_t> rowptr((size_t)o.n + 1, 0);
    std::vector<uint32_t> col;
    const size_t rows = map.getRows();
    for (int64_t v = 0; v < o.n; v++) {
        Point &pt = map.getPoint(o.cells[(size_t)v]);
        if (pt.hasNode()) {
            Node &node = pt.getNode();
            node.first();
            while (!node.is_tail()) {
                PixelRef w = node.cursor();
                int32_t id = o.ord[(size_t)w.x * rows + (size_t)w.y];
                col.push_back(id >= 0 ? (uint32_t)id : (uint32_t)(o.n + (-(id + 1))));
                node.next();
            }
        }
        rowptr[(size_t)v + 1] = col.size();
    }
    vga_graph *g = nullptr;
    if (vga_graph_from_csr(gpu(), o.n, o.ghosts, rowptr.data(), col.data(), nullptr, &g) != VGA_OK)
        throw depthmapX::RuntimeException(std::string("GPU path: ") + vga_last_error());
    set_refs(g, map, o);
    return g;
}

}  // namespace vga_shim
