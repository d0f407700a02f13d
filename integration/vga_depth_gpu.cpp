// Shim: VGAVisualGlobalDepth::run on the GPU (replaces the translation unit
// salalib/vgamodules/vgavisualglobaldepth.cpp of the reference; `-m STEPDEPTH -sdt visual`).
// The selection set becomes the BFS source set of vga_step_depth; every reached cell gets float(level) in
// "Visual Step Depth".  The reference writes the rows level by level, here they are written in row order: the
// serialised column statistics are a minimum, a maximum and a double sum of small integers, which do not depend on
// the order.
#include "shim_common.h"

#include "salalib/vgamodules/vgavisualglobaldepth.h"

bool VGAVisualGlobalDepth::run(Communicator *, PointMap &map, bool) {
    using namespace vga_shim;
    AttributeTable &attributes = map.getAttributeTable();
    int col = attributes.insertOrResetColumn("Visual Step Depth");
    Ordinals o = make_ordinals(map);
    dmx::Contracted contracted;
    std::vector<int32_t> primary;  // merge links: a pair is one vertex (vgavisualglobaldepth.cpp:54-63)
    vga_graph *gr = analysis_graph(map, o, contracted, primary);
    const size_t rows = map.getRows();
    std::vector<int64_t> sources;
    for (auto &sel : map.getSelSet()) {
        PixelRef p = sel;
        if (!map.includes(p)) continue;
        int32_t id = o.ord[(size_t)p.x * rows + (size_t)p.y];
        if (id >= 0) sources.push_back(primary.empty() ? id : primary[(size_t)id]);
    }
    std::vector<int32_t> depth((size_t)o.n, -1);
    int rc = vga_step_depth(gpu(), gr, sources.data(), (int64_t)sources.size(), depth.data());
    vga_graph_free(gr);
    if (rc != VGA_OK) throw depthmapX::RuntimeException(std::string("GPU step depth: ") + vga_last_error());
    dmx::copy_from_primary(primary, depth.data());
    for (int64_t v = 0; v < o.n; v++)
        if (depth[(size_t)v] >= 0) attributes.getRow(AttributeKey(o.cells[(size_t)v])).setValue(col, float(depth[(size_t)v]));
    map.setDisplayedAttribute(-2);
    map.setDisplayedAttribute(col);
    return true;
}
