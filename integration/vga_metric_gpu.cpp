// Shim: VGAMetric::run and VGAAngular::run on the GPU (replaces the translation units
// salalib/vgamodules/vgametric.cpp and vgaangular.cpp of the reference; `-m VGA -vm metric -vr <r>` and `-vm angular`,
// SURVEY §8 row f4).  Column names, their creation order (insertOrResetColumn for metric, getOrInsertColumn for angular),
// the per-row setValue calls and the displayed attribute are the reference's (vgametric.cpp:32-56, 116-120, 131-133;
// vgaangular.cpp:29-53, 109-114, 128-130); the per-source searches run in vga_metric / vga_angular.
#include "shim_common.h"

#include "genlib/stringutils.h"
#include "salalib/vgamodules/vgaangular.h"
#include "salalib/vgamodules/vgametric.h"

namespace {

// Point::blocked || PointMap::blockedAdjacent for every filled cell: the cells Node::extractMetric / extractAngular expand
std::vector<uint8_t> expand_flags(PointMap &map, const vga_shim::Ordinals &o) {
    std::vector<uint8_t> f((size_t)o.n, 0);
    for (int64_t v = 0; v < o.n; v++) {
        const PixelRef p = o.cells[(size_t)v];
        f[(size_t)v] = (map.getPoint(p).blocked() || map.blockedAdjacent(p)) ? 1 : 0;
    }
    return f;
}

// Point::m_merge as ordinals (-1 = not merged); empty when the map has no merge links
std::vector<int32_t> merge_partners(PointMap &map, const vga_shim::Ordinals &o) {
    std::vector<int32_t> partner((size_t)o.n, -1);
    const size_t rows = map.getRows();
    bool any = false;
    for (int64_t v = 0; v < o.n; v++) {
        PixelRef m = map.getPoint(o.cells[(size_t)v]).getMergePixel();
        if (m.empty()) continue;
        if (!map.includes(m) || !map.getPoint(m).filled() || map.getPoint(m).getMergePixel() != o.cells[(size_t)v])
            throw depthmapX::RuntimeException("GPU path: merge links must pair filled cells symmetrically");
        partner[(size_t)v] = o.ord[(size_t)m.x * rows + (size_t)m.y];
        any = true;
    }
    if (!any) partner.clear();
    return partner;
}

}  // namespace

bool VGAMetric::run(Communicator *comm, PointMap &map, bool) {
    using namespace vga_shim;
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    std::string radius_text;
    if (m_radius != -1.0) {
        if (m_radius > 100.0) {
            radius_text = std::string(" R") + dXstring::formatString(m_radius, "%.f");
        } else if (map.getRegion().width() < 1.0) {
            radius_text = std::string(" R") + dXstring::formatString(m_radius, "%.4f");
        } else {
            radius_text = std::string(" R") + dXstring::formatString(m_radius, "%.2f");
        }
    }
    AttributeTable &attributes = map.getAttributeTable();
    int mspa_col = attributes.insertOrResetColumn("Metric Mean Shortest-Path Angle" + radius_text);
    int mspl_col = attributes.insertOrResetColumn("Metric Mean Shortest-Path Distance" + radius_text);
    int dist_col = attributes.insertOrResetColumn("Metric Mean Straight-Line Distance" + radius_text);
    int count_col = attributes.insertOrResetColumn("Metric Node Count" + radius_text);
    if (!m_gates_only) {
        Ordinals o = make_ordinals(map);
        const std::vector<int32_t> partner = merge_partners(map, o);
        vga_graph *gr = graph_from_nodes(map, o);
        const std::vector<uint8_t> flags = expand_flags(map, o);
        std::vector<float> angle((size_t)o.n), path((size_t)o.n), line((size_t)o.n), count((size_t)o.n);
        CommState cs{comm, 0};
        vga_ctx_set_callbacks(gpu(), progress_cb, cancel_cb, &cs);
        int rc = vga_metric(gpu(), gr, flags.data(), partner.empty() ? nullptr : partner.data(), map.getSpacing(), m_radius, nullptr, o.n, angle.data(), path.data(),
                            line.data(), count.data(), nullptr);
        vga_ctx_set_callbacks(gpu(), nullptr, nullptr, nullptr);
        vga_graph_free(gr);
        if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
        if (rc != VGA_OK) throw depthmapX::RuntimeException(std::string("GPU metric analysis: ") + vga_last_error());
        for (int64_t v = 0; v < o.n; v++) {
            AttributeRow &row = attributes.getRow(AttributeKey(o.cells[(size_t)v]));
            row.setValue(mspa_col, angle[(size_t)v]);
            row.setValue(mspl_col, path[(size_t)v]);
            row.setValue(dist_col, line[(size_t)v]);
            row.setValue(count_col, count[(size_t)v]);
        }
    }
    map.overrideDisplayedAttribute(-2);
    map.setDisplayedAttribute(mspl_col);
    return true;
}

bool VGAAngular::run(Communicator *comm, PointMap &map, bool) {
    using namespace vga_shim;
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    std::string radius_text;
    if (m_radius != -1.0) {
        if (map.getRegion().width() > 100.0) {
            radius_text = std::string(" R") + dXstring::formatString(m_radius, "%.f");
        } else if (map.getRegion().width() < 1.0) {
            radius_text = std::string(" R") + dXstring::formatString(m_radius, "%.4f");
        } else {
            radius_text = std::string(" R") + dXstring::formatString(m_radius, "%.2f");
        }
    }
    AttributeTable &attributes = map.getAttributeTable();
    int mean_depth_col = attributes.getOrInsertColumn("Angular Mean Depth" + radius_text);
    int total_depth_col = attributes.getOrInsertColumn("Angular Total Depth" + radius_text);
    int count_col = attributes.getOrInsertColumn("Angular Node Count" + radius_text);
    if (!m_gates_only) {
        Ordinals o = make_ordinals(map);
        const std::vector<int32_t> partner = merge_partners(map, o);
        vga_graph *gr = graph_from_nodes(map, o);
        const std::vector<uint8_t> flags = expand_flags(map, o);
        std::vector<float> mean((size_t)o.n), total((size_t)o.n), count((size_t)o.n);
        CommState cs{comm, 0};
        vga_ctx_set_callbacks(gpu(), progress_cb, cancel_cb, &cs);
        int rc = vga_angular(gpu(), gr, flags.data(), partner.empty() ? nullptr : partner.data(), m_radius, nullptr, o.n, mean.data(), total.data(), count.data(), nullptr);
        vga_ctx_set_callbacks(gpu(), nullptr, nullptr, nullptr);
        vga_graph_free(gr);
        if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
        if (rc != VGA_OK) throw depthmapX::RuntimeException(std::string("GPU angular analysis: ") + vga_last_error());
        for (int64_t v = 0; v < o.n; v++) {
            AttributeRow &row = attributes.getRow(AttributeKey(o.cells[(size_t)v]));
            if (count[(size_t)v] > 0.0f) row.setValue(mean_depth_col, mean[(size_t)v]);
            row.setValue(total_depth_col, total[(size_t)v]);
            row.setValue(count_col, count[(size_t)v]);
        }
    }
    map.setDisplayedAttribute(-2);
    map.setDisplayedAttribute(mean_depth_col);
    return true;
}
