// Shims: VGAVisualGlobal::run and VGAVisualLocal::run on the GPU (replace the translation units
// salalib/vgamodules/vgavisualglobal.cpp and vgavisuallocal.cpp of the reference).  The adjacency is
// flattened from the map's Nodes, the BFS / neighbourhood counting runs in libvga_b200.so, and the
// attribute rows are written with the same column set, the same conditional setValue calls and the same
// row order as the reference, so that the serialised column statistics are identical.
#include "shim_common.h"

#include "genlib/stringutils.h"
#include "salalib/vgamodules/vgavisualglobal.h"
#include "salalib/vgamodules/vgavisuallocal.h"

bool VGAVisualGlobal::run(Communicator *comm, PointMap &map, bool simple_version) {
    using namespace vga_shim;
    CommState cs{comm, 0};
    if (comm) {
        qtimer(cs.atime, 0);
        comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    }
    AttributeTable &attributes = map.getAttributeTable();
    std::string radius_text;
    if (m_radius != -1) radius_text = std::string(" R") + dXstring::formatString(int(m_radius), "%d");
    int entropy_col = -1, rel_entropy_col = -1, integ_dv_col = -1, integ_pv_col = -1, integ_tk_col = -1, depth_col = -1,
        count_col = -1;
    if (!simple_version) entropy_col = attributes.insertOrResetColumn("Visual Entropy" + radius_text);
    integ_dv_col = attributes.insertOrResetColumn("Visual Integration [HH]" + radius_text);
    if (!simple_version) {
        integ_pv_col = attributes.insertOrResetColumn("Visual Integration [P-value]" + radius_text);
        integ_tk_col = attributes.insertOrResetColumn("Visual Integration [Tekl]" + radius_text);
        depth_col = attributes.insertOrResetColumn("Visual Mean Depth" + radius_text);
        count_col = attributes.insertOrResetColumn("Visual Node Count" + radius_text);
        rel_entropy_col = attributes.insertOrResetColumn("Visual Relativised Entropy" + radius_text);
    }

    if (m_gates_only) {  // the reference skips every cell (vgavisualglobal.cpp:75-78): columns only
        map.setDisplayedAttribute(integ_dv_col);
        return true;
    }
    Ordinals o = make_ordinals(map);
    dmx::Contracted contracted;
    std::vector<int32_t> primary;  // merge links: results of a pair's secondary cell are its primary's
    vga_graph *gr = analysis_graph(map, o, contracted, primary);
    const int64_t N = o.n;
    std::vector<int32_t> nodes((size_t)N);
    std::vector<int64_t> depth((size_t)N);
    int32_t maxl = 64, used = 0;
    std::vector<int32_t> dist;
    vga_ctx_set_callbacks(gpu(), progress_cb, cancel_cb, &cs);
    int rc;
    for (;;) {
        dist.assign((size_t)N * maxl, 0);
        rc = vga_global(gpu(), gr, (int)m_radius, 0, N, nodes.data(), depth.data(), dist.data(), maxl, &used);
        if (rc == VGA_ERR_CAPACITY && used > maxl) {
            maxl = used;
            continue;
        }
        break;
    }
    vga_ctx_set_callbacks(gpu(), nullptr, nullptr, nullptr);
    vga_graph_free(gr);
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw depthmapX::RuntimeException(std::string("GPU VGA global: ") + vga_last_error());
    if (!primary.empty()) {
        if ((int)m_radius != -1) {
            GpuLevelTo level_to;
            dmx::radius_correction(contracted, (int)m_radius, level_to, nodes.data(), depth.data(), dist.data(), maxl);
        }
        dmx::copy_from_primary(primary, nodes.data());
        dmx::copy_from_primary(primary, depth.data());
        dmx::copy_from_primary(primary, dist.data(), (size_t)maxl);
    }

    std::vector<float> nc((size_t)N), md((size_t)N), hh((size_t)N), pv((size_t)N), tk((size_t)N), en((size_t)N), re((size_t)N);
    vga_global_attributes(N, nodes.data(), depth.data(), dist.data(), maxl, nc.data(), md.data(), hh.data(), pv.data(),
                          tk.data(), en.data(), re.data());
    // same calls, same order per row as the reference's formula stage (which value is written when)
    for (int64_t v = 0; v < N; v++) {
        if (!o.skip.empty() && o.skip[(size_t)v]) continue;  // not a source in the reference
        AttributeRow &row = attributes.getRow(AttributeKey(o.cells[(size_t)v]));
        const int tn = nodes[(size_t)v];
        if (!simple_version) row.setValue(count_col, nc[(size_t)v]);
        if (tn > 1) {
            if (!simple_version) row.setValue(depth_col, md[(size_t)v]);
            // integration columns are always written when there is more than one node (value or -1)
            row.setValue(integ_dv_col, hh[(size_t)v]);
            if (!simple_version) {
                row.setValue(integ_pv_col, pv[(size_t)v]);
                row.setValue(integ_tk_col, tk[(size_t)v]);
                row.setValue(entropy_col, en[(size_t)v]);
                row.setValue(rel_entropy_col, re[(size_t)v]);
            }
        } else if (!simple_version) {
            row.setValue(depth_col, (float)-1);
            row.setValue(entropy_col, (float)-1);
            row.setValue(rel_entropy_col, (float)-1);
        }
    }
    map.setDisplayedAttribute(integ_dv_col);
    return true;
}

bool VGAVisualLocal::run(Communicator *comm, PointMap &map, bool simple_version) {
    using namespace vga_shim;
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    int cluster_col = -1, control_col = -1, controllability_col = -1;
    if (!simple_version) {
        cluster_col = map.getAttributeTable().insertOrResetColumn("Visual Clustering Coefficient");
        control_col = map.getAttributeTable().insertOrResetColumn("Visual Control");
        controllability_col = map.getAttributeTable().insertOrResetColumn("Visual Controllability");
    }
    if (m_gates_only) {  // vgavisuallocal.cpp:43-46
        if (!simple_version) map.setDisplayedAttribute(cluster_col);
        return true;
    }
    Ordinals o = make_ordinals(map);
    vga_graph *gr = graph_from_nodes(map, o);
    const int64_t N = o.n;
    std::vector<int64_t> cluster((size_t)N);
    std::vector<int32_t> k((size_t)N), total((size_t)N);
    std::vector<float> control((size_t)N);
    int rc = vga_local(gpu(), gr, 0, N, cluster.data(), k.data(), total.data(), control.data());
    vga_graph_free(gr);
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw depthmapX::RuntimeException(std::string("GPU VGA local: ") + vga_last_error());
    if (!simple_version) {
        std::vector<float> a((size_t)N), b((size_t)N), c((size_t)N);
        vga_local_attributes(N, cluster.data(), k.data(), total.data(), control.data(), a.data(), b.data(), c.data());
        for (int64_t v = 0; v < N; v++) {
            if (!o.skip.empty() && o.skip[(size_t)v]) continue;  // vgavisuallocal.cpp:43
            AttributeRow &row = map.getAttributeTable().getRow(AttributeKey(o.cells[(size_t)v]));
            row.setValue(cluster_col, a[(size_t)v]);
            row.setValue(control_col, b[(size_t)v]);
            row.setValue(controllability_col, c[(size_t)v]);
        }
        map.setDisplayedAttribute(cluster_col);
    }
    return true;
}
