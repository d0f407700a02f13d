// Shim: PointMap::sparkGraph2 on the GPU.  Replaces the body at salalib/pointdata.cpp:1246-1341 (the
// reference object keeps every other member; its sparkGraph2 symbol is weakened at link time, see
// integration/Makefile).  Host side effects are reproduced in the reference's order; the per-source
// work (sparkPixel2 / sieve2 / sparkSieve2) runs in libvga_b200.so; the 32-bin Node layout is produced
// by the reference's own Node::make from the rows the library returns.
#include "shim_common.h"

#include "salalib/ngraph.h"

bool PointMap::sparkGraph2(Communicator *comm, bool boundarygraph, double maxdist) {
    using namespace vga_shim;
    if (!m_blockedlines) blockLines();
    if (boundarygraph) {
        for (size_t i = 0; i < m_cols; i++)
            for (size_t j = 0; j < m_rows; j++) {
                PixelRef curs((short)i, (short)j);
                if (getPoint(curs).filled() && !getPoint(curs).edge()) {
                    m_points(j, i).m_state &= ~Point::FILLED;
                    m_filled_point_count--;
                }
            }
    }
    int connectivity_col = m_attributes->insertOrResetLockedColumn("Connectivity");
    int m1_col = m_attributes->insertOrResetColumn("Point First Moment");
    int m2_col = m_attributes->insertOrResetColumn("Point Second Moment");
    int count = tagState(true);
    CommState cs{comm, 0};
    if (comm) {
        qtimer(cs.atime, 0);
        comm->CommPostMessage(Communicator::NUM_RECORDS, count);
    }

    // flat image of Point::m_state / Point::m_lines (x-major)
    const size_t cells = m_cols * m_rows;
    std::vector<uint16_t> state(cells);
    std::vector<uint32_t> off(cells + 1);
    std::vector<double> lines;
    size_t c = 0;
    for (size_t i = 0; i < m_cols; i++)
        for (size_t j = 0; j < m_rows; j++, c++) {
            Point &p = getPoint(PixelRef((short)i, (short)j));
            state[c] = (uint16_t)p.m_state;
            off[c] = (uint32_t)(lines.size() / 5);
            for (const Line &l : p.m_lines) {
                lines.push_back(l.bottom_left.x);
                lines.push_back(l.bottom_left.y);
                lines.push_back(l.top_right.x);
                lines.push_back(l.top_right.y);
                lines.push_back(l.sign() > 0 ? 1.0 : 0.0);
            }
        }
    off[cells] = (uint32_t)(lines.size() / 5);
    vga_grid grid;
    grid.cols = (int32_t)m_cols;
    grid.rows = (int32_t)m_rows;
    grid.spacing = m_spacing;
    grid.bl_x = m_bottom_left.x;
    grid.bl_y = m_bottom_left.y;
    grid.maxdist = maxdist;
    grid.state = state.data();
    grid.line_off = off.data();
    grid.lines = lines.empty() ? nullptr : lines.data();

    vga_ctx_set_callbacks(gpu(), progress_cb, cancel_cb, &cs);
    vga_graph *gr = nullptr;
    int rc = vga_graph_build(gpu(), &grid, 0, -1, &gr);
    vga_ctx_set_callbacks(gpu(), nullptr, nullptr, nullptr);
    if (rc == VGA_ERR_CANCELLED) {
        tagState(false);
        m_attributes->clear();
        m_displayed_attribute = -2;
        throw Communicator::CancelledException();
    }
    if (rc != VGA_OK) throw depthmapX::RuntimeException(std::string("GPU makegraph: ") + vga_last_error());

    const int64_t N = vga_graph_num_cells(gr), E = vga_graph_num_edges(gr), G = vga_graph_num_ghosts(gr);
    std::vector<uint64_t> rowptr((size_t)N + 1);
    std::vector<uint32_t> col((size_t)E + 1);
    std::vector<uint8_t> bin((size_t)E + 1), acc((size_t)E + 1);
    vga_graph_csr(gr, rowptr.data(), col.data(), bin.data(), acc.data());
    std::vector<int32_t> ref((size_t)(N + G));
    vga_graph_cell_refs(gr, ref.data());
    std::vector<int32_t> conn((size_t)N);
    std::vector<double> sd((size_t)N), sd2((size_t)N);
    std::vector<float> far((size_t)N * 32);
    vga_graph_node_stats(gr, conn.data(), sd.data(), sd2.data(), far.data(), nullptr, nullptr);
    vga_graph_free(gr);

    PixelRefVector bins_b[32];
    for (int64_t v = 0; v < N; v++) {  // x-major: the order rows are added in the reference
        PixelRef curs(ref[(size_t)v]);
        Point &pt = getPoint(curs);
        pt.m_node = std::unique_ptr<Node>(new Node());
        m_attributes->addRow(AttributeKey(curs));
        for (uint64_t e = rowptr[(size_t)v]; e < rowptr[(size_t)v + 1]; e++)
            if (acc[e]) bins_b[bin[e]].push_back(PixelRef(ref[col[e]]));
        // rows are sorted by (x, y): a diagonal bin's first / last entries are its extremes in x, which is all
        // Bin::make looks at for diagonal bins; the other bins are re-sorted by Bin::make itself
        pt.m_node->make(curs, bins_b, &far[(size_t)v * 32], 0x00FF);  // clears bins_b
        AttributeRow &row = m_attributes->getRow(AttributeKey(curs));
        row.setValue(connectivity_col, float(conn[(size_t)v]));
        row.setValue(m1_col, float(sd[(size_t)v]));
        row.setValue(m2_col, float(sd2[(size_t)v]));
        pt.m_processflag = 0;
    }
    tagState(false);
    unblockLines(false);
    addGridConnections();
    m_processed = true;
    if (boundarygraph) m_boundarygraph = true;
    m_displayed_attribute = -2;
    setDisplayedAttribute(connectivity_col);
    return true;
}
