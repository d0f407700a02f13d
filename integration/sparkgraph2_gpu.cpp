// Shim: PointMap::sparkGraph2 on the GPU.  Replaces the body at salalib/pointdata.cpp:1246-1341 (the
// reference object keeps every other member; its sparkGraph2 symbol is weakened at link time, see
// integration/Makefile).  Host side effects are reproduced in the reference's order; the per-source
// work (sparkPixel2 / sieve2 / sparkSieve2) runs in libvga_b200.so; the 32-bin Node layout is produced
// directly from the sorted rows the library returns: Bin::make (salalib/ngraph.cpp:234-304) sorts every bin's pixels
// through a std::set (half of the reference's makegraph time) only to find the runs that the sorted
// rows already spell out.  Node / Bin keep those members protected; a maintainer would add a `Node::makeFromRuns`
// member or a friend declaration -- this translation unit opens them for itself instead (standard headers first, so
// that only the reference's headers are affected).  VGA_SHIM_NODE_MAKE=1 selects the reference's own Node::make.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <deque>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>
#define protected public
#define private public
#include "salalib/ngraph.h"
#undef protected
#undef private
#include "shim_common.h"

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

    const bool use_node_make = std::getenv("VGA_SHIM_NODE_MAKE") != nullptr;
    PixelRefVector bins_b[32];
    std::vector<PixelRef> hbuf;
    for (int64_t v = 0; v < N; v++) {  // x-major: the order rows are added in the reference
        PixelRef curs(ref[(size_t)v]);
        Point &pt = getPoint(curs);
        pt.m_node = std::unique_ptr<Node>(new Node());
        m_attributes->addRow(AttributeKey(curs));
        for (uint64_t e = rowptr[(size_t)v]; e < rowptr[(size_t)v + 1]; e++)
            if (acc[e]) bins_b[bin[e]].push_back(PixelRef(ref[col[e]]));
        if (use_node_make) {
            // rows are sorted by (x, y): a diagonal bin's first / last entries are its extremes in x, which is all
            // Bin::make looks at for diagonal bins; the other bins are re-sorted by Bin::make itself
            pt.m_node->make(curs, bins_b, &far[(size_t)v * 32], 0x00FF);  // clears bins_b
        } else {
            Node &node = *pt.m_node;
            node.m_pixel = curs;
            for (int i = 0; i < 32; i++) {
                Bin &b = node.m_bins[i];
                PixelRefVector &px = bins_b[i];  // accepted pixels of the bin in (x, y) order
                b.m_distance = far[(size_t)v * 32 + i];
                b.m_pixel_vecs.clear();
                b.m_node_count = 0;
                if (px.empty()) continue;
                const bool diag = (i == 4 || i == 20 || i == 12 || i == 28);
                const bool vert = (i > 4 && i < 12) || (i > 20 && i < 28);
                b.m_dir = (i == 4 || i == 20) ? PixelRef::POSDIAGONAL : (i == 12 || i == 28) ? PixelRef::NEGDIAGONAL
                          : vert ? PixelRef::VERTICAL : PixelRef::HORIZONTAL;
                if (diag) {
                    // one run from the smallest to the largest x (ngraph.cpp:243-259: first pixel, widened by the last)
                    b.m_pixel_vecs.push_back(PixelVec(px.front(), px.back()));
                } else if (vert) {
                    // PixelRefV order = (x, y) = the order of the row: a run breaks where x changes or y jumps
                    size_t s0 = 0;
                    for (size_t k = 1; k <= px.size(); k++)
                        if (k == px.size() || px[k].x != px[k - 1].x || px[k].y != px[k - 1].y + 1) {
                            b.m_pixel_vecs.push_back(PixelVec(px[s0], px[k - 1]));
                            s0 = k;
                        }
                } else {
                    // PixelRefH order = (y, x)
                    hbuf.assign(px.begin(), px.end());
                    std::sort(hbuf.begin(), hbuf.end(), [](const PixelRef &a, const PixelRef &c) { return a.y != c.y ? a.y < c.y : a.x < c.x; });
                    size_t s0 = 0;
                    for (size_t k = 1; k <= hbuf.size(); k++)
                        if (k == hbuf.size() || hbuf[k].y != hbuf[k - 1].y || hbuf[k].x != hbuf[k - 1].x + 1) {
                            b.m_pixel_vecs.push_back(PixelVec(hbuf[s0], hbuf[k - 1]));
                            s0 = k;
                        }
                }
                b.m_node_count = (unsigned short)px.size();
                px.clear();
            }
        }
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
