// dmx::PointMap -- host pre-steps (grid, per-cell wall lists, flood fill) and the GPU-backed
// sparkGraph2.  See pointmap.h for the reference entry points each method stands in for.
#include "pointmap.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>

namespace dmx {

// ------------------------------------------------------------------------------------ context

static vga_ctx *g_ctx = nullptr;

vga_ctx *shared_context() {
    if (g_ctx) return g_ctx;
    int dev = 0;
    if (const char *e = std::getenv("VGA_DEVICE"))
        dev = std::atoi(e);
    else if (const char *r = std::getenv("LOCAL_RANK"))
        dev = std::atoi(r);
    int rc = vga_ctx_create(dev, &g_ctx);
    if (rc != VGA_OK) throw RuntimeException(std::string("GPU context: ") + vga_last_error());
    return g_ctx;
}

void release_shared_context() {
    if (g_ctx) vga_ctx_destroy(g_ctx);
    g_ctx = nullptr;
}

// ------------------------------------------------------------------------------------ attributes

int AttributeTable::insertOrResetColumn(const std::string &name) {
    int idx = getColumnIndex(name);
    if (idx < 0) {
        m_names.push_back(name);
        m_cols.emplace_back(m_keys.size(), -1.0f);
        return (int)m_names.size() - 1;
    }
    std::fill(m_cols[idx].begin(), m_cols[idx].end(), -1.0f);
    return idx;
}

int AttributeTable::getColumnIndex(const std::string &name) const {
    for (size_t i = 0; i < m_names.size(); i++)
        if (m_names[i] == name) return (int)i;
    return -1;
}

void AttributeTable::setRows(const std::vector<int> &keys) {
    m_keys = keys;
    for (auto &c : m_cols) c.assign(m_keys.size(), -1.0f);
}

void AttributeTable::clear() {
    m_names.clear();
    m_cols.clear();
    m_keys.clear();
}

// ------------------------------------------------------------------------------------ PointMap

PointMap::PointMap(const Region &parentRegion, const std::vector<Line> &walls, const std::string &name)
    : m_name(name), m_parent(parentRegion), m_walls(walls) {}

PointMap::~PointMap() {
    if (m_graph) vga_graph_free(m_graph);
}

vga_ctx *PointMap::context() { return shared_context(); }

void PointMap::adoptGraph(vga_graph *g) {
    if (m_graph) vga_graph_free(m_graph);
    m_graph = g;
}

// Grid centres sit on multiples of the spacing ("origin at 0"): pointdata.cpp:122-171
bool PointMap::setGrid(double spacing, const Point2f &offset) {
    m_spacing = spacing;
    double xo = std::fmod(m_parent.bl.x + offset.x, m_spacing);
    double yo = std::fmod(m_parent.bl.y + offset.y, m_spacing);
    if (xo < m_spacing / 2.0) xo += m_spacing;
    if (xo > m_spacing / 2.0) xo -= m_spacing;
    if (yo < m_spacing / 2.0) yo += m_spacing;
    if (yo > m_spacing / 2.0) yo -= m_spacing;
    m_offset = Point2f(-xo, -yo);
    if (!m_points.empty()) m_filled_point_count = 0;
    m_undocounter = 0;
    m_cols = (size_t)((int)std::floor((xo + m_parent.width()) / m_spacing + 0.5) + 1);
    m_rows = (size_t)((int)std::floor((yo + m_parent.height()) / m_spacing + 0.5) + 1);
    m_bottom_left = Point2f(m_parent.bl.x + m_offset.x, m_parent.bl.y + m_offset.y);
    m_region = Region(Point2f(m_bottom_left.x - m_spacing / 2.0, m_bottom_left.y - m_spacing / 2.0),
                      Point2f(m_bottom_left.x + double(m_cols - 1) * m_spacing + m_spacing / 2.0,
                              m_bottom_left.y + double(m_rows - 1) * m_spacing + m_spacing / 2.0));
    m_points.assign(m_cols * m_rows, Point());
    m_filled_point_count = 0;
    m_initialised = true;
    m_blockedlines = false;
    m_processed = false;
    m_boundarygraph = false;
    adoptGraph(nullptr);
    return true;
}

PixelRef PointMap::pixelate(const Point2f &p, bool constrain) const {
    int x = int(std::floor((p.x - m_bottom_left.x + (m_spacing / 2.0)) / m_spacing));
    int y = int(std::floor((p.y - m_bottom_left.y + (m_spacing / 2.0)) / m_spacing));
    if (constrain) {
        x = std::max(0, std::min(x, (int)m_cols - 1));
        y = std::max(0, std::min(y, (int)m_rows - 1));
    }
    return PixelRef(x, y);
}

Region PointMap::regionate(const PixelRef &p, double border) const {
    return Region(Point2f(m_bottom_left.x + m_spacing * (double(p.x) - 0.5 - border),
                          m_bottom_left.y + m_spacing * (double(p.y) - 0.5 - border)),
                  Point2f(m_bottom_left.x + m_spacing * (double(p.x) + 0.5 + border),
                          m_bottom_left.y + m_spacing * (double(p.y) + 0.5 + border)));
}

// Cells a wall touches (touching counts for both cells): spacepix.cpp:144-214.  The segment is
// mapped to grid units, walked along its major axis, and at every step the minor-axis cell at
// both ends of the step is taken (plus the one in between when they differ by 2).
std::vector<PixelRef> PointMap::pixelateLineTouching(Line l, double tol) const {
    std::vector<PixelRef> out;
    auto to_grid = [&](Point2f &p) {
        p.x = m_region.width() ? (p.x - m_region.bl.x) / m_region.width() : 0.0;
        p.y = m_region.height() ? (p.y - m_region.bl.y) / m_region.height() : 0.0;
        p.x *= double(m_cols);
        p.y *= double(m_rows);
    };
    // scale top-right first, then bottom-left (order is irrelevant: independent values)
    to_grid(l.tr);
    to_grid(l.bl);
    const bool along_x = l.width() > l.height();
    const double sgn = l.sign();
    auto inside = [&](int x, int y) {
        // the reference builds a PixelRef (shorts) before the bounds test
        short sx = (short)x, sy = (short)y;
        return sx >= 0 && sx < (short)m_cols && sy >= 0 && sy < (short)m_rows;
    };
    if (along_x) {
        const double grad = sgn * l.height() / l.width();
        const double constant = l.ay() - grad * l.ax();
        const int first = (int)std::floor(l.ax() - tol);
        const int last = (int)std::floor(l.bx() + tol);
        for (int i = first; i <= last; i++) {
            const int j1 = (int)std::floor((first == i ? l.ax() : double(i)) * grad + constant - sgn * tol);
            const int j2 = (int)std::floor((last == i ? l.bx() : double(i + 1)) * grad + constant + sgn * tol);
            if (inside(i, j1)) out.emplace_back(i, j1);
            if (j1 != j2) {
                if (inside(i, j2)) out.emplace_back(i, j2);
                if (std::abs(j2 - j1) == 2) {
                    const int j3 = (j1 + j2) / 2;
                    if (inside(i, j3)) out.emplace_back(i, j3);
                }
            }
        }
    } else {
        const double grad = sgn * l.width() / l.height();
        const double constant = l.ax() - grad * l.ay();
        const int first = (int)std::floor(l.bl.y - tol);
        const int last = (int)std::floor(l.tr.y + tol);
        for (int i = first; i <= last; i++) {
            const int j1 = (int)std::floor((first == i ? l.bl.y : double(i)) * grad + constant - sgn * tol);
            const int j2 = (int)std::floor((last == i ? l.tr.y : double(i + 1)) * grad + constant + sgn * tol);
            if (inside(j1, i)) out.emplace_back(j1, i);
            if (j1 != j2) {
                if (inside(j2, i)) out.emplace_back(j2, i);
                if (std::abs(j2 - j1) == 2) {
                    const int j3 = (j1 + j2) / 2;
                    if (inside(j3, i)) out.emplace_back(j3, i);
                }
            }
        }
    }
    return out;
}

void PointMap::blockLine(const Line &li) {
    for (const PixelRef &p : pixelateLineTouching(li, 1e-10)) {
        Point &pt = getPoint(p);
        pt.lines.push_back(li);
        pt.state |= Point::BLOCKED;
    }
}

void PointMap::unblockLines(bool clearblockedflag) {
    for (Point &pt : m_points) {
        pt.lines.clear();
        pt.lines.shrink_to_fit();
        if (clearblockedflag) pt.state &= ~Point::BLOCKED;
    }
}

// pointdata.cpp:296-343: push every wall into the cells it touches, then clip each copy to its
// cell (border 1e-10 cells) and drop copies that fall outside
bool PointMap::blockLines() {
    if (!m_initialised || m_points.empty()) return false;
    if (m_blockedlines) return true;
    unblockLines();
    for (const Line &w : m_walls) blockLine(Line(w.start(), w.end()));
    for (size_t i = 0; i < m_cols; i++)
        for (size_t j = 0; j < m_rows; j++) {
            PixelRef curs((int)i, (int)j);
            Point &pt = getPoint(curs);
            if (pt.lines.empty()) continue;
            const Region viewport = regionate(curs, 1e-10);
            std::vector<Line> kept;
            kept.reserve(pt.lines.size());
            for (Line l : pt.lines)
                if (l.crop(viewport)) kept.push_back(l);
            pt.lines.swap(kept);
        }
    m_blockedlines = true;
    return true;
}

// one fill step p1 -> p2 (pointdata.cpp:483-514): 1 off grid, 2 already filled, 4 blocked, 8 filled
int PointMap::expand(const PixelRef p1, const PixelRef p2, std::vector<PixelRef> &list, int filltype) {
    if (!includes(p2)) return 1;
    if (getPoint(p2).state & Point::FILLED) return 2;
    const Line sight(depixelate(p1), depixelate(p2));
    const double tol = m_spacing * 1e-10;
    for (const Line &w : getPoint(p1).lines)
        if (blocks(sight, w, tol)) return 4;
    for (const Line &w : getPoint(p2).lines)
        if (blocks(sight, w, tol)) return 4;
    getPoint(p2).set(filltype, m_undocounter);
    m_filled_point_count++;
    list.push_back(p2);
    return 8;
}

// flood fill from a seed over the 8-neighbourhood (pointdata.cpp:402-481)
bool PointMap::makePoints(const Point2f &seed, int fill_type, Communicator *comm) {
    if (!m_initialised || m_points.empty()) return false;
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, (int)(m_rows * m_cols));
    const PixelRef seedref = pixelate(seed, false);
    if (!includes(seedref) || getPoint(seedref).filled()) return false;
    for (const Line &w : getPoint(seedref).lines)
        if (lines_cross_no_touch(w, Line(seed, depixelate(seedref)))) return false;
    if (!m_blockedlines) blockLines();
    m_undocounter++;
    if (fill_type != 0) throw RuntimeException("makePoints: only FULLFILL (fill_type 0) is supported by the GPU path");
    const int filltype = Point::FILLED;
    getPoint(seedref).set(filltype, m_undocounter);
    m_filled_point_count++;
    std::vector<PixelRef> a, b;
    a.push_back(seedref);
    int added = 0;
    auto last = std::chrono::steady_clock::now();
    while (!a.empty()) {
        const PixelRef cur = a.back();
        int result = 0;
        result |= expand(cur, PixelRef(cur.x, cur.y + 1), b, filltype);
        result |= expand(cur, PixelRef(cur.x, cur.y - 1), b, filltype);
        result |= expand(cur, PixelRef(cur.x - 1, cur.y), b, filltype);
        result |= expand(cur, PixelRef(cur.x + 1, cur.y), b, filltype);
        result |= expand(cur, PixelRef(cur.x - 1, cur.y + 1), b, filltype);
        result |= expand(cur, PixelRef(cur.x + 1, cur.y + 1), b, filltype);
        result |= expand(cur, PixelRef(cur.x - 1, cur.y - 1), b, filltype);
        result |= expand(cur, PixelRef(cur.x + 1, cur.y - 1), b, filltype);
        if ((result & 4) || getPoint(cur).blocked()) getPoint(cur).state |= Point::EDGE;
        a.pop_back();
        if (a.empty()) a.swap(b);
        added++;
        if (comm) {
            auto now = std::chrono::steady_clock::now();
            if (now - last > std::chrono::milliseconds(500)) {
                last = now;
                if (comm->IsCancelled()) throw Communicator::CancelledException();
                comm->CommPostMessage(Communicator::CURRENT_RECORD, added);
            }
        }
    }
    return true;
}

void PointMap::flatten(Flat &out) const {
    const size_t cells = m_cols * m_rows;
    out.state.resize(cells);
    out.line_off.resize(cells + 1);
    out.lines.clear();
    uint32_t n = 0;
    for (size_t c = 0; c < cells; c++) {
        const Point &pt = m_points[c];
        out.state[c] = (uint16_t)pt.state;
        out.line_off[c] = n;
        for (const Line &l : pt.lines) {
            out.lines.push_back(l.bl.x);
            out.lines.push_back(l.bl.y);
            out.lines.push_back(l.tr.x);
            out.lines.push_back(l.tr.y);
            out.lines.push_back(l.parity ? 1.0 : 0.0);
            n++;
        }
    }
    out.line_off[cells] = n;
}

bool PointMap::unmake() {
    for (Point &pt : m_points)
        if (pt.filled()) {
            pt.grid_connections = 0;
            pt.lines.clear();
            pt.state &= ~Point::BLOCKED;
        }
    m_blockedlines = false;
    m_attributes.clear();
    m_processed = false;
    m_boundarygraph = false;
    m_displayed_attribute = -2;
    adoptGraph(nullptr);
    return true;
}

namespace {
struct CbState {
    Communicator *comm;
    std::chrono::steady_clock::time_point last;
    bool cancelled;
};
void progress_cb(void *u, int64_t done, int64_t) {
    CbState *s = (CbState *)u;
    if (s->comm) s->comm->CommPostMessage(Communicator::CURRENT_RECORD, (int)done);
}
int cancel_cb(void *u) {
    CbState *s = (CbState *)u;
    if (!s->comm) return 0;
    auto now = std::chrono::steady_clock::now();
    if (now - s->last < std::chrono::milliseconds(500)) return 0;
    s->last = now;
    if (s->comm->IsCancelled()) s->cancelled = true;
    return s->cancelled ? 1 : 0;
}
}  // namespace

// pointdata.cpp:1246-1341.  Host: boundary un-fill, columns, rows, flags.  GPU: everything per source.
bool PointMap::sparkGraph2(Communicator *comm, bool boundarygraph, double maxdist) {
    if (!m_blockedlines) blockLines();
    if (boundarygraph) {
        for (Point &pt : m_points)
            if (pt.filled() && !pt.edge()) {
                pt.state &= ~Point::FILLED;
                m_filled_point_count--;
            }
    }
    const int connectivity_col = m_attributes.insertOrResetColumn("Connectivity");
    const int m1_col = m_attributes.insertOrResetColumn("Point First Moment");
    const int m2_col = m_attributes.insertOrResetColumn("Point Second Moment");

    std::vector<int> keys;
    for (size_t i = 0; i < m_cols; i++)
        for (size_t j = 0; j < m_rows; j++)
            if (m_points[i * m_rows + j].filled()) keys.push_back(int(PixelRef((int)i, (int)j)));
    const int64_t n = (int64_t)keys.size();
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, (int)n);

    Flat flat;
    flatten(flat);
    vga_grid grid;
    grid.cols = (int32_t)m_cols;
    grid.rows = (int32_t)m_rows;
    grid.spacing = m_spacing;
    grid.bl_x = m_bottom_left.x;
    grid.bl_y = m_bottom_left.y;
    grid.maxdist = maxdist;
    grid.state = flat.state.data();
    grid.line_off = flat.line_off.data();
    grid.lines = flat.lines.empty() ? nullptr : flat.lines.data();

    vga_ctx *ctx = context();
    CbState cb{comm, std::chrono::steady_clock::now(), false};
    vga_ctx_set_callbacks(ctx, progress_cb, cancel_cb, &cb);
    vga_graph *g = nullptr;
    int rc = vga_graph_build(ctx, &grid, 0, -1, &g);
    vga_ctx_set_callbacks(ctx, nullptr, nullptr, nullptr);
    if (rc == VGA_ERR_CANCELLED) {
        // same clean-up as the reference's cancel path (pointdata.cpp:1303-1312)
        m_attributes.clear();
        m_displayed_attribute = -2;
        throw Communicator::CancelledException();
    }
    if (rc != VGA_OK) throw RuntimeException(std::string("sparkGraph2: ") + vga_last_error());
    adoptGraph(g);

    m_attributes.setRows(keys);
    std::vector<int32_t> conn((size_t)n);
    std::vector<double> sd((size_t)n), sd2((size_t)n);
    std::vector<uint8_t> gc((size_t)n);
    rc = vga_graph_node_stats(g, conn.data(), sd.data(), sd2.data(), nullptr, nullptr, gc.data());
    if (rc != VGA_OK) throw RuntimeException(std::string("sparkGraph2: ") + vga_last_error());
    for (int64_t v = 0; v < n; v++) {
        m_attributes.setValue((size_t)v, connectivity_col, float(conn[(size_t)v]));
        m_attributes.setValue((size_t)v, m1_col, float(sd[(size_t)v]));
        m_attributes.setValue((size_t)v, m2_col, float(sd2[(size_t)v]));
        getPoint(PixelRef(keys[(size_t)v])).grid_connections = gc[(size_t)v];
    }
    unblockLines(false);
    m_processed = true;
    if (boundarygraph) m_boundarygraph = true;
    m_displayed_attribute = -2;
    setDisplayedAttribute(connectivity_col);
    return true;
}

// ------------------------------------------------------------------------------------ analyses

static void check_supported(const PointMap &map, bool gates_only, const char *who) {
    if (gates_only) throw RuntimeException(std::string(who) + ": gates_only is not supported by the GPU path");
    if (!map.graph()) throw RuntimeException(std::string(who) + ": the map has no visibility graph (run sparkGraph2 first)");
}

// vgavisualglobal.cpp:23-216
bool VGAVisualGlobal::run(Communicator *comm, PointMap &map, bool simple_version) {
    check_supported(map, m_gates_only, "VGAVisualGlobal");
    AttributeTable &attributes = map.getAttributeTable();
    const int64_t n = (int64_t)attributes.getNumRows();
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    std::string radius_text;
    if (m_radius != -1) radius_text = std::string(" R") + std::to_string(int(m_radius));
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

    vga_ctx *ctx = map.context();
    CbState cb{comm, std::chrono::steady_clock::now(), false};
    vga_ctx_set_callbacks(ctx, progress_cb, cancel_cb, &cb);
    std::vector<int32_t> nodes((size_t)n);
    std::vector<int64_t> depth((size_t)n);
    int32_t maxl = 32, used = 0;
    std::vector<int32_t> dist;
    int rc;
    while (true) {
        dist.assign((size_t)n * maxl, 0);
        rc = vga_global(ctx, map.graph(), (int)m_radius, 0, n, nodes.data(), depth.data(), dist.data(), maxl, &used);
        if (rc == VGA_ERR_CAPACITY && used > maxl) {
            maxl = used;
            continue;
        }
        break;
    }
    vga_ctx_set_callbacks(ctx, nullptr, nullptr, nullptr);
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw RuntimeException(std::string("VGAVisualGlobal: ") + vga_last_error());

    std::vector<float> nc((size_t)n), md((size_t)n), hh((size_t)n), pv((size_t)n), tk((size_t)n), en((size_t)n), re((size_t)n);
    vga_global_attributes(n, nodes.data(), depth.data(), dist.data(), maxl, nc.data(), md.data(), hh.data(), pv.data(),
                          tk.data(), en.data(), re.data());
    attributes.column(integ_dv_col) = hh;
    if (!simple_version) {
        attributes.column(count_col) = nc;
        attributes.column(depth_col) = md;
        attributes.column(integ_pv_col) = pv;
        attributes.column(integ_tk_col) = tk;
        attributes.column(entropy_col) = en;
        attributes.column(rel_entropy_col) = re;
    }
    map.setDisplayedAttribute(integ_dv_col);
    return true;
}

// vgavisuallocal.cpp:23-117
bool VGAVisualLocal::run(Communicator *comm, PointMap &map, bool simple_version) {
    check_supported(map, m_gates_only, "VGAVisualLocal");
    AttributeTable &attributes = map.getAttributeTable();
    const int64_t n = (int64_t)attributes.getNumRows();
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    int cluster_col = -1, control_col = -1, controllability_col = -1;
    if (!simple_version) {
        cluster_col = attributes.insertOrResetColumn("Visual Clustering Coefficient");
        control_col = attributes.insertOrResetColumn("Visual Control");
        controllability_col = attributes.insertOrResetColumn("Visual Controllability");
    }
    vga_ctx *ctx = map.context();
    std::vector<int64_t> cluster((size_t)n);
    std::vector<int32_t> k((size_t)n), total((size_t)n);
    std::vector<float> control((size_t)n);
    int rc = vga_local(ctx, map.graph(), 0, n, cluster.data(), k.data(), total.data(), control.data());
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw RuntimeException(std::string("VGAVisualLocal: ") + vga_last_error());
    if (!simple_version) {
        vga_local_attributes(n, cluster.data(), k.data(), total.data(), control.data(),
                             attributes.column(cluster_col).data(), attributes.column(control_col).data(),
                             attributes.column(controllability_col).data());
        map.setDisplayedAttribute(cluster_col);
    }
    return true;
}

}  // namespace dmx
