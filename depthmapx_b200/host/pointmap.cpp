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

AttributeTable::AttributeTable() {
    // LayerManagerImpl of a new table: one layer "Everything", visible (layermanagerimpl.cpp:22-27), written by
    // LayerManagerImpl::write (:106-149) as available-layers mask, visible mask, count, key 1, name.  The
    // available-layers word is what the reference's `0xffffffff << (32 + 0xfffffffe)` evaluates to in its build
    // (an unsigned 32-bit shift by 30: 0xC0000000); tests/test_graphfile.py checks these bytes against files the
    // reference wrote.
    const int64_t available = 0xC0000000LL, visible = 1, key = 1;
    const int32_t count = 1;
    const char name[] = "Everything";
    const uint32_t len = sizeof(name) - 1;
    m_layers.append((const char *)&available, 8).append((const char *)&visible, 8).append((const char *)&count, 4);
    m_layers.append((const char *)&key, 8).append((const char *)&len, 4).append(name, len);
}

int AttributeTable::insertOrResetColumn(const std::string &name) {
    int idx = getColumnIndex(name);
    if (idx < 0) {
        Column c;
        c.name = name;
        m_columns.push_back(c);
        m_cols.emplace_back(m_keys.size(), -1.0f);
        return (int)m_columns.size() - 1;
    }
    // reset: fresh statistics, unlocked, then setValue(-1) on every row (total -1, min -1, max -1)
    Column &c = m_columns[idx];
    c.min = c.max = c.total = -1.0;
    c.locked = false;
    for (size_t r = 0; r < m_keys.size(); r++) setValue(r, idx, -1.0f);
    return idx;
}

int AttributeTable::getOrInsertColumn(const std::string &name) {
    const int idx = getColumnIndex(name);
    return idx >= 0 ? idx : insertOrResetColumn(name);
}

int AttributeTable::insertOrResetLockedColumn(const std::string &name) {
    int idx = insertOrResetColumn(name);
    m_columns[idx].locked = true;
    return idx;
}

int AttributeTable::getColumnIndex(const std::string &name) const {
    for (size_t i = 0; i < m_columns.size(); i++)
        if (m_columns[i].name == name) return (int)i;
    return -1;
}

// position of the column in name order (the order columns are serialised in); -1 / -2 pass through
int AttributeTable::getColumnSortedIndex(int idx) const {
    if (idx == -1 || idx == -2) return idx;
    if (idx < 0 || idx >= (int)m_columns.size()) return -1;
    int pos = 0;
    for (size_t i = 0; i < m_columns.size(); i++)
        if (m_columns[i].name < m_columns[idx].name) pos++;
    return pos;
}

void AttributeTable::setValue(size_t row, int col, float v) {
    float old = m_cols[col][row];
    m_cols[col][row] = v;
    if (old < 0.0f) old = 0.0f;
    Column &c = m_columns[col];
    if (c.total < 0) {
        c.total = v;
    } else {
        c.total += v;
        c.total -= old;
    }
    if (v > c.max) c.max = v;
    if (c.min < 0 || v < c.min) c.min = v;
}

void AttributeTable::setRows(const std::vector<int> &keys) {
    m_keys = keys;
    m_layer_keys.assign(keys.size(), 1);
    for (auto &c : m_cols) c.assign(m_keys.size(), -1.0f);
}

void AttributeTable::clear() {
    m_columns.clear();
    m_cols.clear();
    m_keys.clear();
    m_layer_keys.clear();
}

// ------------------------------------------------------------------------------------ PointMap

PointMap::PointMap(const Region &parentRegion, const std::vector<Line> &walls, const std::string &name)
    : m_name(name), m_parent(parentRegion), m_walls(walls) {}

PointMap::~PointMap() {
    if (m_graph) vga_graph_free(m_graph);
    if (m_merged_graph) vga_graph_free(m_merged_graph);
}

vga_ctx *PointMap::context() { return shared_context(); }

void PointMap::adoptGraph(vga_graph *g) {
    if (m_graph) vga_graph_free(m_graph);
    m_graph = g;
    dropMergedGraph();
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
    for (size_t j = 0; j < m_cols; j++)
        for (size_t k = 0; k < m_rows; k++) m_points[j * m_rows + k].location = depixelate(PixelRef((int)j, (int)k));
    m_filled_point_count = 0;
    m_nodes.clear();
    m_nodes_valid = false;
    m_selection_set.clear();
    m_has_selection = false;
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
    if (fill_type != 0 && fill_type != 1)
        throw RuntimeException("makePoints: fill types 0 (full) and 1 (semi / context fill) are supported");
    const int filltype = fill_type == 0 ? Point::FILLED : (Point::FILLED | Point::CONTEXTFILLED);
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

bool PointMap::unmake(bool removeLinks) {
    for (Point &pt : m_points)
        if (pt.filled()) {
            if (removeLinks) pt.merge = PixelRef();
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
    m_nodes.clear();
    m_nodes_valid = false;
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

std::vector<int> PointMap::filledKeys() const {
    std::vector<int> keys;
    for (size_t i = 0; i < m_cols; i++)
        for (size_t j = 0; j < m_rows; j++)
            if (m_points[i * m_rows + j].filled()) keys.push_back(int(PixelRef((int)i, (int)j)));
    return keys;
}

std::vector<uint8_t> PointMap::contextSkipFlags() const {
    std::vector<uint8_t> flags;
    bool any = false;
    for (size_t i = 0; i < m_cols; i++)
        for (size_t j = 0; j < m_rows; j++) {
            const Point &pt = m_points[i * m_rows + j];
            if (!pt.filled()) continue;
            const bool skip = pt.contextfilled() && !(i % 2 == 0 && j % 2 == 0);
            flags.push_back(skip ? 1 : 0);
            any = any || skip;
        }
    if (!any) flags.clear();
    return flags;
}

std::vector<uint8_t> PointMap::blockedAdjacentFlags() const {
    std::vector<uint8_t> flags;
    for (size_t i = 0; i < m_cols; i++)
        for (size_t j = 0; j < m_rows; j++) {
            if (!m_points[i * m_rows + j].filled()) continue;
            bool near = false;
            for (int dx = -1; dx <= 1 && !near; dx++)
                for (int dy = -1; dy <= 1 && !near; dy++) {
                    const long x = (long)i + dx, y = (long)j + dy;
                    near = x >= 0 && x < (long)m_cols && y >= 0 && y < (long)m_rows && m_points[(size_t)x * m_rows + (size_t)y].blocked();
                }
            flags.push_back(near ? 1 : 0);
        }
    return flags;
}

std::vector<int32_t> PointMap::mergePartners() const {
    std::vector<int32_t> ord(m_cols * m_rows, -1), partner;
    int32_t n = 0;
    for (size_t c = 0; c < m_cols * m_rows; c++)
        if (m_points[c].filled()) ord[c] = n++;
    bool any = false;
    for (size_t c = 0; c < m_cols * m_rows; c++) {
        const Point &pt = m_points[c];
        if (!pt.filled()) continue;
        int32_t q = -1;
        if (pt.merged()) {
            const PixelRef m = pt.merge;
            if (m.x < 0 || m.y < 0 || (size_t)m.x >= m_cols || (size_t)m.y >= m_rows || ord[(size_t)m.x * m_rows + (size_t)m.y] < 0)
                throw RuntimeException("merge links must pair filled cells");
            q = ord[(size_t)m.x * m_rows + (size_t)m.y];
            any = true;
        }
        partner.push_back(q);
    }
    if (!any) partner.clear();
    return partner;
}

// pointdata.cpp:1250-1264
void PointMap::beginSparkGraph(bool boundarygraph) {
    if (!m_blockedlines) blockLines();
    if (boundarygraph) {
        for (Point &pt : m_points)
            if (pt.filled() && !pt.edge()) {
                pt.state &= ~Point::FILLED;
                m_filled_point_count--;
            }
    }
}

// pointdata.cpp:1268-1341 without the per-source work: columns (Connectivity is a locked column), one row per
// filled cell in x-major order with its three setValue calls, tagState's selection reset, unblockLines(false),
// grid connections, flags, displayed attribute
void PointMap::finishSparkGraph(bool boundarygraph, const int32_t *connectivity, const double *sum_d,
                                const double *sum_d2, const uint8_t *grid_connections) {
    const int connectivity_col = m_attributes.insertOrResetLockedColumn("Connectivity");
    const int m1_col = m_attributes.insertOrResetColumn("Point First Moment");
    const int m2_col = m_attributes.insertOrResetColumn("Point Second Moment");
    m_selection_set.clear();  // tagState (pointdata.cpp:1198-1199); Point::SELECTED bits stay as they are
    m_has_selection = false;
    const std::vector<int> keys = filledKeys();
    m_attributes.setRows(keys);
    for (size_t v = 0; v < keys.size(); v++) {
        m_attributes.setValue(v, connectivity_col, float(connectivity[v]));
        m_attributes.setValue(v, m1_col, float(sum_d[v]));
        m_attributes.setValue(v, m2_col, float(sum_d2[v]));
        getPoint(PixelRef(keys[v])).grid_connections = grid_connections[v];
    }
    unblockLines(false);
    m_processed = true;
    if (boundarygraph) m_boundarygraph = true;
    m_displayed_attribute = -2;
    setDisplayedAttribute(connectivity_col);
    m_nodes.clear();
    m_nodes_valid = false;
}

// pointdata.cpp:1246-1341.  Host: boundary un-fill, columns, rows, flags.  GPU: everything per source.
bool PointMap::sparkGraph2(Communicator *comm, bool boundarygraph, double maxdist) {
    beginSparkGraph(boundarygraph);
    const int64_t n = (int64_t)filledKeys().size();
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

    std::vector<int32_t> conn((size_t)n);
    std::vector<double> sd((size_t)n), sd2((size_t)n);
    std::vector<uint8_t> gc((size_t)n);
    rc = vga_graph_node_stats(g, conn.data(), sd.data(), sd2.data(), nullptr, nullptr, gc.data());
    if (rc != VGA_OK) throw RuntimeException(std::string("sparkGraph2: ") + vga_last_error());
    finishSparkGraph(boundarygraph, conn.data(), sd.data(), sd2.data(), gc.data());
    return true;
}

// ------------------------------------------------------------------------------------ selection

// pointdata.cpp:939-987
bool PointMap::setCurSel(const Region &r, bool add) {
    if (!m_has_selection)
        add = false;
    else if (!add)
        clearSel();
    const PixelRef bl = pixelate(r.bl, true), tr = pixelate(r.tr, true);
    for (int i = bl.x; i <= tr.x; i++)
        for (int j = bl.y; j <= tr.y; j++) {
            Point &pnt = getPoint(PixelRef(i, j));
            if ((pnt.state & Point::FILLED) && (~pnt.state & Point::SELECTED)) {
                pnt.state |= Point::SELECTED;
                const PixelRef ref(i, j);
                auto pos = std::lower_bound(m_selection_set.begin(), m_selection_set.end(), ref,
                                            [](const PixelRef &a, const PixelRef &b) { return int(a) < int(b); });
                if (pos == m_selection_set.end() || int(*pos) != int(ref)) m_selection_set.insert(pos, ref);
                m_has_selection = true;
            }
        }
    return true;
}

// pointdata.cpp:923-937
bool PointMap::clearSel() {
    if (!m_has_selection) return false;
    for (const PixelRef &p : m_selection_set) getPoint(p).state &= ~Point::SELECTED;
    m_selection_set.clear();
    m_has_selection = false;
    return true;
}

// ------------------------------------------------------------------------------------ analyses

static void check_supported(PointMap &map, const char *who) {
    map.ensureGraph();
    if (!map.graph()) throw RuntimeException(std::string(who) + ": the map has no visibility graph (run sparkGraph2 first)");
}

// vgavisualglobal.cpp:33-63 (columns), 131-193 (formulas, which value is written when), 214 (display)
void VGAVisualGlobal::writeAttributes(PointMap &map, double radius, bool simple_version, const int32_t *nodes,
                                      const int64_t *depth, const int32_t *dist, int32_t maxl, const uint8_t *skip) {
    AttributeTable &attributes = map.getAttributeTable();
    const int64_t n = (int64_t)attributes.getNumRows();
    std::string radius_text;
    if (radius != -1) radius_text = std::string(" R") + std::to_string(int(radius));
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
    if (!nodes) {  // gates_only: no cell is analysed, the columns stay at -1
        map.setDisplayedAttribute(integ_dv_col);
        return;
    }
    std::vector<float> nc((size_t)n), md((size_t)n), hh((size_t)n), pv((size_t)n), tk((size_t)n), en((size_t)n), re((size_t)n);
    vga_global_attributes(n, nodes, depth, dist, maxl, nc.data(), md.data(), hh.data(), pv.data(), tk.data(), en.data(),
                          re.data());
    // same setValue calls in the same order as the reference, so the column statistics (double running totals)
    // come out identical
    for (int64_t v = 0; v < n; v++) {
        const size_t r = (size_t)v;
        if (skip && skip[r]) continue;  // not a source in the reference: the row keeps its -1s
        if (!simple_version) attributes.setValue(r, count_col, nc[r]);
        if (nodes[r] > 1) {
            if (!simple_version) attributes.setValue(r, depth_col, md[r]);
            attributes.setValue(r, integ_dv_col, hh[r]);
            if (!simple_version) {
                attributes.setValue(r, integ_pv_col, pv[r]);
                attributes.setValue(r, integ_tk_col, tk[r]);
                attributes.setValue(r, entropy_col, en[r]);
                attributes.setValue(r, rel_entropy_col, re[r]);
            }
        } else if (!simple_version) {
            attributes.setValue(r, depth_col, -1.0f);
            attributes.setValue(r, entropy_col, -1.0f);
            attributes.setValue(r, rel_entropy_col, -1.0f);
        }
    }
    map.setDisplayedAttribute(integ_dv_col);
}

namespace {
// level of every source TO a vertex set = vga_step_depth from the set over the transposed adjacency
struct GpuLevelTo : PointMap::LevelTo {
    vga_ctx *ctx;
    vga_graph *graph = nullptr;
    int64_t n = 0;
    explicit GpuLevelTo(vga_ctx *c) : ctx(c) {}
    ~GpuLevelTo() override {
        if (graph) vga_graph_free(graph);
    }
    void prepare(int64_t cells, const std::vector<uint64_t> &t_rowptr, const std::vector<uint32_t> &t_col) override {
        n = cells;
        if (vga_graph_from_csr(ctx, n, 0, t_rowptr.data(), t_col.data(), nullptr, &graph) != VGA_OK)
            throw RuntimeException(std::string("GPU path: ") + vga_last_error());
    }
    void run(const std::vector<int64_t> &seeds, std::vector<int32_t> &level) override {
        level.assign((size_t)n, -1);
        if (vga_step_depth(ctx, graph, seeds.data(), (int64_t)seeds.size(), level.data()) != VGA_OK)
            throw RuntimeException(std::string("GPU path: ") + vga_last_error());
    }
};
}  // namespace

// vgavisualglobal.cpp:23-216
bool VGAVisualGlobal::run(Communicator *comm, PointMap &map, bool simple_version) {
    check_supported(map, "VGAVisualGlobal");
    const int64_t n = (int64_t)map.getAttributeTable().getNumRows();
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    if (m_gates_only) {
        // the reference skips every cell when gates_only is set (vgavisualglobal.cpp:75-78): columns only
        writeAttributes(map, m_radius, simple_version, nullptr, nullptr, nullptr, 0);
        return true;
    }
    std::vector<int32_t> primary;
    vga_graph *graph = map.analysisGraph(&primary);
    const std::vector<uint8_t> skip = map.contextSkipFlags();
    if (!skip.empty() && !primary.empty())
        throw RuntimeException("VGAVisualGlobal: context-filled cells together with merge links are not supported by the GPU path");
    if (vga_graph_set_noexpand(graph, skip.empty() ? nullptr : skip.data()) != VGA_OK)
        throw RuntimeException(std::string("VGAVisualGlobal: ") + vga_last_error());
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
        rc = vga_global(ctx, graph, (int)m_radius, 0, n, nodes.data(), depth.data(), dist.data(), maxl, &used);
        if (rc == VGA_ERR_CAPACITY && used > maxl) {
            maxl = used;
            continue;
        }
        break;
    }
    vga_ctx_set_callbacks(ctx, nullptr, nullptr, nullptr);
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw RuntimeException(std::string("VGAVisualGlobal: ") + vga_last_error());
    if (!primary.empty()) {
        if ((int)m_radius != -1) {
            GpuLevelTo level_to(ctx);
            map.radiusCorrection((int)m_radius, level_to, nodes.data(), depth.data(), dist.data(), maxl);
        }
        copy_from_primary(primary, nodes.data());
        copy_from_primary(primary, depth.data());
        copy_from_primary(primary, dist.data(), (size_t)maxl);
    }
    writeAttributes(map, m_radius, simple_version, nodes.data(), depth.data(), dist.data(), maxl, skip.empty() ? nullptr : skip.data());
    return true;
}

// vgavisuallocal.cpp:31-35, 84-96, 112
void VGAVisualLocal::writeAttributes(PointMap &map, bool simple_version, const int64_t *cluster, const int32_t *k,
                                     const int32_t *total, const float *control, const uint8_t *skip) {
    if (simple_version) return;
    AttributeTable &attributes = map.getAttributeTable();
    const int64_t n = (int64_t)attributes.getNumRows();
    const int cluster_col = attributes.insertOrResetColumn("Visual Clustering Coefficient");
    const int control_col = attributes.insertOrResetColumn("Visual Control");
    const int controllability_col = attributes.insertOrResetColumn("Visual Controllability");
    if (!cluster) {  // gates_only (vgavisuallocal.cpp:43-46)
        map.setDisplayedAttribute(cluster_col);
        return;
    }
    std::vector<float> a((size_t)n), b((size_t)n), c((size_t)n);
    vga_local_attributes(n, cluster, k, total, control, a.data(), b.data(), c.data());
    for (int64_t v = 0; v < n; v++) {
        if (skip && skip[(size_t)v]) continue;
        attributes.setValue((size_t)v, cluster_col, a[(size_t)v]);
        attributes.setValue((size_t)v, control_col, b[(size_t)v]);
        attributes.setValue((size_t)v, controllability_col, c[(size_t)v]);
    }
    map.setDisplayedAttribute(cluster_col);
}

// vgavisuallocal.cpp:23-117
bool VGAVisualLocal::run(Communicator *comm, PointMap &map, bool simple_version) {
    check_supported(map, "VGAVisualLocal");
    const int64_t n = (int64_t)map.getAttributeTable().getNumRows();
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    if (m_gates_only) {
        writeAttributes(map, simple_version, nullptr, nullptr, nullptr, nullptr);
        return true;
    }
    vga_ctx *ctx = map.context();
    std::vector<int64_t> cluster((size_t)n);
    std::vector<int32_t> k((size_t)n), total((size_t)n);
    std::vector<float> control((size_t)n);
    int rc = vga_local(ctx, map.graph(), 0, n, cluster.data(), k.data(), total.data(), control.data());
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw RuntimeException(std::string("VGAVisualLocal: ") + vga_last_error());
    const std::vector<uint8_t> skip = map.contextSkipFlags();
    writeAttributes(map, simple_version, cluster.data(), k.data(), total.data(), control.data(), skip.empty() ? nullptr : skip.data());
    return true;
}

// vgavisualglobaldepth.cpp:27-28, 51-52, 71-73: every reached filled cell gets float(level); others keep -1
void VGAVisualGlobalDepth::writeAttributes(PointMap &map, const int32_t *depth) {
    AttributeTable &attributes = map.getAttributeTable();
    const int col = attributes.insertOrResetColumn("Visual Step Depth");
    const size_t n = attributes.getNumRows();
    // the reference writes level by level; the statistics are sums of small integers (exact in double in any
    // order), minimum and maximum
    for (size_t v = 0; v < n; v++)
        if (depth[v] >= 0) attributes.setValue(v, col, float(depth[v]));
    map.setDisplayedAttribute(-2);
    map.setDisplayedAttribute(col);
}

// vgavisualglobaldepth.cpp:23-75: BFS from the selection set
bool VGAVisualGlobalDepth::run(Communicator *, PointMap &map, bool) {
    check_supported(map, "VGAVisualGlobalDepth");
    const std::vector<int> &keys = map.getAttributeTable().keys();
    const int64_t n = (int64_t)keys.size();
    std::vector<int32_t> primary;
    vga_graph *graph = map.analysisGraph(&primary);
    std::vector<int64_t> sources;
    for (const PixelRef &sel : map.getSelSet()) {
        auto pos = std::lower_bound(keys.begin(), keys.end(), int(sel));
        if (pos == keys.end() || *pos != int(sel)) continue;
        const int64_t v = (int64_t)(pos - keys.begin());
        sources.push_back(primary.empty() ? v : (int64_t)primary[(size_t)v]);
    }
    const std::vector<uint8_t> skip = map.contextSkipFlags();
    if (!skip.empty() && !primary.empty())
        throw RuntimeException("VGAVisualGlobalDepth: context-filled cells together with merge links are not supported by the GPU path");
    if (vga_graph_set_noexpand(graph, skip.empty() ? nullptr : skip.data()) != VGA_OK)
        throw RuntimeException(std::string("VGAVisualGlobalDepth: ") + vga_last_error());
    std::vector<int32_t> depth((size_t)n, -1);
    int rc = vga_step_depth(map.context(), graph, sources.data(), (int64_t)sources.size(), depth.data());
    if (rc != VGA_OK) throw RuntimeException(std::string("VGAVisualGlobalDepth: ") + vga_last_error());
    copy_from_primary(primary, depth.data());
    writeAttributes(map, depth.data());
    return true;
}

static std::string radius_suffix(double radius, bool by_width, double width) {
    // vgametric.cpp:33-41 (first test on the radius), vgaangular.cpp:30-38 (first test on the region's width)
    if (radius == -1.0) return std::string();
    char buf[64];
    const char *fmt = (by_width ? width > 100.0 : radius > 100.0) ? "%.f" : (width < 1.0 ? "%.4f" : "%.2f");
    std::snprintf(buf, sizeof buf, fmt, radius);
    return std::string(" R") + buf;
}

void VGAMetric::writeAttributes(PointMap &map, double radius, const float *mean_angle, const float *mean_path_dist,
                                const float *mean_line_dist, const float *node_count) {
    AttributeTable &attributes = map.getAttributeTable();
    const std::string rt = radius_suffix(radius, false, map.getRegion().width());
    const int mspa_col = attributes.insertOrResetColumn("Metric Mean Shortest-Path Angle" + rt);
    const int mspl_col = attributes.insertOrResetColumn("Metric Mean Shortest-Path Distance" + rt);
    const int dist_col = attributes.insertOrResetColumn("Metric Mean Straight-Line Distance" + rt);
    const int count_col = attributes.insertOrResetColumn("Metric Node Count" + rt);
    if (node_count) {
        const size_t n = attributes.getNumRows();
        for (size_t v = 0; v < n; v++) {
            attributes.setValue(v, mspa_col, mean_angle[v]);
            attributes.setValue(v, mspl_col, mean_path_dist[v]);
            attributes.setValue(v, dist_col, mean_line_dist[v]);
            attributes.setValue(v, count_col, node_count[v]);
        }
    }
    map.overrideDisplayedAttribute(-2);
    map.setDisplayedAttribute(mspl_col);
}

bool VGAMetric::run(Communicator *comm, PointMap &map, bool) {
    check_supported(map, "VGAMetric");
    const int64_t n = (int64_t)map.getAttributeTable().getNumRows();
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    if (m_gates_only) {  // the reference skips every cell (vgametric.cpp:65-68): columns only
        writeAttributes(map, m_radius, nullptr, nullptr, nullptr, nullptr);
        return true;
    }
    const std::vector<uint8_t> flags = map.blockedAdjacentFlags();
    const std::vector<int32_t> partner = map.mergePartners();
    vga_ctx *ctx = map.context();
    CbState cb{comm, std::chrono::steady_clock::now(), false};
    vga_ctx_set_callbacks(ctx, progress_cb, cancel_cb, &cb);
    std::vector<float> angle((size_t)n), path((size_t)n), line((size_t)n), count((size_t)n);
    const int rc = vga_metric(ctx, map.graph(), flags.data(), partner.empty() ? nullptr : partner.data(), map.getSpacing(), m_radius,
                              nullptr, n, angle.data(), path.data(), line.data(), count.data(), nullptr);
    vga_ctx_set_callbacks(ctx, nullptr, nullptr, nullptr);
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw RuntimeException(std::string("VGAMetric: ") + vga_last_error());
    writeAttributes(map, m_radius, angle.data(), path.data(), line.data(), count.data());
    return true;
}

void VGAAngular::writeAttributes(PointMap &map, double radius, const float *mean_depth, const float *total_depth,
                                 const float *node_count) {
    AttributeTable &attributes = map.getAttributeTable();
    const std::string rt = radius_suffix(radius, true, map.getRegion().width());
    // getOrInsertColumn: existing columns keep their values and statistics (vgaangular.cpp:44-52)
    const int mean_depth_col = attributes.getOrInsertColumn("Angular Mean Depth" + rt);
    const int total_depth_col = attributes.getOrInsertColumn("Angular Total Depth" + rt);
    const int count_col = attributes.getOrInsertColumn("Angular Node Count" + rt);
    if (node_count) {
        const size_t n = attributes.getNumRows();
        for (size_t v = 0; v < n; v++) {
            if (node_count[v] > 0.0f) attributes.setValue(v, mean_depth_col, mean_depth[v]);
            attributes.setValue(v, total_depth_col, total_depth[v]);
            attributes.setValue(v, count_col, node_count[v]);
        }
    }
    map.setDisplayedAttribute(-2);
    map.setDisplayedAttribute(mean_depth_col);
}

bool VGAAngular::run(Communicator *comm, PointMap &map, bool) {
    check_supported(map, "VGAAngular");
    const int64_t n = (int64_t)map.getAttributeTable().getNumRows();
    if (comm) comm->CommPostMessage(Communicator::NUM_RECORDS, map.getFilledPointCount());
    if (m_gates_only) {
        writeAttributes(map, m_radius, nullptr, nullptr, nullptr);
        return true;
    }
    const std::vector<uint8_t> flags = map.blockedAdjacentFlags();
    const std::vector<int32_t> partner = map.mergePartners();
    vga_ctx *ctx = map.context();
    CbState cb{comm, std::chrono::steady_clock::now(), false};
    vga_ctx_set_callbacks(ctx, progress_cb, cancel_cb, &cb);
    std::vector<float> mean((size_t)n), total((size_t)n), count((size_t)n);
    const int rc = vga_angular(ctx, map.graph(), flags.data(), partner.empty() ? nullptr : partner.data(), m_radius, nullptr, n,
                               mean.data(), total.data(), count.data(), nullptr);
    vga_ctx_set_callbacks(ctx, nullptr, nullptr, nullptr);
    if (rc == VGA_ERR_CANCELLED) throw Communicator::CancelledException();
    if (rc != VGA_OK) throw RuntimeException(std::string("VGAAngular: ") + vga_last_error());
    writeAttributes(map, m_radius, mean.data(), total.data(), count.data());
    return true;
}

}  // namespace dmx
