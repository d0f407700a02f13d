// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin C-ABI harness around the UNMODIFIED reference sources (compiled where they
// lie under /root/reference by oracle/Makefile into oracle/_ref/libdmxref.so).
// It drives the reference's own entry points
//   PointMap::setGrid / makePoints / sparkGraph2   (salalib/pointdata.cpp:122,402,1246)
//   VGAVisualGlobal::run / VGAVisualLocal::run     (salalib/vgamodules/*.cpp:23)
// and flattens their results so that tests can (i) pin the C restatement in
// oracle/vga_oracle.c, (ii) generate tests/golden fixtures, and (iii) time the
// reference on host cores for bench.py's reference arm.
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
// may load this library.
//
// The harness itself is our code; it contains no reference source.  It needs access to
// a few protected members (Point::m_lines, Point::m_node) for dumping and for the
// sampled per-source timing, hence the access override below, confined to this TU.

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
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
#include <queue>
#include <stack>
#include <functional>
#include <exception>
#include <iomanip>
#include <cmath>
#include <ctime>
#include <cfloat>
#include <numeric>
#include <optional>
#include <unordered_map>
#include <unordered_set>
#include <typeinfo>
#include <cassert>
#include <random>
#include <mutex>
#include <thread>
#include <limits>
#include <array>
#include <tuple>
#include <utility>
#include <iterator>

#define protected public
#define private public
#include "salalib/mgraph.h"
#include "salalib/pointdata.h"
#include "salalib/ngraph.h"
#include "salalib/vgamodules/vgavisualglobal.h"
#include "salalib/vgamodules/vgavisuallocal.h"
#include "salalib/vgamodules/vgavisualglobaldepth.h"
#include "salalib/vgamodules/vgametric.h"
#include "salalib/vgamodules/vgaangular.h"
#undef protected
#undef private

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

struct Ref {
    std::vector<SpacePixelFile> drawing;
    QtRegion region;
    std::unique_ptr<PointMap> owned;
    std::unique_ptr<MetaGraph> mgraph;  // set when the map came from a .graph file
    PointMap *map = nullptr;
    std::string err;
};

double now_s() {
    using namespace std::chrono;
    return duration<double>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// walls: 4 doubles per segment (x1,y1,x2,y2).  Mirrors salaTest/testpointmap.cpp:321-341.
void *dmxref_create(const double *walls, int nwalls, double spacing) {
    Ref *r = new Ref();
    r->drawing.emplace_back("drawing");
    r->drawing.back().m_spacePixels.emplace_back("walls");
    ShapeMap &sm = r->drawing.back().m_spacePixels.back();
    for (int i = 0; i < nwalls; i++) {
        sm.makeLineShape(Line(Point2f(walls[4 * i], walls[4 * i + 1]), Point2f(walls[4 * i + 2], walls[4 * i + 3])));
    }
    r->drawing.back().m_region = sm.getRegion();
    r->region = r->drawing.back().m_region;
    r->owned.reset(new PointMap(r->region, r->drawing, "map"));
    r->map = r->owned.get();
    r->map->setGrid(spacing, Point2f(0, 0));
    return r;
}

void dmxref_destroy(void *h) { delete static_cast<Ref *>(h); }

// ---- .graph files through the reference's own MetaGraph (SURVEY.md §8 row f2 checks) --------------------

// MetaGraph::readFromFile; the handle then works with every dmxref_* accessor on the displayed point map.
void *dmxref_graph_open(const char *path) {
    Ref *r = new Ref();
    r->mgraph.reset(new MetaGraph());
    if (r->mgraph->readFromFile(path) != MetaGraph::OK || r->mgraph->getPointMaps().empty()) {
        delete r;
        return nullptr;
    }
    r->map = &r->mgraph->getDisplayedPointMap();
    return r;
}

// MetaGraph::write(filename, METAGRAPH_VERSION, false) of a handle from dmxref_graph_open
int dmxref_graph_save(void *h, const char *path) {
    Ref *r = static_cast<Ref *>(h);
    if (!r->mgraph) return 0;
    return r->mgraph->write(path, METAGRAPH_VERSION, false) == MetaGraph::OK ? 1 : 0;
}

// read + write with nothing in between (also for files without point maps): what the reference itself makes of
// a file it loads, e.g. the SELECTED flags it drops and the displayed-attribute index it re-maps
int dmxref_graph_rewrite(const char *in, const char *out) {
    MetaGraph mg;
    if (mg.readFromFile(in) != MetaGraph::OK) return 0;
    return mg.write(out, METAGRAPH_VERSION, false) == MetaGraph::OK ? 1 : 0;
}

void dmxref_grid(void *h, int *cols, int *rows, double *spacing, double *blx, double *bly) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    *cols = (int)m.getCols();
    *rows = (int)m.getRows();
    *spacing = m.getSpacing();
    *blx = m.m_bottom_left.x;
    *bly = m.m_bottom_left.y;
}

// PointMap::mergePixels on the cells containing the two points (what -m LINK -lnk does, salalib/linkutils.cpp:20-98)
int dmxref_merge(void *h, double ax, double ay, double bx, double by) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    PixelRef a = m.pixelate(Point2f(ax, ay), false), b = m.pixelate(Point2f(bx, by), false);
    if (!m.includes(a) || !m.includes(b) || !m.getPoint(a).filled() || !m.getPoint(b).filled()) return 0;
    if (m.isPixelMerged(a) || m.isPixelMerged(b)) return 0;
    return m.mergePixels(a, b) ? 1 : 0;
}

int dmxref_fill(void *h, double x, double y) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    return m.makePoints(Point2f(x, y), 0) ? 1 : 0;
}

// makePoints with a fill type: 0 = full fill, 1 = semi-fill (FILLED | CONTEXTFILLED, the GUI's "context fill")
int dmxref_fill_type(void *h, double x, double y, int fill_type) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    return m.makePoints(Point2f(x, y), fill_type) ? 1 : 0;
}

int dmxref_block_lines(void *h) { return static_cast<Ref *>(h)->map->blockLines() ? 1 : 0; }

int dmxref_filled_count(void *h) { return static_cast<Ref *>(h)->map->getFilledPointCount(); }

// state[x*rows+y] = Point::m_state (x-major, the order of every hot loop)
void dmxref_state(void *h, uint16_t *state) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    size_t rows = m.getRows(), cols = m.getCols();
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++)
            state[x * rows + y] = (uint16_t)m.getPoint(PixelRef((short)x, (short)y)).m_state;
}

// per-cell cropped wall lines: line_off[cols*rows+1]; lines = 5 doubles (blx,bly,trx,try,parity).
// Pass lines == nullptr to only count.  Returns total number of segments.
int64_t dmxref_cell_lines(void *h, uint32_t *line_off, double *lines) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    size_t rows = m.getRows(), cols = m.getCols();
    int64_t n = 0;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++) {
            if (line_off) line_off[x * rows + y] = (uint32_t)n;
            for (const Line &l : m.getPoint(PixelRef((short)x, (short)y)).m_lines) {
                if (lines) {
                    double *o = lines + 5 * n;
                    o[0] = l.bottom_left.x;
                    o[1] = l.bottom_left.y;
                    o[2] = l.top_right.x;
                    o[3] = l.top_right.y;
                    o[4] = l.bits.parity ? 1.0 : 0.0;
                }
                n++;
            }
        }
    if (line_off) line_off[rows * cols] = (uint32_t)n;
    return n;
}

// Runs the reference's sparkGraph2; returns seconds (negative on failure).
double dmxref_makegraph(void *h, int boundarygraph, double maxdist) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    double t0 = now_s();
    bool ok = m.sparkGraph2(nullptr, boundarygraph != 0, maxdist);
    double t1 = now_s();
    return ok ? (t1 - t0) : -1.0;
}

// Iterated adjacency (Node::first/next order: bin 0..31, runs in stored order), per filled
// cell in x-major order.  ref = int(PixelRef) = (x<<16)+(y&0xffff).  nullptr = count only.
int64_t dmxref_edges(void *h, uint64_t *rowptr, int32_t *ref, uint8_t *bin) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    size_t rows = m.getRows(), cols = m.getCols();
    int64_t n = 0;
    size_t v = 0;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++) {
            Point &p = m.getPoint(PixelRef((short)x, (short)y));
            if (!p.filled()) continue;
            if (rowptr) rowptr[v] = (uint64_t)n;
            v++;
            if (!p.hasNode()) continue;
            Node &node = p.getNode();
            for (int b = 0; b < 32; b++) {
                const Bin &bn = node.bin(b);
                bn.first();
                while (!bn.is_tail()) {
                    if (ref) ref[n] = (int32_t)(int)bn.cursor();
                    if (bin) bin[n] = (uint8_t)b;
                    n++;
                    bn.next();
                }
            }
        }
    if (rowptr) rowptr[v] = (uint64_t)n;
    return n;
}

// Per filled cell (x-major): 32 bin node counts (as stored: unsigned short) and far distances,
// plus the grid-connection byte.
void dmxref_bins(void *h, uint16_t *count /*N*32*/, float *dist /*N*32*/, uint8_t *gridconn /*N*/) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    size_t rows = m.getRows(), cols = m.getCols();
    size_t v = 0;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++) {
            Point &p = m.getPoint(PixelRef((short)x, (short)y));
            if (!p.filled()) continue;
            for (int b = 0; b < 32; b++) {
                count[v * 32 + b] = p.hasNode() ? (uint16_t)p.getNode().bin(b).count() : 0;
                dist[v * 32 + b] = p.hasNode() ? p.getNode().bin(b).distance() : 0.0f;
            }
            gridconn[v] = (uint8_t)p.m_grid_connections;
            v++;
        }
}

// Attribute column by name for every filled cell in x-major order; returns 0 if column missing.
int dmxref_attr(void *h, const char *name, float *out) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    AttributeTable &t = m.getAttributeTable();
    if (!t.hasColumn(name)) return 0;
    size_t col = t.getColumnIndex(name);
    size_t rows = m.getRows(), cols = m.getCols();
    size_t v = 0;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++) {
            PixelRef pr((short)x, (short)y);
            if (!m.getPoint(pr).filled()) continue;
            out[v++] = t.getRow(AttributeKey(pr)).getValue(col);
        }
    return 1;
}

// Column names, '\n'-separated, into buf; returns number of columns.
int dmxref_columns(void *h, char *buf, int buflen) {
    AttributeTable &t = static_cast<Ref *>(h)->map->getAttributeTable();
    std::string s;
    for (size_t i = 0; i < t.getNumColumns(); i++) {
        s += t.getColumnName(i);
        s += "\n";
    }
    strncpy(buf, s.c_str(), buflen - 1);
    buf[buflen - 1] = 0;
    return (int)t.getNumColumns();
}

double dmxref_vga_global(void *h, double radius, int simple) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    double t0 = now_s();
    bool ok = VGAVisualGlobal(radius, false).run(nullptr, m, simple != 0);
    double t1 = now_s();
    return ok ? (t1 - t0) : -1.0;
}

double dmxref_vga_local(void *h, int simple) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    double t0 = now_s();
    bool ok = VGAVisualLocal(false).run(nullptr, m, simple != 0);
    double t1 = now_s();
    return ok ? (t1 - t0) : -1.0;
}

// Metric / angular VGA: the reference's VGAMetric::run / VGAAngular::run (salalib/vgamodules/vgametric.cpp,
// vgaangular.cpp) unmodified; results are the "Metric ..." / "Angular ..." columns via dmxref_attr.
double dmxref_vga_metric(void *h, double radius) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    double t0 = now_s();
    bool ok = VGAMetric(radius, false).run(nullptr, m, false);
    double t1 = now_s();
    return ok ? (t1 - t0) : -1.0;
}

double dmxref_vga_angular(void *h, double radius) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    double t0 = now_s();
    bool ok = VGAAngular(radius, false).run(nullptr, m, false);
    double t1 = now_s();
    return ok ? (t1 - t0) : -1.0;
}

// Visual step depth from a selection of cells (x-major ordinals of filled cells): drives the reference's
// VGAVisualGlobalDepth::run (salalib/vgamodules/vgavisualglobaldepth.cpp:23-75) through the map's
// selection set, as the CLI's -m STEPDEPTH does.  Result: column "Visual Step Depth" via dmxref_attr.
double dmxref_step_depth(void *h, const int32_t *src, int k) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    size_t rows = m.getRows(), cols = m.getCols();
    std::vector<PixelRef> ord;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++)
            if (m.getPoint(PixelRef((short)x, (short)y)).filled()) ord.push_back(PixelRef((short)x, (short)y));
    m.getSelSet().clear();
    for (int i = 0; i < k; i++) m.getSelSet().insert((int)ord[src[i]]);
    double t0 = now_s();
    VGAVisualGlobalDepth d;
    bool ok = d.run(nullptr, m, false);
    double t1 = now_s();
    m.getSelSet().clear();
    return ok ? (t1 - t0) : -1.0;
}

// ---- sampled per-source timing (SURVEY.md §8d "sampled baseline") -------------------------

// Full per-source construction cost for K sampled sources: exactly the body of the reference's
// source loop (pointdata.cpp:1288-1292: new Node, addRow, sparkPixel2(curs,1,maxdist)), which
// includes Node::make / Bin::make.  Must be called after fill and before dmxref_makegraph.
// src = x-major ordinals of filled cells.  Returns seconds for the K sources.
double dmxref_sample_makegraph(void *h, const int32_t *src, int k, double maxdist, int64_t *edges_out) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    if (!m.m_blockedlines) m.blockLines();
    size_t rows = m.getRows(), cols = m.getCols();
    std::vector<PixelRef> ord;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++)
            if (m.getPoint(PixelRef((short)x, (short)y)).filled()) ord.push_back(PixelRef((short)x, (short)y));
    m.m_attributes->insertOrResetLockedColumn("Connectivity");
    m.m_attributes->insertOrResetColumn("Point First Moment");
    m.m_attributes->insertOrResetColumn("Point Second Moment");
    m.tagState(true);
    int64_t edges = 0;
    double t0 = now_s();
    for (int i = 0; i < k; i++) {
        PixelRef curs = ord[src[i]];
        m.getPoint(curs).m_node = std::unique_ptr<Node>(new Node());
        // a source sampled again (bench.py draws a fresh sample every step) keeps its row: the reference's addRow
        // throws on a duplicate key (attributetable.cpp:278); sparkGraph2 itself adds each row once
        if (!m.m_attributes->getRowPtr(AttributeKey(curs))) m.m_attributes->addRow(AttributeKey(curs));
        m.getPoint(curs).m_processflag = 0x00FF;
        m.sparkPixel2(curs, 1, maxdist);
    }
    double t1 = now_s();
    for (int i = 0; i < k; i++) edges += m.getPoint(ord[src[i]]).getNode().count();
    m.tagState(false);
    if (edges_out) *edges_out = edges;
    return t1 - t0;
}

// ---- the graph of a large plan built by several processes (bench.py --impl reference on the 10^6-cell workload) -----
// The reference is single-threaded and non-reentrant (function-local statics in sparkPixel2), but independent PROCESSES
// can each make the Nodes of a share of the sources (the body of sparkGraph2's source loop, as in
// dmxref_sample_makegraph) and hand them over through the reference's own Node serialisation (Node::write / Node::read,
// salalib/ngraph.cpp:195-220 -- what a .graph file holds).  File: int32 count, then per source int32 ordinal + Node.

static std::vector<PixelRef> filled_order(PointMap &m) {
    size_t rows = m.getRows(), cols = m.getCols();
    std::vector<PixelRef> ord;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++)
            if (m.getPoint(PixelRef((short)x, (short)y)).filled()) ord.push_back(PixelRef((short)x, (short)y));
    return ord;
}

// Makes the Nodes of the listed sources and writes them to `path`; the Nodes are dropped again afterwards (the caller
// loads all parts with dmxref_load_nodes).  Returns seconds spent in the construction, negative on failure.
double dmxref_build_nodes_to_file(void *h, const int32_t *src, int k, double maxdist, const char *path) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    if (!m.m_blockedlines) m.blockLines();
    std::vector<PixelRef> ord = filled_order(m);
    m.m_attributes->insertOrResetLockedColumn("Connectivity");
    m.m_attributes->insertOrResetColumn("Point First Moment");
    m.m_attributes->insertOrResetColumn("Point Second Moment");
    m.tagState(true);
    std::ofstream out(path, std::ios::binary);
    if (!out) return -1.0;
    int32_t count = k;
    out.write((const char *)&count, sizeof(count));
    double spent = 0.0;
    for (int i = 0; i < k; i++) {
        PixelRef curs = ord[src[i]];
        double t0 = now_s();
        m.getPoint(curs).m_node = std::unique_ptr<Node>(new Node());
        if (!m.m_attributes->getRowPtr(AttributeKey(curs))) m.m_attributes->addRow(AttributeKey(curs));
        m.getPoint(curs).m_processflag = 0x00FF;
        m.sparkPixel2(curs, 1, maxdist);
        spent += now_s() - t0;
        int32_t o = src[i];
        out.write((const char *)&o, sizeof(o));
        m.getPoint(curs).getNode().write(out);
        m.getPoint(curs).m_node.reset();
    }
    m.tagState(false);
    out.close();
    return out ? spent : -1.0;
}

// Reads a file written by dmxref_build_nodes_to_file into the map's Points.  Returns the number of Nodes read, -1 on failure.
int64_t dmxref_load_nodes(void *h, const char *path) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    std::vector<PixelRef> ord = filled_order(m);
    std::ifstream in(path, std::ios::binary);
    if (!in) return -1;
    int32_t count = 0;
    in.read((char *)&count, sizeof(count));
    for (int32_t i = 0; i < count; i++) {
        int32_t o = -1;
        in.read((char *)&o, sizeof(o));
        if (!in || o < 0 || (size_t)o >= ord.size()) return -1;
        Point &p = m.getPoint(ord[o]);
        p.m_node = std::unique_ptr<Node>(new Node());
        p.m_node->read(in);
        if (!in) return -1;
    }
    return count;
}

// number of filled cells that hold a Node
int64_t dmxref_node_count(void *h) {
    PointMap &m = *static_cast<Ref *>(h)->map;
    int64_t c = 0;
    for (const PixelRef &p : filled_order(m))
        if (m.getPoint(p).m_node) c++;
    return c;
}

// Global BFS for K sampled sources on a made graph: the per-source body of
// VGAVisualGlobal::run (vgavisualglobal.cpp:80-130, no merges) re-stated around the
// reference's own public VGAVisualGlobal::extractUnseen.  Outputs integers per source.
double dmxref_sample_global(void *h, const int32_t *src, int k, int radius, int32_t *total_nodes,
                            int64_t *total_depth) {
    PointMap &map = *static_cast<Ref *>(h)->map;
    size_t rows = map.getRows(), cols = map.getCols();
    std::vector<PixelRef> ord;
    for (size_t x = 0; x < cols; x++)
        for (size_t y = 0; y < rows; y++)
            if (map.getPoint(PixelRef((short)x, (short)y)).filled()) ord.push_back(PixelRef((short)x, (short)y));
    VGAVisualGlobal g((double)radius, false);
    depthmapX::RowMatrix<int> miscs(rows, cols);
    depthmapX::RowMatrix<PixelRef> extents(rows, cols);
    double t0 = now_s();
    for (int s = 0; s < k; s++) {
        PixelRef curs = ord[src[s]];
        for (size_t ii = 0; ii < cols; ii++)
            for (size_t jj = 0; jj < rows; jj++) {
                miscs(jj, ii) = 0;
                extents(jj, ii) = PixelRef((short)ii, (short)jj);
            }
        int td = 0, tn = 0;
        std::vector<PixelRefVector> tree;
        tree.push_back(PixelRefVector());
        tree.back().push_back(curs);
        int level = 0;
        while (tree[level].size()) {
            tree.push_back(PixelRefVector());
            const PixelRefVector &cur = tree[level];
            for (auto it = cur.rbegin(); it != cur.rend(); ++it) {
                int &pmisc = miscs(it->y, it->x);
                Point &p = map.getPoint(*it);
                if (p.filled() && pmisc != ~0) {
                    td += level;
                    tn += 1;
                    if (radius == -1 || level < radius) {
                        g.extractUnseen(p.getNode(), tree[level + 1], miscs, extents);
                    }
                    pmisc = ~0;
                }
            }
            level++;
        }
        if (total_nodes) total_nodes[s] = tn;
        if (total_depth) total_depth[s] = td;
    }
    double t1 = now_s();
    return t1 - t0;
}

}  // extern "C"
