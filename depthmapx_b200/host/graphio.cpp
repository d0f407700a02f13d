// .graph serialisation of the visibility-graph path: ByteReader, the attribute table,
// Point / Node / Bin / PixelVec records, PointMap::read / write, the Node encoder used after a GPU build, and
// the GraphFile container.  Every record cites the reference routine whose byte layout it follows; nothing here
// computes visibility -- the adjacency comes from libvga_b200 or from the file.
#include <algorithm>
#include <cstring>
#include <fstream>
#include <numeric>
#include <ostream>
#include <sstream>

#include "graphfile.h"
#include "merge_contract.h"
#include "pointmap.h"

namespace dmx {

namespace {
enum { NODIR = 0x00, HORIZONTAL = 0x01, VERTICAL = 0x02, POSDIAGONAL = 0x04, NEGDIAGONAL = 0x08, DIAGONAL = 0x0c };

template <typename T> void put(std::ostream &out, const T &v) { out.write(reinterpret_cast<const char *>(&v), sizeof(T)); }
void put_string(std::ostream &out, const std::string &s) {  // dXstring::writeString, genlib/stringutils.cpp:52-58
    put<uint32_t>(out, (uint32_t)s.size());
    if (!s.empty()) out.write(s.data(), (std::streamsize)s.size());
}
// direction class of bin i (Node::make, ngraph.cpp:41-52)
inline int bin_dir(int i) {
    if (i == 4 || i == 20) return POSDIAGONAL;
    if (i == 12 || i == 28) return NEGDIAGONAL;
    if ((i > 4 && i < 12) || (i > 20 && i < 28)) return VERTICAL;
    return HORIZONTAL;
}
// the coordinate a run advances along (PixelRef::col, pixelref.h:47-54) and the step (PixelRef::move :55-66)
inline short run_col(const PixelRef &p, int dir) { return (dir & VERTICAL) ? p.y : p.x; }
inline void run_move(PixelRef &p, int dir) {
    switch (dir) {
    case POSDIAGONAL: p.x++; p.y++; break;
    case NEGDIAGONAL: p.x++; p.y--; break;
    case HORIZONTAL: p.x++; break;
    case VERTICAL: p.y++; break;
    default: break;
    }
}
}  // namespace

// ------------------------------------------------------------------------------------ ByteReader

void ByteReader::raw(void *dst, size_t n) {
    if (n > m_size - m_pos) throw RuntimeException("graph file: unexpected end of data");
    std::memcpy(dst, m_data + m_pos, n);
    m_pos += n;
}

void ByteReader::skip(size_t n) {
    if (n > m_size - m_pos) throw RuntimeException("graph file: unexpected end of data");
    m_pos += n;
}

void ByteReader::expect(uint64_t count, uint64_t min_bytes_each) const {
    if (min_bytes_each && count > (m_size - m_pos) / min_bytes_each) throw RuntimeException("graph file: a count exceeds the data that is left");
}

std::string ByteReader::str() {  // dXstring::readString, genlib/stringutils.cpp:40-50
    const uint32_t len = get<uint32_t>();
    if (len > m_size - m_pos) throw RuntimeException("graph file: string runs past the end of data");
    std::string s(m_data + m_pos, len);
    m_pos += len;
    return s;
}

void NodeStore::clear() {
    node_of_cell.clear();
    bins.clear();
    runs.clear();
    occl_off.clear();
    occl.clear();
}

// ------------------------------------------------------------------------------------ attribute table

// skip / capture a serialised LayerManagerImpl (layermanagerimpl.cpp:90-104): two int64, count, (int64 key + name)*
static std::string read_layers(ByteReader &in) {
    const size_t begin = in.pos();
    in.skip(16);
    const int32_t count = in.get<int32_t>();
    for (int i = 0; i < count; i++) {
        in.skip(8);
        in.str();
    }
    return std::string(in.at(begin), in.pos() - begin);
}

// AttributeTable::read (attributetable.cpp:397-425), AttributeColumnImpl::read (:91-109), AttributeRowImpl::read
// (:190-194).  Columns are stored in name order together with their physical index; memory keeps physical order.
bool AttributeTable::read(ByteReader &in) {
    clear();
    m_layers = read_layers(in);
    const int32_t colcount = in.get<int32_t>();
    if (colcount < 0) throw RuntimeException("graph file: negative column count");
    in.expect((uint64_t)colcount, 4 + 4 + 4 + 8 + 4 + 1 + 1 + sizeof(DisplayParams) + 4);
    std::map<size_t, Column> tmp;
    for (int j = 0; j < colcount; j++) {
        Column c;
        c.name = in.str();
        c.min = in.get<float>();
        c.max = in.get<float>();
        c.total = in.get<double>();
        const int32_t physical = in.get<int32_t>();
        c.hidden = in.get<uint8_t>() != 0;
        c.locked = in.get<uint8_t>() != 0;
        in.raw(&c.display, sizeof(DisplayParams));
        c.formula = in.str();
        tmp[(size_t)physical] = c;
    }
    for (auto &kv : tmp) m_columns.push_back(kv.second);
    m_cols.assign(m_columns.size(), std::vector<float>());
    const int32_t rowcount = in.get<int32_t>();
    if (rowcount < 0) throw RuntimeException("graph file: negative row count");
    in.expect((uint64_t)rowcount, 4 + 8 + 4 + 4 * m_columns.size());
    m_keys.resize((size_t)rowcount);
    m_layer_keys.resize((size_t)rowcount);
    for (auto &c : m_cols) c.resize((size_t)rowcount);
    std::vector<float> row;
    for (int32_t r = 0; r < rowcount; r++) {
        m_keys[(size_t)r] = in.get<int32_t>();
        m_layer_keys[(size_t)r] = in.get<int64_t>();
        const uint32_t n = in.get<uint32_t>();
        if (n != m_columns.size()) throw RuntimeException("graph file: attribute row width differs from the column count");
        row.resize(n);
        if (n) in.raw(row.data(), n * sizeof(float));
        for (uint32_t c = 0; c < n; c++) m_cols[c][(size_t)r] = row[c];
    }
    if (!std::is_sorted(m_keys.begin(), m_keys.end()))
        throw RuntimeException("graph file: attribute rows are not in key order");
    in.raw(&m_display, sizeof(DisplayParams));
    return true;
}

// AttributeTable::write (attributetable.cpp:427-457), AttributeColumnImpl::write (:111-124)
void AttributeTable::write(std::ostream &out) const {
    out.write(m_layers.data(), (std::streamsize)m_layers.size());
    put<int32_t>(out, (int32_t)m_columns.size());
    std::vector<size_t> order(m_columns.size());
    std::iota(order.begin(), order.end(), (size_t)0);
    std::sort(order.begin(), order.end(), [&](size_t a, size_t b) { return m_columns[a].name < m_columns[b].name; });
    for (size_t idx : order) {
        const Column &c = m_columns[idx];
        put_string(out, c.name);
        put<float>(out, (float)c.min);
        put<float>(out, (float)c.max);
        put<double>(out, c.total);
        put<int32_t>(out, (int32_t)idx);
        put<uint8_t>(out, c.hidden ? 1 : 0);
        put<uint8_t>(out, c.locked ? 1 : 0);
        put(out, c.display);
        put_string(out, c.formula);
    }
    put<int32_t>(out, (int32_t)m_keys.size());
    std::vector<float> row(m_columns.size());
    for (size_t r = 0; r < m_keys.size(); r++) {
        put<int32_t>(out, m_keys[r]);
        put<int64_t>(out, m_layer_keys[r]);
        put<uint32_t>(out, (uint32_t)row.size());
        for (size_t c = 0; c < row.size(); c++) row[c] = m_cols[c][r];
        if (!row.empty()) out.write(reinterpret_cast<const char *>(row.data()), (std::streamsize)(row.size() * sizeof(float)));
    }
    put(out, m_display);
}

// ------------------------------------------------------------------------------------ Node codec

// Node::read -> 32 x Bin::read (ngraph.cpp:195-207, 420-445), PixelVec::read (:491-516, 541-563), then 32 occlusion
// vectors
static void read_node(ByteReader &in, NodeStore &ns) {
    for (int i = 0; i < 32; i++) {
        NodeStore::Bin b;
        b.dir = in.get<signed char>();
        b.count = in.get<uint16_t>();
        b.distance = in.get<float>();
        b.occ_distance = in.get<float>();
        b.first_run = (uint32_t)ns.runs.size();
        if (b.count) {
            if (b.dir & DIAGONAL) {
                NodeStore::Run r;
                r.start = in.get<PixelRef>();
                const uint16_t len = in.get<uint16_t>();
                r.end.x = (short)(r.start.x + len);
                r.end.y = (short)(b.dir == POSDIAGONAL ? r.start.y + len : r.start.y - len);
                ns.runs.push_back(r);
            } else {
                const uint16_t length = in.get<uint16_t>();
                if (length == 0) throw RuntimeException("graph file: bin with nodes but no runs");
                for (uint16_t k = 0; k < length; k++) {
                    NodeStore::Run r;
                    if (k == 0) {
                        r.start = in.get<PixelRef>();
                        const uint16_t len = in.get<uint16_t>();
                        r.end = r.start;
                        if (b.dir == VERTICAL)
                            r.end.y = (short)(r.start.y + len);
                        else
                            r.end.x = (short)(r.start.x + len);
                    } else {
                        const short primary = in.get<short>();
                        const uint16_t sl = in.get<uint16_t>();  // struct { shift:4; runlength:12; }
                        const int shift = sl & 0xf, len = sl >> 4;
                        const NodeStore::Run &prev = ns.runs.back();
                        if (b.dir == VERTICAL) {
                            r.start.x = (short)(prev.start.x + shift);
                            r.start.y = primary;
                            r.end.x = r.start.x;
                            r.end.y = (short)(r.start.y + len);
                        } else {
                            r.start.x = primary;
                            r.start.y = (short)(prev.start.y + shift);
                            r.end.x = (short)(r.start.x + len);
                            r.end.y = r.start.y;
                        }
                    }
                    ns.runs.push_back(r);
                }
            }
        }
        b.nruns = (uint32_t)ns.runs.size() - b.first_run;
        ns.bins.push_back(b);
    }
    for (int i = 0; i < 32; i++) {
        const uint32_t n = in.get<uint32_t>();
        in.expect(n, sizeof(PixelRef));
        ns.occl_off.push_back((uint32_t)ns.occl.size());
        for (uint32_t k = 0; k < n; k++) ns.occl.push_back(in.get<PixelRef>());
    }
}

// Node::write -> Bin::write (ngraph.cpp:209-220, 447-472), PixelVec::write (:518-534, 565-583)
static void write_node(std::ostream &out, const NodeStore &ns, size_t node) {
    for (int i = 0; i < 32; i++) {
        const NodeStore::Bin &b = ns.bins[node * 32 + (size_t)i];
        put<signed char>(out, b.dir);
        put<uint16_t>(out, b.count);
        put<float>(out, b.distance);
        put<float>(out, b.occ_distance);
        if (!b.count) continue;
        const NodeStore::Run *runs = ns.runs.data() + b.first_run;
        if (b.dir & DIAGONAL) {
            put(out, runs[0].start);
            put<uint16_t>(out, (uint16_t)(runs[0].end.x - runs[0].start.x));
        } else {
            put<uint16_t>(out, (uint16_t)b.nruns);
            put(out, runs[0].start);
            put<uint16_t>(out, (uint16_t)(b.dir == VERTICAL ? runs[0].end.y - runs[0].start.y : runs[0].end.x - runs[0].start.x));
            for (uint32_t k = 1; k < (uint16_t)b.nruns; k++) {
                const NodeStore::Run &r = runs[k], &prev = runs[k - 1];
                unsigned shift, len;
                if (b.dir == VERTICAL) {
                    put<short>(out, r.start.y);
                    len = (unsigned)(r.end.y - r.start.y);
                    shift = (unsigned)(r.start.x - prev.start.x);
                } else {
                    put<short>(out, r.start.x);
                    len = (unsigned)(r.end.x - r.start.x);
                    shift = (unsigned)(r.start.y - prev.start.y);
                }
                put<uint16_t>(out, (uint16_t)((shift & 0xf) | ((len & 0xfff) << 4)));
            }
        }
    }
    for (int i = 0; i < 32; i++) {
        const size_t o = node * 32 + (size_t)i;
        const uint32_t begin = ns.occl_off[o];
        const uint32_t end = o + 1 < ns.occl_off.size() ? ns.occl_off[o + 1] : (uint32_t)ns.occl.size();
        put<uint32_t>(out, end - begin);
        if (end > begin) out.write(reinterpret_cast<const char *>(ns.occl.data() + begin), (std::streamsize)((end - begin) * sizeof(PixelRef)));
    }
}

// ------------------------------------------------------------------------------------ PointMap::read / write

// PointMap::read (pointdata.cpp:1073-1152) with Point::read (point.cpp:25-49)
bool PointMap::read(ByteReader &in) {
    adoptGraph(nullptr);
    m_name = in.str();
    m_displayed_attribute = -1;
    m_spacing = in.get<double>();
    const int32_t rows = in.get<int32_t>(), cols = in.get<int32_t>();
    if (rows < 0 || cols < 0 || rows > 32767 || cols > 32767) throw RuntimeException("graph file: bad grid size");
    in.expect((uint64_t)rows * (uint64_t)cols, 34);  // a Point record without a Node is 34 bytes
    m_rows = (size_t)rows;
    m_cols = (size_t)cols;
    m_filled_point_count = in.get<int32_t>();
    m_bottom_left = in.get<Point2f>();
    m_region = Region(Point2f(m_bottom_left.x - m_spacing / 2.0, m_bottom_left.y - m_spacing / 2.0),
                      Point2f(m_bottom_left.x + double(m_cols - 1) * m_spacing + m_spacing / 2.0,
                              m_bottom_left.y + double(m_rows - 1) * m_spacing + m_spacing / 2.0));
    const int32_t displayed_attribute = in.get<int32_t>();
    m_attributes.read(in);
    m_points.assign(m_cols * m_rows, Point());
    m_nodes.clear();
    m_nodes.node_of_cell.assign(m_cols * m_rows, -1);
    const int keep = Point::EMPTY | Point::FILLED | Point::MERGED | Point::BLOCKED | Point::CONTEXTFILLED | Point::EDGE;
    for (size_t c = 0; c < m_cols * m_rows; c++) {
        Point &pt = m_points[c];
        pt.state = in.get<int32_t>();
        pt.block = in.get<int32_t>();
        in.skip(4);  // dummy
        pt.grid_connections = in.get<uint8_t>();
        pt.merge = in.get<PixelRef>();
        const bool has_node = in.get<uint8_t>() != 0;
        if (has_node) {
            m_nodes.node_of_cell[c] = (int32_t)m_nodes.numNodes();
            read_node(in, m_nodes);
        }
        pt.location = in.get<Point2f>();
        pt.state &= keep;  // drops the SELECTED flag etc. (pointdata.cpp:1126)
    }
    m_nodes.occl_off.push_back((uint32_t)m_nodes.occl.size());
    m_nodes_valid = true;
    {
        int filled = 0;
        for (const Point &pt : m_points) filled += pt.filled() ? 1 : 0;
        if (filled != m_filled_point_count) throw RuntimeException("graph file: the stored filled-point count does not match the cell states");
        // attribute rows belong to filled cells that have a Node, in key order
        for (int key : m_attributes.keys()) {
            const PixelRef p(key);
            if (!includes(p) || !getPoint(p).filled()) throw RuntimeException("graph file: attribute row of a cell that is not filled");
        }
    }
    m_selection_set.clear();
    m_has_selection = false;
    m_initialised = true;
    m_blockedlines = false;
    m_processed = in.get<uint8_t>() != 0;
    m_boundarygraph = in.get<uint8_t>() != 0;
    // the stored (name-order) index is taken as it is (pointdata.cpp:1109, 1148-1149)
    m_displayed_attribute = -2;
    setDisplayedAttribute(displayed_attribute);
    return true;
}

// PointMap::write (pointdata.cpp:1154-1188) with Point::write (point.cpp:51-73)
void PointMap::write(std::ostream &out) {
    const NodeStore &ns = nodes();
    put_string(out, m_name);
    put<double>(out, m_spacing);
    put<int32_t>(out, (int32_t)m_rows);
    put<int32_t>(out, (int32_t)m_cols);
    put<int32_t>(out, (int32_t)m_filled_point_count);
    put(out, m_bottom_left);
    put<int32_t>(out, (int32_t)m_attributes.getColumnSortedIndex(m_displayed_attribute));
    m_attributes.write(out);
    for (size_t c = 0; c < m_cols * m_rows; c++) {
        const Point &pt = m_points[c];
        put<int32_t>(out, pt.state);
        put<int32_t>(out, pt.block);
        put<int32_t>(out, 0);
        put<uint8_t>(out, pt.grid_connections);
        put(out, pt.merge);
        const int32_t node = ns.node_of_cell.empty() ? -1 : ns.node_of_cell[c];
        put<uint8_t>(out, node >= 0 ? 1 : 0);
        if (node >= 0) write_node(out, ns, (size_t)node);
        put(out, pt.location);
    }
    put<uint8_t>(out, m_processed ? 1 : 0);
    put<uint8_t>(out, m_boundarygraph ? 1 : 0);
}

// ------------------------------------------------------------------------------------ nodes <-> flat rows

// Node::first / next over Bin::first / next (ngraph.cpp:158-191, 392-416): bins 0..31, runs in order, each run
// from start while col <= end.col
void PointMap::flattenNodes(FlatRows &out) {
    const NodeStore &ns = nodes();
    out.rowptr.assign(1, 0);
    out.ref.clear();
    out.bin.clear();
    for (size_t c = 0; c < m_cols * m_rows; c++) {
        if (!m_points[c].filled()) continue;
        const int32_t node = ns.node_of_cell.empty() ? -1 : ns.node_of_cell[c];
        if (node >= 0) {
            for (int i = 0; i < 32; i++) {
                const NodeStore::Bin &b = ns.bins[(size_t)node * 32 + (size_t)i];
                for (uint32_t k = 0; k < b.nruns; k++) {
                    const NodeStore::Run &r = ns.runs[b.first_run + k];
                    if (b.dir == NODIR) continue;
                    // start .. end along the run's axis; counted in int so that a run ending at 32767 terminates
                    const int cells_in_run = (int)run_col(r.end, b.dir) - (int)run_col(r.start, b.dir) + 1;
                    PixelRef p = r.start;
                    for (int k2 = 0; k2 < cells_in_run; k2++, run_move(p, b.dir)) {
                        out.ref.push_back(int(p));
                        out.bin.push_back((uint8_t)i);
                    }
                }
            }
        }
        out.rowptr.push_back(out.ref.size());
    }
}

// Node::make / Bin::make (ngraph.cpp:27-58, 234-304) from the flat rows of all filled cells.  Horizontal bins are
// run-length encoded in (y, x) order, vertical bins in (x, y) order; a diagonal bin is one run from its
// smallest-x to its largest-x pixel (the sieve emits a diagonal in order of distance, so first / last are the
// extremes) and its stored count excludes the fill-in pixels.
void PointMap::encodeNodes(const uint64_t *rowptr, const int32_t *ref, const uint8_t *bin, const uint8_t *accepted,
                           const float *far_bin_dists) {
    m_nodes.clear();
    m_nodes.node_of_cell.assign(m_cols * m_rows, -1);
    std::vector<PixelRef> pix[32];
    std::vector<uint32_t> nacc(32);
    size_t v = 0;
    for (size_t c = 0; c < m_cols * m_rows; c++) {
        if (!m_points[c].filled()) continue;
        m_nodes.node_of_cell[c] = (int32_t)v;
        for (int i = 0; i < 32; i++) {
            pix[i].clear();
            nacc[(size_t)i] = 0;
        }
        for (uint64_t e = rowptr[v]; e < rowptr[v + 1]; e++) {
            const int b = bin[e];
            pix[b].push_back(PixelRef(ref[e]));
            if (!accepted || accepted[e]) nacc[(size_t)b]++;
        }
        for (int i = 0; i < 32; i++) {
            NodeStore::Bin b;
            b.distance = far_bin_dists ? far_bin_dists[v * 32 + (size_t)i] : 0.0f;
            b.first_run = (uint32_t)m_nodes.runs.size();
            std::vector<PixelRef> &p = pix[i];
            if (!p.empty()) {
                const int dir = bin_dir(i);
                b.dir = (signed char)dir;
                b.count = (uint16_t)nacc[(size_t)i];
                if (dir & DIAGONAL) {
                    auto mm = std::minmax_element(p.begin(), p.end(), [](const PixelRef &a, const PixelRef &c2) { return a.x < c2.x; });
                    m_nodes.runs.push_back(NodeStore::Run{*mm.first, *mm.second});
                } else {
                    if (dir == HORIZONTAL)
                        std::sort(p.begin(), p.end(), [](const PixelRef &a, const PixelRef &c2) { return a.y < c2.y || (a.y == c2.y && a.x < c2.x); });
                    else
                        std::sort(p.begin(), p.end(), [](const PixelRef &a, const PixelRef &c2) { return a.x < c2.x || (a.x == c2.x && a.y < c2.y); });
                    p.erase(std::unique(p.begin(), p.end(), [](const PixelRef &a, const PixelRef &c2) { return int(a) == int(c2); }), p.end());
                    NodeStore::Run cur{p[0], p[0]};
                    for (size_t k = 1; k < p.size(); k++) {
                        const bool joins = dir == HORIZONTAL ? (p[k - 1].y == p[k].y && p[k - 1].x + 1 == p[k].x)
                                                             : (p[k - 1].x == p[k].x && p[k - 1].y + 1 == p[k].y);
                        if (!joins) {
                            cur.end = p[k - 1];
                            m_nodes.runs.push_back(cur);
                            cur = NodeStore::Run{p[k], p[k]};
                        }
                    }
                    cur.end = p.back();
                    m_nodes.runs.push_back(cur);
                }
            }
            b.nruns = (uint32_t)m_nodes.runs.size() - b.first_run;
            m_nodes.bins.push_back(b);
        }
        for (int i = 0; i < 32; i++) m_nodes.occl_off.push_back(0);
        v++;
    }
    m_nodes.occl_off.push_back(0);
    m_nodes_valid = true;
}

// the run-length adjacency of the current graph; after a GPU build it is encoded from the device rows on demand
const NodeStore &PointMap::nodes() {
    if (m_nodes_valid || !m_graph) return m_nodes;
    const int64_t n = vga_graph_num_cells(m_graph), e = vga_graph_num_edges(m_graph);
    const int64_t total = n + vga_graph_num_ghosts(m_graph);
    std::vector<uint64_t> rowptr((size_t)n + 1);
    std::vector<uint32_t> col((size_t)e);
    std::vector<uint8_t> bin((size_t)e), acc((size_t)e);
    std::vector<int32_t> refs((size_t)total);
    std::vector<float> far((size_t)n * 32);
    if (vga_graph_csr(m_graph, rowptr.data(), col.data(), bin.data(), acc.data()) != VGA_OK ||
        vga_graph_cell_refs(m_graph, refs.data()) != VGA_OK ||
        vga_graph_node_stats(m_graph, nullptr, nullptr, nullptr, far.data(), nullptr, nullptr) != VGA_OK)
        throw RuntimeException(std::string("PointMap::nodes: ") + vga_last_error());
    std::vector<int32_t> ref((size_t)e);
    for (int64_t i = 0; i < e; i++) ref[(size_t)i] = refs[col[(size_t)i]];
    encodeNodes(rowptr.data(), ref.data(), bin.data(), acc.data(), far.data());
    return m_nodes;
}

// filled cells get their x-major ordinal, unfilled cells N + their x-major rank among the unfilled (the numbering
// vga_graph_build uses)
void PointMap::ordinals(std::vector<int32_t> &ord, int64_t &n, int64_t &ghosts) const {
    const size_t cells = m_cols * m_rows;
    ord.resize(cells);
    n = 0;
    ghosts = 0;
    for (size_t c = 0; c < cells; c++)
        if (m_points[c].filled()) ord[c] = (int32_t)n++;
    for (size_t c = 0; c < cells; c++)
        if (!m_points[c].filled()) ord[c] = (int32_t)(n + ghosts++);
}

// upload the adjacency of a loaded map (vga_graph_from_csr)
void PointMap::ensureGraph() {
    if (m_graph || !m_nodes_valid || m_nodes.numNodes() == 0) return;
    const size_t cells = m_cols * m_rows;
    std::vector<int32_t> ord;
    int64_t n, ghosts;
    ordinals(ord, n, ghosts);
    std::vector<int32_t> refs((size_t)(n + ghosts));
    for (size_t c = 0; c < cells; c++) refs[(size_t)ord[c]] = int(PixelRef((int)(c / m_rows), (int)(c % m_rows)));
    FlatRows rows;
    flattenNodes(rows);
    std::vector<uint32_t> col(rows.ref.size());
    for (size_t e = 0; e < col.size(); e++) {
        const PixelRef p(rows.ref[e]);
        if (!includes(p)) throw RuntimeException("graph file: a node run leaves the grid");
        col[e] = (uint32_t)ord[(size_t)p.x * m_rows + (size_t)p.y];
    }
    vga_graph *g = nullptr;
    if (vga_graph_from_csr(context(), n, ghosts, rows.rowptr.data(), col.data(), rows.bin.data(), &g) != VGA_OK)
        throw RuntimeException(std::string("GPU path: ") + vga_last_error());
    vga_graph_set_cell_refs(g, refs.data(), (int64_t)refs.size());
    m_graph = g;  // m_nodes stays valid: it describes this graph
}

// ------------------------------------------------------------------------------------ merge links

bool PointMap::hasMerges() const {
    for (const Point &pt : m_points)
        if (pt.filled() && pt.merged()) return true;
    return false;
}

void PointMap::dropMergedGraph() {
    if (m_merged_graph) vga_graph_free(m_merged_graph);
    m_merged_graph = nullptr;
    m_merged_primary.clear();
}

// pointdata.cpp:1643-1651
bool PointMap::unmergePixel(PixelRef a) {
    const PixelRef c = getPoint(a).merge;
    if (includes(c)) {
        getPoint(c).merge = PixelRef();
        getPoint(c).state &= ~Point::MERGED;
    }
    getPoint(a).merge = PixelRef();
    getPoint(a).state &= ~Point::MERGED;
    dropMergedGraph();
    return true;
}

// pointdata.cpp:1653-1685
bool PointMap::mergePixels(PixelRef a, PixelRef b) {
    if (int(a) == int(b) && getPoint(a).merged()) unmergePixel(a);
    if (int(a) != int(b) && int(getPoint(a).merge) != int(b)) {
        for (const PixelRef &x : {a, b})
            if (getPoint(x).merged()) {
                const PixelRef c = getPoint(x).merge;
                getPoint(c).merge = PixelRef();
                getPoint(c).state &= ~Point::MERGED;
            }
        getPoint(a).merge = b;
        getPoint(a).state |= Point::MERGED;
        getPoint(b).merge = a;
        getPoint(b).state |= Point::MERGED;
    }
    dropMergedGraph();
    return true;
}

void PointMap::contractedRows(Contracted &out) {
    std::vector<int32_t> ord;
    ordinals(ord, out.n, out.ghosts);
    const int64_t n = out.n;
    // the original rows as ordinals
    std::vector<uint64_t> rowptr;
    std::vector<uint32_t> col;
    if (m_nodes_valid) {
        FlatRows rows;
        flattenNodes(rows);
        rowptr.swap(rows.rowptr);
        col.resize(rows.ref.size());
        for (size_t e = 0; e < col.size(); e++) {
            const PixelRef p(rows.ref[e]);
            if (!includes(p)) throw RuntimeException("graph file: a node run leaves the grid");
            col[e] = (uint32_t)ord[(size_t)p.x * m_rows + (size_t)p.y];
        }
    } else if (m_graph) {
        rowptr.resize((size_t)n + 1);
        col.resize((size_t)vga_graph_num_edges(m_graph));
        if (vga_graph_csr(m_graph, rowptr.data(), col.data(), nullptr, nullptr) != VGA_OK)
            throw RuntimeException(std::string("contractedRows: ") + vga_last_error());
    } else {
        throw RuntimeException("contractedRows: the map has no visibility graph");
    }
    // pairing: symmetric and exclusive, both cells filled
    std::vector<int32_t> partner((size_t)n, -1);
    for (size_t c = 0; c < m_cols * m_rows; c++) {
        const Point &pt = m_points[c];
        if (!pt.filled()) continue;
        const int32_t v = ord[c];
        if (!pt.merged()) continue;
        const PixelRef m = pt.merge;
        if (!includes(m) || !getPoint(m).filled() || int(getPoint(m).merge) != int(PixelRef((int)(c / m_rows), (int)(c % m_rows))))
            throw RuntimeException("GPU path: merge links must pair filled cells symmetrically");
        partner[(size_t)v] = ord[(size_t)m.x * m_rows + (size_t)m.y];
    }
    contract_rows(n, rowptr.data(), col.data(), partner.data(), out);
}

void PointMap::radiusCorrection(int radius, LevelTo &level_to, int32_t *total_nodes, int64_t *total_depth, int32_t *dist,
                                int32_t max_levels) {
    if (radius < 1 || radius >= max_levels) return;
    Contracted c;
    contractedRows(c);
    radius_correction(c, radius, level_to, total_nodes, total_depth, dist, max_levels);
}

vga_graph *PointMap::analysisGraph(std::vector<int32_t> *primary) {
    ensureGraph();
    if (!hasMerges()) {
        if (primary) primary->clear();
        return m_graph;
    }
    if (!m_merged_graph) {
        Contracted c;
        contractedRows(c);
        std::vector<int32_t> ord, refs((size_t)(c.n + c.ghosts));
        int64_t n, ghosts;
        ordinals(ord, n, ghosts);
        for (size_t i = 0; i < ord.size(); i++) refs[(size_t)ord[i]] = int(PixelRef((int)(i / m_rows), (int)(i % m_rows)));
        vga_graph *g = nullptr;
        if (vga_graph_from_csr(context(), c.n, c.ghosts, c.rowptr.data(), c.col.data(), nullptr, &g) != VGA_OK)
            throw RuntimeException(std::string("GPU path: ") + vga_last_error());
        vga_graph_set_cell_refs(g, refs.data(), (int64_t)refs.size());
        m_merged_graph = g;
        m_merged_primary = c.primary;
    }
    if (primary) *primary = m_merged_primary;
    return m_merged_graph;
}

// ------------------------------------------------------------------------------------ GraphFile

namespace {
// ShapeMap::read (shapemap.cpp:2273-2383) of one drawing layer: returns the layer's line segments when shown
void read_drawing_layer(ByteReader &in, std::vector<Line> &walls) {
    in.str();                                 // name
    in.skip(4);                               // map type
    const bool show = in.get<uint8_t>() != 0;
    in.skip(1);                               // editable
    in.skip(32 + 4 + 4 + 4 + 4);              // region, rows, cols, next object ref, deprecated int
    const int32_t nshapes = in.get<int32_t>();
    if (nshapes < 0) throw RuntimeException("graph file: negative shape count");
    in.expect((uint64_t)nshapes, 4 + 1 + 40 + 32 + 4);
    std::vector<Line> lines;
    for (int32_t j = 0; j < nshapes; j++) {
        in.skip(4);  // key (std::map order = file order)
        // SalaShape::read (shapemap.cpp:49-64)
        const uint8_t type = in.get<uint8_t>();
        struct {
            double blx, bly, trx, tr_y;
            signed char parity, direction, pad[6];
        } region;
        static_assert(sizeof(region) == 40, "Line image");
        in.raw(&region, 40);
        in.skip(16 + 8 + 8);  // centroid, area, perimeter
        const uint32_t npts = in.get<uint32_t>();
        in.expect(npts, sizeof(Point2f));
        std::vector<Point2f> pts(npts);
        if (npts) in.raw(pts.data(), npts * sizeof(Point2f));
        const bool closed = (type & 0x40) != 0, poly = (type & 0x04) != 0;
        if (type == 0x02) {  // SHAPE_LINE: the region member is the line itself
            Line l;
            l.bl = Point2f(region.blx, region.bly);
            l.tr = Point2f(region.trx, region.tr_y);
            l.parity = region.parity != 0;
            lines.push_back(Line(l.start(), l.end()));
        } else if (poly && npts > 0) {  // polyline / polygon (ShapeMap::getAllShapesAsLines)
            for (size_t n = 0; n + 1 < pts.size(); n++) lines.push_back(Line(pts[n], pts[n + 1]));
            if (closed) lines.push_back(Line(pts.back(), pts.front()));
        }
    }
    const int32_t nobjects = in.get<int32_t>();
    for (int32_t k = 0; k < nobjects; k++) {
        in.skip(4);
        const uint32_t size = in.get<uint32_t>();
        in.skip((size_t)size * 4);
    }
    AttributeTable attributes;
    attributes.read(in);
    in.skip(4);  // displayed attribute
    const int32_t nconnectors = in.get<int32_t>();
    for (int32_t i = 0; i < nconnectors; i++) {  // Connector::read (connector.cpp:28-43)
        const uint32_t nconn = in.get<uint32_t>();
        in.skip((size_t)nconn * 4 + 4);
        for (int m = 0; m < 2; m++) {
            const uint32_t nseg = in.get<uint32_t>();
            in.skip((size_t)nseg * (8 + 4));
        }
    }
    for (int m = 0; m < 2; m++) {  // links, unlinks
        const uint32_t n = in.get<uint32_t>();
        in.skip((size_t)n * 8);
    }
    const char tail = in.get<char>();
    if (tail == 'm') throw RuntimeException("graph file: MapInfo data in a drawing layer is not supported");
    if (show) walls.insert(walls.end(), lines.begin(), lines.end());
}
}  // namespace

int GraphFile::read(const std::string &filename) {
    std::ifstream f(filename.c_str(), std::ios::binary);
    if (!f) {
        m_error = "cannot open " + filename;
        return DISK_ERROR;
    }
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string buf = ss.str();
    return readFromBuffer(buf.data(), buf.size());
}

// MetaGraph::readFromStream (mgraph.cpp:2492-2654)
int GraphFile::readFromBuffer(const char *data, size_t size) {
    m_point_maps.clear();
    m_walls.clear();
    m_head.clear();
    m_tail.clear();
    m_displayed_pointmap = -1;
    m_has_drawing = false;
    try {
        ByteReader in(data, size);
        if (size < 7 || std::memcmp(data, "grf", 3) != 0) return NOT_A_GRAPH;
        in.skip(3);
        const int32_t version = in.get<int32_t>();
        if (version > METAGRAPH_VERSION) return NEWER_VERSION;
        if (version < METAGRAPH_VERSION) {
            m_error = "graph files older than version 440 need the reference's converter";
            return UNSUPPORTED;
        }
        m_state = in.get<int32_t>();
        m_view_class = in.get<int32_t>();
        in.skip(2);  // showgrid, showtext
        auto finish = [&](size_t head_end) {
            m_head.assign(data, head_end);
            m_tail.assign(data + in.pos(), size - in.pos());
            return (int)OK;
        };
        if (in.eof()) return finish(in.pos());
        char type = in.get<char>();
        if (type == 'd' || type == 'v') {
            m_error = "deprecated data-layer / virtual-memory sections need the reference's converter";
            return UNSUPPORTED;
        }
        size_t type_pos = in.pos() - 1;
        bool have_type = true;
        auto next_type = [&]() {
            have_type = !in.eof();
            type_pos = in.pos();
            if (have_type) type = in.get<char>();
        };
        if (type == 'x') {
            for (int i = 0; i < 7; i++) in.str();  // FileProperties::read (fileproperties.h:64-75)
            next_type();
        }
        // The reference renames an unnamed drawing and unnamed drawing files to "<unknown>" when it reads them
        // (mgraph.cpp:2607-2609, spacepixfile.cpp:39-41), so its next write carries that name: same here.
        std::vector<size_t> unnamed;
        auto name_at = [&]() {
            const size_t at = in.pos();
            if (in.str().empty()) unnamed.push_back(at);
        };
        auto take_head = [&](size_t end) {
            m_head.assign(data, end);
            static const char unknown[] = "\x09\x00\x00\x00<unknown>";
            for (auto it = unnamed.rbegin(); it != unnamed.rend(); ++it) m_head.replace(*it, 4, unknown, 13);
        };
        if (have_type && type == 'l') {
            name_at();
            m_region = in.get<Region>();
            const int32_t nfiles = in.get<int32_t>();
            for (int32_t i = 0; i < nfiles; i++) {  // SpacePixelFile::read (spacepixfile.cpp:28-43)
                name_at();
                in.skip(32);
                const int32_t nlayers = in.get<int32_t>();
                for (int32_t k = 0; k < nlayers; k++) read_drawing_layer(in, m_walls);
            }
            m_has_drawing = true;
            next_type();
        }
        if (have_type && type == 'p') {
            m_displayed_pointmap = in.get<int32_t>();
            const int32_t count = in.get<int32_t>();
            if (count < 0 || m_displayed_pointmap < -1 || m_displayed_pointmap >= std::max(count, 1))
                throw RuntimeException("graph file: bad point map count / displayed map");
            for (int32_t i = 0; i < count; i++) {
                m_point_maps.emplace_back(new PointMap(m_region, m_walls));
                m_point_maps.back()->read(in);
            }
            take_head(type_pos);
            m_tail.assign(data + in.pos(), size - in.pos());
            return OK;
        }
        // no point maps: they would be inserted where the next section starts
        take_head(have_type ? type_pos : size);
        m_tail.assign(data + (have_type ? type_pos : size), size - (have_type ? type_pos : size));
        return OK;
    } catch (const std::exception &e) {
        m_error = e.what();
        return DAMAGED_FILE;
    }
}

// MetaGraph::write with currentlayer = false (mgraph.cpp:2656-2763) + writePointMaps (:2833-2845)
int GraphFile::write(const std::string &filename) {
    std::ofstream out(filename.c_str(), std::ios::binary | std::ios::trunc);
    if (!out) {
        m_error = "cannot write " + filename;
        return DISK_ERROR;
    }
    std::string head = m_head;
    std::memcpy(&head[7], &m_state, 4);
    std::memcpy(&head[11], &m_view_class, 4);
    out.write(head.data(), (std::streamsize)head.size());
    if (m_state & POINTMAPS) {
        out.put('p');
        put<int32_t>(out, m_displayed_pointmap);
        put<int32_t>(out, (int32_t)m_point_maps.size());
        for (auto &m : m_point_maps) m->write(out);
    }
    out.write(m_tail.data(), (std::streamsize)m_tail.size());
    out.close();
    return out ? OK : DISK_ERROR;
}

// MetaGraph::addNewPointMap (mgraph.cpp:2799-2819)
int GraphFile::addNewPointMap(const std::string &name) {
    std::string myname = name;
    int counter = 1;
    bool duplicate = true;
    while (duplicate) {
        duplicate = false;
        for (auto &m : m_point_maps)
            if (m->getName() == myname) {
                duplicate = true;
                myname = name + " " + std::to_string(counter++);
                break;
            }
    }
    m_point_maps.emplace_back(new PointMap(m_region, m_walls, myname));
    m_displayed_pointmap = (int)m_point_maps.size() - 1;
    return m_displayed_pointmap;
}

// MetaGraph::setGrid (mgraph.cpp:222-234) incl. setViewClass(SHOWVGATOP) (:167-177)
bool GraphFile::setGrid(double spacing, const Point2f &offset) {
    m_state &= ~POINTMAPS;
    getDisplayedPointMap().setGrid(spacing, offset);
    m_state |= POINTMAPS;
    showVgaTop();
    return true;
}

// MetaGraph::setViewClass(SHOWVGATOP) (mgraph.cpp:101-107, 167-177)
void GraphFile::showVgaTop() {
    if (~m_state & POINTMAPS) return;
    if (m_view_class & VIEWAXIAL)
        m_view_class = VIEWBACKAXIAL | VIEWVGA;
    else if (m_view_class & VIEWDATA)
        m_view_class = VIEWBACKDATA | VIEWVGA;
    else
        m_view_class = VIEWVGA | (m_view_class & (VIEWBACKAXIAL | VIEWBACKDATA));
}

// MetaGraph::makeGraph (mgraph.cpp:264-298)
bool GraphFile::makeGraph(Communicator *comm, bool boundarygraph, double maxdist) {
    m_state |= ANGULARGRAPH;
    bool made = false;
    try {
        made = getDisplayedPointMap().sparkGraph2(comm, boundarygraph, maxdist);
    } catch (const Communicator::CancelledException &) {
        made = false;
    }
    if (made) showVgaTop();
    return made;
}

// the state / view-class part of MetaGraph::makeGraph for graphs finished through PointMap::finishSparkGraph
void GraphFile::graphMade() {
    m_state |= ANGULARGRAPH;
    showVgaTop();
}

}  // namespace dmx
