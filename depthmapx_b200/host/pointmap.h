// GPU-backed mirror of the reference's operator interface for the visibility-graph hot path.
//
//   dmx::PointMap            <->  PointMap            (salalib/pointdata.h:76-350)
//     setGrid / blockLines / makePoints : host pre-steps, same semantics as pointdata.cpp:122-171,
//                                         296-357, 402-514 (stays on host)
//     sparkGraph2(comm, boundarygraph, maxdist) : pointdata.cpp:1246-1341, construction on the GPU
//                                         through the C ABI (include/vga_b200.h)
//   dmx::VGAVisualGlobal     <->  VGAVisualGlobal(radius, gates_only)::run  (vgavisualglobal.cpp:23)
//   dmx::VGAVisualLocal      <->  VGAVisualLocal(gates_only)::run           (vgavisuallocal.cpp:23)
//   dmx::Communicator        <->  Communicator (genlib/comm.h): progress + cooperative cancel
//   dmx::AttributeTable      <->  the float32 row store keyed by int(PixelRef)
//                                 (salalib/attributetable.h:174-235), rows in x-major order
//
// Same names, argument meaning, column names, -1 sentinels and error behaviour (return false /
// CancelledException).  There is no CPU compute path: without a CUDA device sparkGraph2 and the
// run() methods throw dmx::RuntimeException.
#pragma once

#include <cstdint>
#include <iosfwd>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vga_b200.h"
#include "geometry.h"
#include "merge_contract.h"

namespace dmx {

class RuntimeException : public std::runtime_error {
  public:
    explicit RuntimeException(const std::string &m) : std::runtime_error(m) {}
};

class Communicator {
  public:
    class CancelledException {};
    enum { NUM_STEPS, CURRENT_STEP, NUM_RECORDS, CURRENT_RECORD };
    virtual ~Communicator() {}
    virtual void CommPostMessage(int /*m*/, int /*x*/) const {}
    virtual bool IsCancelled() const { return false; }
};

struct PixelRef {
    short x = -1, y = -1;
    PixelRef() = default;
    PixelRef(int ax, int ay) : x((short)ax), y((short)ay) {}
    explicit PixelRef(int packed) : x((short)(packed >> 16)), y((short)(packed & 0xffff)) {}
    // (int(x) << 16) + (int(y) & 0xffff) of the reference (pixelref.h:81-82), written without shifting a negative value
    operator int() const { return (int)(((unsigned)(unsigned short)x << 16) | (unsigned)(unsigned short)y); }
};

struct Point {
    enum { EMPTY = 0x0001, FILLED = 0x0002, BLOCKED = 0x0004, CONTEXTFILLED = 0x0008, SELECTED = 0x0010, EDGE = 0x0020, MERGED = 0x0040 };
    int state = EMPTY;
    int misc = 0;
    int block = 0;            // Point::m_block: unused by the reference but serialised (point.cpp:28-31)
    unsigned char grid_connections = 0;
    PixelRef merge;           // Point::m_merge, (-1,-1) = not merged
    Point2f location;         // Point::m_location = centre of the cell (pointdata.cpp:159)
    std::vector<Line> lines;  // wall segments touching the cell, clipped to it
    bool filled() const { return (state & FILLED) != 0; }
    bool blocked() const { return (state & BLOCKED) != 0; }
    bool edge() const { return (state & EDGE) != 0; }
    bool contextfilled() const { return (state & CONTEXTFILLED) != 0; }
    bool merged() const { return !(merge.x == -1 && merge.y == -1); }
    void set(int s, int undo) {
        state = s | (state & BLOCKED);
        misc = undo;
    }
};

// Display parameters stored with every column and with the table (salalib/displayparams.h:3-11).
struct DisplayParams {
    float blue = 0.0f, red = 1.0f;
    int colorscale = 0;
};

// Column store; a row exists for every cell that had a Node made (x-major order = AttributeKey order).
// Keeps what the reference serialises per column (salalib/attributetable.cpp:91-124): running statistics
// (min / max / total, updated by every setValue exactly as AttributeColumnImpl::updateStats :65-84 does),
// hidden / locked flags, display parameters and formula; per row the layer key.
class AttributeTable {
  public:
    struct Column {
        std::string name;
        double min = -1.0, max = -1.0, total = -1.0;
        bool hidden = false, locked = false;
        DisplayParams display;
        std::string formula;
    };
    AttributeTable();
    int insertOrResetColumn(const std::string &name);        // attributetable.cpp:303-319
    int insertOrResetLockedColumn(const std::string &name);  // :321-326
    int getOrInsertColumn(const std::string &name);          // :290-301: an existing column keeps its values and statistics
    int getColumnIndex(const std::string &name) const;       // -1 if absent
    bool hasColumn(const std::string &name) const { return getColumnIndex(name) >= 0; }
    size_t getNumColumns() const { return m_columns.size(); }
    const std::string &getColumnName(size_t i) const { return m_columns[i].name; }
    const Column &getColumn(size_t i) const { return m_columns[i]; }
    int getColumnSortedIndex(int idx) const;                 // :479-485
    size_t getNumRows() const { return m_keys.size(); }
    void setRows(const std::vector<int> &keys);  // addRow for every key (ascending), values -1, layer key 1
    void clear();
    const std::vector<int> &keys() const { return m_keys; }
    const std::vector<float> &column(int idx) const { return m_cols[idx]; }
    void setValue(size_t row, int col, float v);  // AttributeRowImpl::setValue :153-166 (updates the statistics)
    float getValue(size_t row, int col) const { return m_cols[col][row]; }
    // AttributeTable::read / write (attributetable.cpp:397-457) incl. the layer manager block
    // (layermanagerimpl.cpp:90-149, kept as opaque bytes)
    bool read(class ByteReader &in);
    void write(std::ostream &out) const;

  private:
    std::vector<Column> m_columns;
    std::vector<std::vector<float>> m_cols;
    std::vector<int> m_keys;
    std::vector<int64_t> m_layer_keys;
    DisplayParams m_display;
    std::string m_layers;  // serialised LayerManagerImpl
};

// Bounds-checked cursor over a file image (all .graph integers are little-endian, structs are raw x86-64 images).
class ByteReader {
  public:
    ByteReader(const char *data, size_t size) : m_data(data), m_size(size) {}
    size_t pos() const { return m_pos; }
    size_t size() const { return m_size; }
    bool eof() const { return m_pos >= m_size; }
    const char *at(size_t p) const { return m_data + p; }
    void raw(void *dst, size_t n);
    void skip(size_t n);
    template <typename T> T get() {
        T v;
        raw(&v, sizeof(T));
        return v;
    }
    std::string str();  // dXstring::readString: u32 length + bytes
    // a count read from the file must be coverable by the bytes that are left (guards allocations against damaged files)
    void expect(uint64_t count, uint64_t min_bytes_each) const;

  private:
    const char *m_data;
    size_t m_size, m_pos = 0;
};

// Run-length encoded adjacency of every cell that has a Node (salalib/ngraph.h:31-149): 32 bins per node, each a
// direction class, the stored node count, the far distance, the occlusion distance and its runs (PixelVec
// start..end).  Filled by PointMap::read, or from the device rows after sparkGraph2 when the map is written.
struct NodeStore {
    struct Run {
        PixelRef start, end;
    };
    struct Bin {
        signed char dir = 0;  // PixelRef::NODIR / HORIZONTAL 1 / VERTICAL 2 / POSDIAGONAL 4 / NEGDIAGONAL 8
        uint16_t count = 0;
        float distance = 0.0f, occ_distance = 0.0f;
        uint32_t first_run = 0, nruns = 0;
    };
    std::vector<int32_t> node_of_cell;  // [cells] node index or -1
    std::vector<Bin> bins;              // [nodes*32]
    std::vector<Run> runs;
    std::vector<uint32_t> occl_off;     // [nodes*32+1] into occl (Node::m_occlusion_bins)
    std::vector<PixelRef> occl;
    size_t numNodes() const { return bins.size() / 32; }
    void clear();
};

class PointMap {
  public:
    // parentRegion = bounding box of the drawing; walls = its line shapes in drawing order
    PointMap(const Region &parentRegion, const std::vector<Line> &walls, const std::string &name = "VGA Map");
    ~PointMap();
    PointMap(const PointMap &) = delete;
    PointMap &operator=(const PointMap &) = delete;

    bool setGrid(double spacing, const Point2f &offset = Point2f());
    bool blockLines();
    void unblockLines(bool clearblockedflag = true);
    bool makePoints(const Point2f &seed, int fill_type, Communicator *comm = nullptr);
    bool sparkGraph2(Communicator *comm, bool boundarygraph, double maxdist);
    bool unmake(bool removeLinks = false);  // pointdata.cpp:1343-1374

    size_t getCols() const { return m_cols; }
    size_t getRows() const { return m_rows; }
    double getSpacing() const { return m_spacing; }
    const Region &getRegion() const { return m_region; }
    const Point2f &getBottomLeft() const { return m_bottom_left; }
    int getFilledPointCount() const { return m_filled_point_count; }
    bool isProcessed() const { return m_processed; }
    bool isBoundaryGraph() const { return m_boundarygraph; }
    bool includes(const PixelRef &p) const { return p.x >= 0 && p.x < (int)m_cols && p.y >= 0 && p.y < (int)m_rows; }
    Point &getPoint(const PixelRef &p) { return m_points[(size_t)p.x * m_rows + (size_t)p.y]; }
    const Point &getPoint(const PixelRef &p) const { return m_points[(size_t)p.x * m_rows + (size_t)p.y]; }
    PixelRef pixelate(const Point2f &p, bool constrain = true) const;
    Point2f depixelate(const PixelRef &p) const {
        return Point2f(m_bottom_left.x + m_spacing * 1.0 * double(p.x), m_bottom_left.y + m_spacing * 1.0 * double(p.y));
    }
    Region regionate(const PixelRef &p, double border) const;
    AttributeTable &getAttributeTable() { return m_attributes; }
    const AttributeTable &getAttributeTable() const { return m_attributes; }
    int getDisplayedAttribute() const { return m_displayed_attribute; }
    void setDisplayedAttribute(int col) { m_displayed_attribute = col; }
    void overrideDisplayedAttribute(int col) { m_displayed_attribute = col; }  // pointdata.h: sets the member, nothing else
    const std::string &getName() const { return m_name; }

    // flat image of the hot-path inputs (the vga_grid of the C ABI); arrays owned by the map
    struct Flat {
        std::vector<uint16_t> state;
        std::vector<uint32_t> line_off;
        std::vector<double> lines;
    };
    void flatten(Flat &out) const;
    // The PointMap section of a .graph file: the bytes PointMap::read / write handle
    // (salalib/pointdata.cpp:1073-1188; Point::read / write salalib/point.cpp:25-73; Node / Bin / PixelVec
    // run-length codec salalib/ngraph.cpp:195-220, 420-583; attribute table salalib/attributetable.cpp:91-124,
    // 397-457).  read() restores grid geometry, cell states, grid connections, merge links, every attribute
    // column and the run-length adjacency, so `-m VGA` on a loaded map needs no reference object; write()
    // emits exactly the bytes the reference would write for the same map.
    bool read(ByteReader &in);
    void write(std::ostream &out);
    // Run-length adjacency (the reference's Nodes).  After sparkGraph2 it is produced on demand from the device
    // rows (the GPU graph is the master copy); after read() it is what the file held.
    const NodeStore &nodes();
    // Flat adjacency of the node store in Node::first/next order: rowptr [N+1] over filled cells in x-major
    // order, packed PixelRef and bin id of every iterated pixel.
    struct FlatRows {
        std::vector<uint64_t> rowptr;
        std::vector<int32_t> ref;
        std::vector<uint8_t> bin;
    };
    void flattenNodes(FlatRows &out);
    // Rebuild the node store from flat rows (any order inside a row; `accepted` = 0 marks the fill-in pixels of
    // a diagonal first..last run, NULL = all accepted): the encoder used after a GPU build, Node::make /
    // Bin::make (ngraph.cpp:27-58, 234-304).  far_bin_dists [N*32].
    void encodeNodes(const uint64_t *rowptr, const int32_t *ref, const uint8_t *bin, const uint8_t *accepted,
                     const float *far_bin_dists);
    // make sure graph() is valid: uploads the loaded adjacency when there is one (needs the GPU)
    void ensureGraph();
    // Merge links (`-m LINK`; pointdata.cpp:1643-1685): a symmetric, exclusive pairing of filled cells.
    bool mergePixels(PixelRef a, PixelRef b);
    bool unmergePixel(PixelRef a);
    bool isPixelMerged(const PixelRef &a) const { return getPoint(a).merged(); }
    bool hasMerges() const;
    // Point::blocked || PointMap::blockedAdjacent (pointdata.cpp:1016-1066) per attribute row: the cells VGAMetric /
    // VGAAngular expand; Point::m_merge per row as row indices (-1 = not merged), empty when nothing is merged
    std::vector<uint8_t> blockedAdjacentFlags() const;
    std::vector<int32_t> mergePartners() const;
    // The adjacency the BFS analyses (global, step depth) run on when cells are merged: every merged pair contracted
    // into its smaller-ordinal cell, see merge_contract.h.
    typedef dmx::Contracted Contracted;
    typedef dmx::LevelTo LevelTo;
    void contractedRows(Contracted &out);
    // graph() when nothing is merged, otherwise the contracted adjacency on the device (built on demand)
    vga_graph *analysisGraph(std::vector<int32_t> *primary = nullptr);
    // radius-limited global analysis of a merged map: the at-the-radius second counts (merge_contract.h)
    void radiusCorrection(int radius, LevelTo &level_to, int32_t *total_nodes, int64_t *total_depth, int32_t *dist,
                          int32_t max_levels);
    // x-major packed PixelRefs of the filled cells (= attribute row keys once the graph is made)
    std::vector<int> filledKeys() const;
    // per filled cell (x-major): 1 for context-filled cells (semi-fill, fill_type 1) that are not "even"
    // (PixelRef::iseven, pixelref.h:68-69): the analyses skip them as sources and, under a radius limit, count but do not
    // expand them (vgavisualglobal.cpp:75, 108-110; vgavisuallocal.cpp:43; vgavisualglobaldepth.cpp:53).  Empty when the
    // map has no such cell.
    std::vector<uint8_t> contextSkipFlags() const;
    // Second half of sparkGraph2 (pointdata.cpp:1268-1341 minus the per-source work): attribute rows, the three
    // columns, grid connections and flags from per-source results that libvga_b200 produced -- on this GPU
    // (sparkGraph2 calls it) or on other ranks of a multi-GPU build (gathered by the caller).
    void finishSparkGraph(bool boundarygraph, const int32_t *connectivity, const double *sum_d, const double *sum_d2,
                          const uint8_t *grid_connections);
    // first half: blockLines + boundary un-fill (pointdata.cpp:1250-1264)
    void beginSparkGraph(bool boundarygraph);
    // selection (pointdata.cpp:939-987, 905-937): filled cells inside r get Point::SELECTED and join the set
    bool setCurSel(const Region &r, bool add = false);
    bool clearSel();
    const std::vector<PixelRef> &getSelSet() const { return m_selection_set; }  // x-major (std::set<PixelRef> order)
    // adjacency handle on the device (valid after sparkGraph2 / adoptGraph / ensureGraph)
    vga_graph *graph() const { return m_graph; }
    vga_ctx *context();
    // attach an adjacency produced elsewhere (e.g. flattened from a loaded .graph)
    void adoptGraph(vga_graph *g);
    // per-cell connection lists in the reference's iteration order are not kept on the host;
    // the sorted rows can be fetched with vga_graph_csr(graph(), ...).

  private:
    std::vector<PixelRef> pixelateLineTouching(Line l, double tolerance) const;
    void blockLine(const Line &li);
    int expand(const PixelRef p1, const PixelRef p2, std::vector<PixelRef> &list, int filltype);

    std::string m_name;
    Region m_parent;
    std::vector<Line> m_walls;
    double m_spacing = 0.0;
    Point2f m_offset, m_bottom_left;
    Region m_region;
    size_t m_cols = 0, m_rows = 0;
    std::vector<Point> m_points;
    int m_filled_point_count = 0;
    int m_undocounter = 0;
    bool m_initialised = false, m_blockedlines = false, m_processed = false, m_boundarygraph = false;
    int m_displayed_attribute = -2;
    AttributeTable m_attributes;
    vga_graph *m_graph = nullptr;
    vga_graph *m_merged_graph = nullptr;  // contracted adjacency, valid while the merges do not change
    std::vector<int32_t> m_merged_primary;
    void dropMergedGraph();
    void ordinals(std::vector<int32_t> &ord, int64_t &n, int64_t &ghosts) const;
    std::vector<PixelRef> m_selection_set;
    bool m_has_selection = false;
    NodeStore m_nodes;
    bool m_nodes_valid = false;  // m_nodes describes the current graph
};

class IVGA {
  public:
    virtual std::string getAnalysisName() const = 0;
    virtual bool run(Communicator *comm, PointMap &map, bool simple_version) = 0;
    virtual ~IVGA() {}
};

class VGAVisualGlobal : public IVGA {
    double m_radius;
    bool m_gates_only;

  public:
    std::string getAnalysisName() const override { return "Global Visibility Analysis"; }
    bool run(Communicator *comm, PointMap &map, bool simple_version) override;
    VGAVisualGlobal(double radius, bool gates_only) : m_radius(radius), m_gates_only(gates_only) {}
    // column set-up + formula stage + row writes (vgavisualglobal.cpp:33-63, 131-193, 214) from the BFS integers
    // of all N cells in x-major order -- computed by vga_global on this GPU (run() calls it) or gathered from the
    // ranks of a multi-GPU run
    // skip: optional per-row flags of the sources the reference does not analyse (PointMap::contextSkipFlags)
    static void writeAttributes(PointMap &map, double radius, bool simple_version, const int32_t *total_nodes,
                                const int64_t *total_depth, const int32_t *dist, int32_t max_levels,
                                const uint8_t *skip = nullptr);
};

class VGAVisualLocal : public IVGA {
    bool m_gates_only;

  public:
    std::string getAnalysisName() const override { return "Local Visibility Analysis"; }
    bool run(Communicator *comm, PointMap &map, bool simple_version) override;
    explicit VGAVisualLocal(bool gates_only) : m_gates_only(gates_only) {}
    // vgavisuallocal.cpp:31-35, 84-96, 112 from the integers of vga_local
    static void writeAttributes(PointMap &map, bool simple_version, const int64_t *cluster, const int32_t *k,
                                const int32_t *total, const float *control, const uint8_t *skip = nullptr);
};

// Visual step depth from the map's selection (salalib/vgamodules/vgavisualglobaldepth.cpp:23-75)
class VGAVisualGlobalDepth : public IVGA {
  public:
    std::string getAnalysisName() const override { return "Global Visibility Depth"; }
    bool run(Communicator *comm, PointMap &map, bool simple_version) override;
    // column "Visual Step Depth" from the per-cell depths of vga_step_depth (-1 = not reached: value stays -1)
    static void writeAttributes(PointMap &map, const int32_t *depth);
};

// Metric VGA (salalib/vgamodules/vgametric.cpp:25-136): per-source shortest metric paths; radius in map units
class VGAMetric : public IVGA {
    double m_radius;
    bool m_gates_only;

  public:
    std::string getAnalysisName() const override { return "Metric Analysis"; }
    bool run(Communicator *comm, PointMap &map, bool simple_version) override;
    VGAMetric(double radius, bool gates_only) : m_radius(radius), m_gates_only(gates_only) {}
    // columns (vgametric.cpp:32-56), row writes (:116-120) and display (:131-133) from the four columns of vga_metric; NULL
    // arrays = columns only (gates_only)
    static void writeAttributes(PointMap &map, double radius, const float *mean_angle, const float *mean_path_dist,
                                const float *mean_line_dist, const float *node_count);
};

// Angular VGA (salalib/vgamodules/vgaangular.cpp:22-133): least-turn paths; radius in units of 90 degrees
class VGAAngular : public IVGA {
    double m_radius;
    bool m_gates_only;

  public:
    std::string getAnalysisName() const override { return "Angular Analysis"; }
    bool run(Communicator *comm, PointMap &map, bool simple_version) override;
    VGAAngular(double radius, bool gates_only) : m_radius(radius), m_gates_only(gates_only) {}
    static void writeAttributes(PointMap &map, double radius, const float *mean_depth, const float *total_depth,
                                const float *node_count);
};

// process-wide GPU context (one process drives one GPU; device from VGA_DEVICE or LOCAL_RANK)
vga_ctx *shared_context();
void release_shared_context();

}  // namespace dmx
