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
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/vga_b200.h"
#include "geometry.h"

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
    operator int() const { return (int(x) << 16) + (int(y) & 0xffff); }
};

struct Point {
    enum { EMPTY = 0x0001, FILLED = 0x0002, BLOCKED = 0x0004, CONTEXTFILLED = 0x0008, EDGE = 0x0020, MERGED = 0x0040 };
    int state = EMPTY;
    int misc = 0;
    unsigned char grid_connections = 0;
    std::vector<Line> lines;  // wall segments touching the cell, clipped to it
    bool filled() const { return (state & FILLED) != 0; }
    bool blocked() const { return (state & BLOCKED) != 0; }
    bool edge() const { return (state & EDGE) != 0; }
    void set(int s, int undo) {
        state = s | (state & BLOCKED);
        misc = undo;
    }
};

// Column store; a row exists for every cell that had a Node made (x-major order).
class AttributeTable {
  public:
    int insertOrResetColumn(const std::string &name);
    int getColumnIndex(const std::string &name) const;  // -1 if absent
    bool hasColumn(const std::string &name) const { return getColumnIndex(name) >= 0; }
    size_t getNumColumns() const { return m_names.size(); }
    const std::string &getColumnName(size_t i) const { return m_names[i]; }
    size_t getNumRows() const { return m_keys.size(); }
    void setRows(const std::vector<int> &keys);
    void clear();
    const std::vector<int> &keys() const { return m_keys; }
    std::vector<float> &column(int idx) { return m_cols[idx]; }
    const std::vector<float> &column(int idx) const { return m_cols[idx]; }
    void setValue(size_t row, int col, float v) { m_cols[col][row] = v; }
    float getValue(size_t row, int col) const { return m_cols[col][row]; }

  private:
    std::vector<std::string> m_names;
    std::vector<std::vector<float>> m_cols;
    std::vector<int> m_keys;
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
    bool unmake();

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
    const std::string &getName() const { return m_name; }

    // flat image of the hot-path inputs (the vga_grid of the C ABI); arrays owned by the map
    struct Flat {
        std::vector<uint16_t> state;
        std::vector<uint32_t> line_off;
        std::vector<double> lines;
    };
    void flatten(Flat &out) const;
    // adjacency handle on the device (valid after sparkGraph2 / adoptGraph)
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
};

class VGAVisualLocal : public IVGA {
    bool m_gates_only;

  public:
    std::string getAnalysisName() const override { return "Local Visibility Analysis"; }
    bool run(Communicator *comm, PointMap &map, bool simple_version) override;
    explicit VGAVisualLocal(bool gates_only) : m_gates_only(gates_only) {}
};

// process-wide GPU context (one process drives one GPU; device from VGA_DEVICE or LOCAL_RANK)
vga_ctx *shared_context();
void release_shared_context();

}  // namespace dmx
