// Flat C view of the host layer for ctypes (include/vga_host.h).
#include "../../include/vga_host.h"

#include <cstring>
#include <string>

#include "graphfile.h"
#include "pointmap.h"

static thread_local std::string g_err;

template <typename F> static int guarded(F f) {
    try {
        return f() ? 1 : 0;
    } catch (const dmx::Communicator::CancelledException &) {
        g_err = "cancelled";
        return -2;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

extern "C" {

const char *dmxh_last_error(void) { return g_err.c_str(); }

void *dmxh_map_create(const double *walls, int nwalls, double spacing) {
    std::vector<dmx::Line> ls;
    dmx::Region r;
    for (int i = 0; i < nwalls; i++) {
        dmx::Line l(dmx::Point2f(walls[4 * i], walls[4 * i + 1]), dmx::Point2f(walls[4 * i + 2], walls[4 * i + 3]));
        if (i == 0) {
            r = dmx::Region(l.bl, l.tr);
        } else {
            if (l.bl.x < r.bl.x) r.bl.x = l.bl.x;
            if (l.bl.y < r.bl.y) r.bl.y = l.bl.y;
            if (l.tr.x > r.tr.x) r.tr.x = l.tr.x;
            if (l.tr.y > r.tr.y) r.tr.y = l.tr.y;
        }
        ls.push_back(l);
    }
    dmx::PointMap *m = new dmx::PointMap(r, ls);
    m->setGrid(spacing);
    return m;
}

void dmxh_map_destroy(void *map) { delete static_cast<dmx::PointMap *>(map); }

void dmxh_map_grid(void *map, int32_t *cols, int32_t *rows, double *spacing, double *bl_x, double *bl_y) {
    dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
    *cols = (int32_t)m->getCols();
    *rows = (int32_t)m->getRows();
    *spacing = m->getSpacing();
    *bl_x = m->getBottomLeft().x;
    *bl_y = m->getBottomLeft().y;
}

int dmxh_map_block_lines(void *map) {
    return guarded([&] { return static_cast<dmx::PointMap *>(map)->blockLines(); });
}

int dmxh_map_fill(void *map, double x, double y) {
    return guarded([&] { return static_cast<dmx::PointMap *>(map)->makePoints(dmx::Point2f(x, y), 0); });
}

int dmxh_map_fill_type(void *map, double x, double y, int fill_type) {
    return guarded([&] { return static_cast<dmx::PointMap *>(map)->makePoints(dmx::Point2f(x, y), fill_type); });
}

int dmxh_map_context_skip(void *map, uint8_t *flags) {
    const std::vector<uint8_t> f = static_cast<dmx::PointMap *>(map)->contextSkipFlags();
    if (flags && !f.empty()) std::memcpy(flags, f.data(), f.size());
    return f.empty() ? 0 : 1;
}

int dmxh_map_filled_count(void *map) { return static_cast<dmx::PointMap *>(map)->getFilledPointCount(); }

void dmxh_map_flat(void *map, int64_t *cells, int64_t *nseg, uint16_t *state, uint32_t *line_off, double *lines) {
    dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
    dmx::PointMap::Flat f;
    m->flatten(f);
    if (cells) *cells = (int64_t)f.state.size();
    if (nseg) *nseg = (int64_t)f.lines.size() / 5;
    if (state) std::memcpy(state, f.state.data(), f.state.size() * sizeof(uint16_t));
    if (line_off) std::memcpy(line_off, f.line_off.data(), f.line_off.size() * sizeof(uint32_t));
    if (lines && !f.lines.empty()) std::memcpy(lines, f.lines.data(), f.lines.size() * sizeof(double));
}

int dmxh_map_make_graph(void *map, int boundarygraph, double maxdist) {
    return guarded([&] { return static_cast<dmx::PointMap *>(map)->sparkGraph2(nullptr, boundarygraph != 0, maxdist); });
}

int dmxh_map_vga_global(void *map, double radius, int simple_version) {
    return guarded([&] {
        return dmx::VGAVisualGlobal(radius, false).run(nullptr, *static_cast<dmx::PointMap *>(map), simple_version != 0);
    });
}

int dmxh_map_vga_local(void *map, int simple_version) {
    return guarded([&] {
        return dmx::VGAVisualLocal(false).run(nullptr, *static_cast<dmx::PointMap *>(map), simple_version != 0);
    });
}

int dmxh_map_vga_metric(void *map, double radius) {
    return guarded([&] { return dmx::VGAMetric(radius, false).run(nullptr, *static_cast<dmx::PointMap *>(map), false); });
}

int dmxh_map_vga_angular(void *map, double radius) {
    return guarded([&] { return dmx::VGAAngular(radius, false).run(nullptr, *static_cast<dmx::PointMap *>(map), false); });
}

int dmxh_map_columns(void *map, char *buf, int buflen) {
    const dmx::AttributeTable &t = static_cast<dmx::PointMap *>(map)->getAttributeTable();
    std::string s;
    for (size_t i = 0; i < t.getNumColumns(); i++) s += t.getColumnName(i) + "\n";
    std::strncpy(buf, s.c_str(), (size_t)buflen - 1);
    buf[buflen - 1] = 0;
    return (int)t.getNumColumns();
}

int64_t dmxh_map_num_rows(void *map) { return (int64_t)static_cast<dmx::PointMap *>(map)->getAttributeTable().getNumRows(); }

int dmxh_map_attr(void *map, const char *name, float *out) {
    const dmx::AttributeTable &t = static_cast<dmx::PointMap *>(map)->getAttributeTable();
    int c = t.getColumnIndex(name);
    if (c < 0) return 0;
    std::memcpy(out, t.column(c).data(), t.getNumRows() * sizeof(float));  // dmxh_map_num_rows floats
    return 1;
}

int dmxh_map_grid_connections(void *map, uint8_t *out) {
    dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
    size_t v = 0;
    for (size_t x = 0; x < m->getCols(); x++)
        for (size_t y = 0; y < m->getRows(); y++) {
            const dmx::Point &p = m->getPoint(dmx::PixelRef((int)x, (int)y));
            if (p.filled()) out[v++] = p.grid_connections;
        }
    return 1;
}

int dmxh_map_step_depth(void *map, const double *points, int npoints) {
    return guarded([&] {
        dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
        for (int i = 0; i < npoints; i++) {
            dmx::Point2f p(points[2 * i], points[2 * i + 1]);
            m->setCurSel(dmx::Region(p, p), true);
        }
        return dmx::VGAVisualGlobalDepth().run(nullptr, *m, false);
    });
}

int dmxh_map_select(void *map, const double *points, int npoints) {
    return guarded([&] {
        dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
        for (int i = 0; i < npoints; i++) {
            dmx::Point2f p(points[2 * i], points[2 * i + 1]);
            m->setCurSel(dmx::Region(p, p), true);
        }
        return true;
    });
}

int64_t dmxh_map_selection(void *map, int32_t *refs) {
    const auto &sel = static_cast<dmx::PointMap *>(map)->getSelSet();
    if (refs)
        for (size_t i = 0; i < sel.size(); i++) refs[i] = int(sel[i]);
    return (int64_t)sel.size();
}

int dmxh_map_flat_rows(void *map, int64_t *n, int64_t *entries, uint64_t *rowptr, int32_t *ref, uint8_t *bin) {
    return guarded([&] {
        dmx::PointMap::FlatRows rows;
        static_cast<dmx::PointMap *>(map)->flattenNodes(rows);
        if (n) *n = (int64_t)rows.rowptr.size() - 1;
        if (entries) *entries = (int64_t)rows.ref.size();
        if (rowptr) std::memcpy(rowptr, rows.rowptr.data(), rows.rowptr.size() * sizeof(uint64_t));
        if (ref && !rows.ref.empty()) std::memcpy(ref, rows.ref.data(), rows.ref.size() * sizeof(int32_t));
        if (bin && !rows.bin.empty()) std::memcpy(bin, rows.bin.data(), rows.bin.size());
        return true;
    });
}

int dmxh_map_bins(void *map, uint16_t *bin_count, float *bin_dist) {
    return guarded([&] {
        dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
        const dmx::NodeStore &ns = m->nodes();
        size_t v = 0;
        for (size_t c = 0; c < m->getCols() * m->getRows(); c++) {
            if (!m->getPoint(dmx::PixelRef((int)(c / m->getRows()), (int)(c % m->getRows()))).filled()) continue;
            const int32_t node = ns.node_of_cell.empty() ? -1 : ns.node_of_cell[c];
            for (int i = 0; i < 32; i++) {
                bin_count[v * 32 + (size_t)i] = node >= 0 ? ns.bins[(size_t)node * 32 + (size_t)i].count : 0;
                bin_dist[v * 32 + (size_t)i] = node >= 0 ? ns.bins[(size_t)node * 32 + (size_t)i].distance : 0.0f;
            }
            v++;
        }
        return true;
    });
}

int dmxh_map_encode_nodes(void *map, const uint64_t *rowptr, const int32_t *ref, const uint8_t *bin,
                          const uint8_t *accepted, const float *far_bin_dists) {
    return guarded([&] {
        static_cast<dmx::PointMap *>(map)->encodeNodes(rowptr, ref, bin, accepted, far_bin_dists);
        return true;
    });
}

int dmxh_map_finish_graph(void *map, int boundarygraph, const int32_t *connectivity, const double *sum_d,
                          const double *sum_d2, const uint8_t *grid_connections) {
    return guarded([&] {
        static_cast<dmx::PointMap *>(map)->finishSparkGraph(boundarygraph != 0, connectivity, sum_d, sum_d2, grid_connections);
        return true;
    });
}

int dmxh_map_begin_graph(void *map, int boundarygraph) {
    return guarded([&] {
        static_cast<dmx::PointMap *>(map)->beginSparkGraph(boundarygraph != 0);
        return true;
    });
}

int dmxh_map_write_global(void *map, double radius, int simple_version, const int32_t *total_nodes,
                          const int64_t *total_depth, const int32_t *dist, int32_t max_levels) {
    return guarded([&] {
        dmx::PointMap &m = *static_cast<dmx::PointMap *>(map);
        const std::vector<uint8_t> skip = m.contextSkipFlags();
        dmx::VGAVisualGlobal::writeAttributes(m, radius, simple_version != 0, total_nodes, total_depth, dist, max_levels,
                                              skip.empty() ? nullptr : skip.data());
        return true;
    });
}

int dmxh_map_write_local(void *map, int simple_version, const int64_t *cluster, const int32_t *k, const int32_t *total,
                         const float *control) {
    return guarded([&] {
        dmx::PointMap &m = *static_cast<dmx::PointMap *>(map);
        const std::vector<uint8_t> skip = m.contextSkipFlags();
        dmx::VGAVisualLocal::writeAttributes(m, simple_version != 0, cluster, k, total, control, skip.empty() ? nullptr : skip.data());
        return true;
    });
}

int dmxh_map_write_step_depth(void *map, const int32_t *depth) {
    return guarded([&] {
        dmx::VGAVisualGlobalDepth::writeAttributes(*static_cast<dmx::PointMap *>(map), depth);
        return true;
    });
}

int dmxh_map_merge(void *map, double ax, double ay, double bx, double by) {
    return guarded([&] {
        dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
        // linkutils::pixelateMergeLines + mergePixelPairs (salalib/linkutils.cpp:20-98)
        const dmx::PixelRef a = m->pixelate(dmx::Point2f(ax, ay), false), b = m->pixelate(dmx::Point2f(bx, by), false);
        if (!m->includes(a) || !m->getPoint(a).filled() || !m->includes(b) || !m->getPoint(b).filled())
            throw dmx::RuntimeException("Line ends not both on painted analysis space");
        if (m->isPixelMerged(a) || m->isPixelMerged(b))
            throw dmx::RuntimeException("Link pixel found that is already linked on the map");
        return m->mergePixels(a, b);
    });
}

int dmxh_map_contracted_rows(void *map, int64_t *n, int64_t *entries, uint64_t *rowptr, uint32_t *col, int32_t *primary) {
    return guarded([&] {
        dmx::PointMap::Contracted c;
        static_cast<dmx::PointMap *>(map)->contractedRows(c);
        if (n) *n = c.n;
        if (entries) *entries = (int64_t)c.col.size();
        if (rowptr) std::memcpy(rowptr, c.rowptr.data(), c.rowptr.size() * sizeof(uint64_t));
        if (col && !c.col.empty()) std::memcpy(col, c.col.data(), c.col.size() * sizeof(uint32_t));
        if (primary && !c.primary.empty()) std::memcpy(primary, c.primary.data(), c.primary.size() * sizeof(int32_t));
        return true;
    });
}

namespace {
struct CallbackLevelTo : dmx::PointMap::LevelTo {
    dmxh_level_prepare_fn prep;
    dmxh_level_run_fn go;
    void *user;
    int64_t n = 0;
    CallbackLevelTo(dmxh_level_prepare_fn p, dmxh_level_run_fn r, void *u) : prep(p), go(r), user(u) {}
    void prepare(int64_t cells, const std::vector<uint64_t> &t_rowptr, const std::vector<uint32_t> &t_col) override {
        n = cells;
        prep(user, cells, t_rowptr.data(), t_col.data());
    }
    void run(const std::vector<int64_t> &seeds, std::vector<int32_t> &level) override {
        level.assign((size_t)n, -1);
        go(user, seeds.data(), (int64_t)seeds.size(), level.data());
    }
};
}  // namespace

int dmxh_map_radius_correction(void *map, int radius, dmxh_level_prepare_fn prepare, dmxh_level_run_fn run, void *user,
                               int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels) {
    return guarded([&] {
        CallbackLevelTo cb(prepare, run, user);
        static_cast<dmx::PointMap *>(map)->radiusCorrection(radius, cb, total_nodes, total_depth, dist, max_levels);
        return true;
    });
}

int dmxh_map_state(void *map, uint16_t *state) {
    dmx::PointMap *m = static_cast<dmx::PointMap *>(map);
    const size_t rows = m->getRows();
    for (size_t c = 0; c < m->getCols() * rows; c++)
        state[c] = (uint16_t)m->getPoint(dmx::PixelRef((int)(c / rows), (int)(c % rows))).state;
    return 1;
}

void *dmxh_graph_open(const char *path) {
    dmx::GraphFile *f = new dmx::GraphFile();
    int rc = f->read(path);
    if (rc != dmx::GraphFile::OK) {
        g_err = "GraphFile::read: code " + std::to_string(rc) + " " + f->lastError();
        delete f;
        return nullptr;
    }
    return f;
}

void dmxh_graph_close(void *file) { delete static_cast<dmx::GraphFile *>(file); }

int dmxh_graph_save(void *file, const char *path) {
    return guarded([&] {
        dmx::GraphFile *f = static_cast<dmx::GraphFile *>(file);
        int rc = f->write(path);
        if (rc != dmx::GraphFile::OK) throw dmx::RuntimeException("GraphFile::write: " + f->lastError());
        return true;
    });
}

int dmxh_graph_num_maps(void *file) { return (int)static_cast<dmx::GraphFile *>(file)->getNumPointMaps(); }
int dmxh_graph_displayed_map(void *file) { return static_cast<dmx::GraphFile *>(file)->getDisplayedPointMapRef(); }
void *dmxh_graph_map(void *file, int i) { return &static_cast<dmx::GraphFile *>(file)->getPointMap((size_t)i); }

int64_t dmxh_graph_walls(void *file, double *out) {
    const auto &w = static_cast<dmx::GraphFile *>(file)->getVisibleDrawingLines();
    if (out)
        for (size_t i = 0; i < w.size(); i++) {
            out[4 * i] = w[i].start().x;
            out[4 * i + 1] = w[i].start().y;
            out[4 * i + 2] = w[i].end().x;
            out[4 * i + 3] = w[i].end().y;
        }
    return (int64_t)w.size();
}

void *dmxh_graph_new_map(void *file, double spacing) {
    dmx::GraphFile *f = static_cast<dmx::GraphFile *>(file);
    void *out = nullptr;
    guarded([&] {
        f->addNewPointMap();
        f->setGrid(spacing);
        out = &f->getDisplayedPointMap();
        return true;
    });
    return out;
}

int dmxh_graph_make_graph(void *file, int boundarygraph, double maxdist) {
    return guarded([&] { return static_cast<dmx::GraphFile *>(file)->makeGraph(nullptr, boundarygraph != 0, maxdist); });
}

void dmxh_graph_made(void *file) { static_cast<dmx::GraphFile *>(file)->graphMade(); }

void *dmxh_map_graph(void *map) { return static_cast<dmx::PointMap *>(map)->graph(); }

void dmxh_release_context(void) { dmx::release_shared_context(); }

}  // extern "C"
