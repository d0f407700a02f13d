// Flat C view of the host layer for ctypes (include/vga_host.h).
#include "../../include/vga_host.h"

#include <cstring>
#include <string>

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

int dmxh_map_columns(void *map, char *buf, int buflen) {
    const dmx::AttributeTable &t = static_cast<dmx::PointMap *>(map)->getAttributeTable();
    std::string s;
    for (size_t i = 0; i < t.getNumColumns(); i++) s += t.getColumnName(i) + "\n";
    std::strncpy(buf, s.c_str(), (size_t)buflen - 1);
    buf[buflen - 1] = 0;
    return (int)t.getNumColumns();
}

int dmxh_map_attr(void *map, const char *name, float *out) {
    const dmx::AttributeTable &t = static_cast<dmx::PointMap *>(map)->getAttributeTable();
    int c = t.getColumnIndex(name);
    if (c < 0) return 0;
    std::memcpy(out, t.column(c).data(), t.getNumRows() * sizeof(float));
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

void *dmxh_map_graph(void *map) { return static_cast<dmx::PointMap *>(map)->graph(); }

void dmxh_release_context(void) { dmx::release_shared_context(); }

}  // extern "C"
