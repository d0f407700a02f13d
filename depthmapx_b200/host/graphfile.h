// dmx::GraphFile -- the .graph container around the PointMap section.
//
// Mirrors what MetaGraph::readFromStream / MetaGraph::write do for the visibility-graph path
// (salalib/mgraph.cpp:2492-2654, 2656-2763, 2811-2845): header "grf" + version 440, state / view class words,
// 'x' file properties, 'l' drawing layers (SpacePixelFile / ShapeMap / SalaShape, salalib/spacepixfile.cpp:28-58,
// salalib/shapemap.cpp:49-78, 2273-2383), 'p' point maps, then shape graphs ('x') and data maps ('s').
//
// The drawing section is parsed to recover the region and the wall segments of the shown layers in
// MetaGraph::getVisibleDrawingLines / ShapeMap::getAllShapesAsLines order (mgraph.cpp:2785-2797,
// shapemap.cpp:3275-3292) -- the input of PointMap::blockLines -- but it is written back verbatim, as are the
// sections that follow the point maps: this reader owns only the 'p' section, which it decodes and re-encodes
// itself.  Version < 440 files and the deprecated 'd' / 'v' sections are rejected (the reference converts them
// through its mgraph440 library, which is outside the path).
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "pointmap.h"

namespace dmx {

class GraphFile {
  public:
    enum { OK = 0, NOT_A_GRAPH = 1, NEWER_VERSION = 2, UNSUPPORTED = 3, DAMAGED_FILE = 4, DISK_ERROR = 5 };
    enum { POINTMAPS = 0x0002, LINEDATA = 0x0004, ANGULARGRAPH = 0x0010, DATAMAPS = 0x0020, SHAPEGRAPHS = 0x0100 };
    enum { VIEWVGA = 0x01, VIEWBACKVGA = 0x02, VIEWAXIAL = 0x04, VIEWBACKAXIAL = 0x08, VIEWDATA = 0x20, VIEWBACKDATA = 0x40 };
    static const int METAGRAPH_VERSION = 440;

    int read(const std::string &filename);
    int readFromBuffer(const char *data, size_t size);
    int write(const std::string &filename);
    std::string lastError() const { return m_error; }

    int getState() const { return m_state; }
    int getViewClass() const { return m_view_class; }
    const Region &getRegion() const { return m_region; }
    // wall segments of the shown drawing layers, in blockLines order
    const std::vector<Line> &getVisibleDrawingLines() const { return m_walls; }
    size_t getNumPointMaps() const { return m_point_maps.size(); }
    PointMap &getPointMap(size_t i) { return *m_point_maps[i]; }
    int getDisplayedPointMapRef() const { return m_displayed_pointmap; }
    PointMap &getDisplayedPointMap() { return *m_point_maps[(size_t)m_displayed_pointmap]; }

    // MetaGraph::addNewPointMap + MetaGraph::setGrid (mgraph.cpp:2799-2819, 222-234): a new map named
    // "VGA Map" (de-duplicated with a counter), displayed, state |= POINTMAPS, view class SHOWVGATOP
    int addNewPointMap(const std::string &name = "VGA Map");
    bool setGrid(double spacing, const Point2f &offset = Point2f());
    // MetaGraph::makePoints / makeGraph / analyseGraph wrappers only touch the state word (mgraph.cpp:264-268)
    bool makeGraph(Communicator *comm, bool boundarygraph, double maxdist);
    void graphMade();


  private:
    void showVgaTop();
    std::string m_error;
    std::string m_head;     // bytes from "grf" up to (not including) the 'p' type byte
    std::string m_tail;     // bytes after the point maps section
    int m_state = 0, m_view_class = 0;
    bool m_has_drawing = false;
    Region m_region;
    std::vector<Line> m_walls;
    std::vector<std::unique_ptr<PointMap>> m_point_maps;
    int m_displayed_pointmap = -1;
};

}  // namespace dmx
