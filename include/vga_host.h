/* vga_host.h -- flat C view of the C++ host layer (depthmapx_b200/host/pointmap.h) for ctypes.
 *
 * The host layer mirrors the reference's operator interface for the hot path:
 *   dmxh_map_create + dmxh_map_fill   PointMap(region, drawing) + setGrid + makePoints
 *                                     (salalib/pointdata.cpp:122-171, 296-357, 402-514)
 *   dmxh_map_make_graph               PointMap::sparkGraph2 (pointdata.cpp:1246)   [GPU]
 *   dmxh_map_vga_global / _local      VGAVisualGlobal::run / VGAVisualLocal::run   [GPU]
 * Functions return 1 for true, 0 for false (the reference's bool), negative on exceptions
 * (message in dmxh_last_error()).
 */
#ifndef VGA_HOST_H
#define VGA_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char *dmxh_last_error(void);
/* walls: 4 doubles per segment (x1,y1,x2,y2); the parent region is their bounding box */
void *dmxh_map_create(const double *walls, int nwalls, double spacing);
void dmxh_map_destroy(void *map);
void dmxh_map_grid(void *map, int32_t *cols, int32_t *rows, double *spacing, double *bl_x, double *bl_y);
int dmxh_map_block_lines(void *map);
int dmxh_map_fill(void *map, double x, double y);
/* fill_type 0 = full fill, 1 = semi-fill (FILLED | CONTEXTFILLED, the GUI's "context fill", pointdata.cpp:435-441) */
int dmxh_map_fill_type(void *map, double x, double y, int fill_type);
/* 1 if the map has context-filled cells that the analyses skip (flags: one byte per filled cell, may be NULL) */
int dmxh_map_context_skip(void *map, uint8_t *flags);
int dmxh_map_filled_count(void *map);
/* flat hot-path inputs (vga_grid arrays); call with NULL arrays to get the sizes */
void dmxh_map_flat(void *map, int64_t *cells, int64_t *nseg, uint16_t *state, uint32_t *line_off, double *lines);
int dmxh_map_make_graph(void *map, int boundarygraph, double maxdist);
int dmxh_map_vga_global(void *map, double radius, int simple_version);
int dmxh_map_vga_local(void *map, int simple_version);
/* dmx::VGAMetric(radius, false).run / dmx::VGAAngular(radius, false).run; radius -1.0 = n */
int dmxh_map_vga_metric(void *map, double radius);
int dmxh_map_vga_angular(void *map, double radius);
int dmxh_map_columns(void *map, char *buf, int buflen); /* '\n' separated, returns the count */
int64_t dmxh_map_num_rows(void *map); /* attribute rows = cells that had a Node made */
int dmxh_map_attr(void *map, const char *name, float *out /* dmxh_map_num_rows values, x-major */);
int dmxh_map_grid_connections(void *map, uint8_t *out /* one per filled cell */);
void *dmxh_map_graph(void *map); /* vga_graph* of include/vga_b200.h, owned by the map */
int dmxh_map_state(void *map, uint16_t *state /* cols*rows, x-major */);

/* Visual step depth (VGAVisualGlobalDepth::run, salalib/vgamodules/vgavisualglobaldepth.cpp:23-75) from the cells
 * containing the given points (x,y pairs; PointMap::setCurSel(QtRegion(p,p), true), pointdata.cpp:939)  [GPU] */
int dmxh_map_step_depth(void *map, const double *points, int npoints);
int dmxh_map_select(void *map, const double *points, int npoints);
int64_t dmxh_map_selection(void *map, int32_t *refs /* packed PixelRefs, NULL = count only */);

/* Run-length adjacency (the reference's Nodes, salalib/ngraph.h:31-149) as flat rows in Node::first/next order:
 * call with NULL arrays for the sizes.  For a map built on the GPU the rows come from the device. */
int dmxh_map_flat_rows(void *map, int64_t *n, int64_t *entries, uint64_t *rowptr, int32_t *ref, uint8_t *bin);
int dmxh_map_bins(void *map, uint16_t *bin_count /* N*32 */, float *bin_dist /* N*32 */);
/* Node::make / Bin::make encoder (ngraph.cpp:27-58, 234-304) from flat rows; `accepted` may be NULL */
int dmxh_map_encode_nodes(void *map, const uint64_t *rowptr, const int32_t *ref, const uint8_t *bin,
                          const uint8_t *accepted, const float *far_bin_dists);

/* Merge links (`-m LINK -lnk x1,y1,x2,y2`: linkutils::pixelateMergeLines + mergePixelPairs, salalib/linkutils.cpp:20-98,
 * PointMap::mergePixels pointdata.cpp:1653-1685).  The BFS analyses then run on the contracted adjacency (a merged pair
 * is one vertex); dmxh_map_contracted_rows returns it: rowptr [N+1], col, primary [N] = the ordinal whose
 * results each cell takes (NULL arrays = sizes only). */
int dmxh_map_merge(void *map, double ax, double ay, double bx, double by);
int dmxh_map_contracted_rows(void *map, int64_t *n, int64_t *entries, uint64_t *rowptr, uint32_t *col, int32_t *primary);

/* Radius-limited global analysis of a merged map: adds the second count of every pair whose two cells are both first
 * reached AT the radius (see dmx::PointMap::radiusCorrection) to the integers of the contracted BFS.  The BFS "level of
 * every source to a vertex set" over the transposed contracted adjacency is supplied by the caller: the analysis classes
 * use vga_step_depth on the GPU; the CPU tests plug in the checker. */
typedef void (*dmxh_level_prepare_fn)(void *user, int64_t n, const uint64_t *t_rowptr, const uint32_t *t_col);
typedef void (*dmxh_level_run_fn)(void *user, const int64_t *seeds, int64_t nseeds, int32_t *level /* [n], preset -1 */);
int dmxh_map_radius_correction(void *map, int radius, dmxh_level_prepare_fn prepare, dmxh_level_run_fn run, void *user,
                               int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels);

/* Attribute stages on their own, fed with the integers libvga_b200 produced (on this GPU or gathered from the
 * ranks of a multi-GPU run): sparkGraph2's host halves (pointdata.cpp:1250-1264, 1268-1341) and the column +
 * formula stages of the three analyses. */
int dmxh_map_begin_graph(void *map, int boundarygraph);
int dmxh_map_finish_graph(void *map, int boundarygraph, const int32_t *connectivity, const double *sum_d,
                          const double *sum_d2, const uint8_t *grid_connections);
int dmxh_map_write_global(void *map, double radius, int simple_version, const int32_t *total_nodes,
                          const int64_t *total_depth, const int32_t *dist, int32_t max_levels);
int dmxh_map_write_local(void *map, int simple_version, const int64_t *cluster, const int32_t *k, const int32_t *total,
                         const float *control);
int dmxh_map_write_step_depth(void *map, const int32_t *depth);

/* .graph container (MetaGraph::readFromStream / write, salalib/mgraph.cpp:2492-2763): the
 * PointMap section is decoded / encoded here, the other sections are carried through verbatim. */
void *dmxh_graph_open(const char *path); /* NULL on failure, see dmxh_last_error() */
void dmxh_graph_close(void *file);
int dmxh_graph_save(void *file, const char *path);
int dmxh_graph_num_maps(void *file);
int dmxh_graph_displayed_map(void *file);
void *dmxh_graph_map(void *file, int i); /* owned by the file; usable with every dmxh_map_* call */
int64_t dmxh_graph_walls(void *file, double *out /* 4 per segment, NULL = count only */);
void *dmxh_graph_new_map(void *file, double spacing); /* MetaGraph::addNewPointMap + setGrid */
int dmxh_graph_make_graph(void *file, int boundarygraph, double maxdist); /* MetaGraph::makeGraph  [GPU] */
void dmxh_graph_made(void *file); /* state / view-class update of MetaGraph::makeGraph after dmxh_map_finish_graph */
void dmxh_release_context(void);

#ifdef __cplusplus
}
#endif
#endif
