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
int dmxh_map_filled_count(void *map);
/* flat hot-path inputs (vga_grid arrays); call with NULL arrays to get the sizes */
void dmxh_map_flat(void *map, int64_t *cells, int64_t *nseg, uint16_t *state, uint32_t *line_off, double *lines);
int dmxh_map_make_graph(void *map, int boundarygraph, double maxdist);
int dmxh_map_vga_global(void *map, double radius, int simple_version);
int dmxh_map_vga_local(void *map, int simple_version);
int dmxh_map_columns(void *map, char *buf, int buflen); /* '\n' separated, returns the count */
int dmxh_map_attr(void *map, const char *name, float *out /* one per filled cell, x-major */);
int dmxh_map_grid_connections(void *map, uint8_t *out /* one per filled cell */);
void *dmxh_map_graph(void *map); /* vga_graph* of include/vga_b200.h, owned by the map */
void dmxh_release_context(void);

#ifdef __cplusplus
}
#endif
#endif
