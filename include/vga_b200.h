/* vga_b200.h -- C ABI of libvga_b200.so: the B200 (sm_100a) implementation of depthmapX's
 * visibility-graph hot path.  Plain C: pointers and sizes only, no C++/torch types.
 *
 * What each entry point replaces in the reference (/root/reference, depthmapX 0.8.0):
 *
 *   vga_graph_build*      PointMap::sparkGraph2 -> sparkPixel2 -> sieve2 -> sparkSieve2
 *                         salalib/pointdata.cpp:1246-1565, salalib/sparksieve2.cpp:33-173,
 *                         whichbin salalib/pointdata.h:432-520, Node::make/Bin::make
 *                         salalib/ngraph.cpp:27-58,234-304 (adjacency content + iteration set),
 *                         addGridConnections salalib/pointdata.cpp:1735-1768
 *   vga_graph_from_csr    the adjacency a loaded .graph holds in Point::m_node
 *                         (Node::first/next iteration, salalib/ngraph.cpp:158-191,392-416)
 *   vga_global            VGAVisualGlobal::run BFS part + extractUnseen
 *                         salalib/vgamodules/vgavisualglobal.cpp:66-130, 218-240
 *   vga_global_attributes the formula stage of VGAVisualGlobal::run :131-193 (host, glibc libm)
 *   vga_local             VGAVisualLocal::run salalib/vgamodules/vgavisuallocal.cpp:41-81
 *   vga_local_attributes  the formula stage :84-96
 *   vga_metric/vga_angular VGAMetric::run / VGAAngular::run salalib/vgamodules/vgametric.cpp:25-136,
 *                         vgaangular.cpp:22-133 (row f4)
 *
 * Conventions: every function returns 0 on success and a negative vga_status otherwise;
 * vga_last_error() gives the message of the last failure on the calling thread.  There is NO CPU
 * fallback: without a CUDA device every compute entry point fails with VGA_ERR_NO_DEVICE.
 * Unsupported inputs (NaN blocks, grids of 2^26 cells or more) are errors.  Merge links and context-filled cells
 * are handled above this ABI (contracted adjacency via vga_graph_from_csr, vga_graph_set_noexpand).
 * One process drives one GPU (vga_ctx_create(device)); multi-GPU runs shard sources across
 * processes (src_begin/src_end arguments) and exchange shards outside this library.
 *
 * Vertex numbering: "ordinal" = index of a FILLED cell in x-major order (for x: for y:), the
 * iteration order of every hot loop of the reference and of its attribute table
 * (AttributeKey = int(PixelRef), salalib/pixelref.h:81-82).  "Ghost" vertices (ordinals
 * N..N+G-1) are the unfilled cells, numbered in x-major order: an unfilled cell can be covered by
 * a diagonal bin's first..last run (Bin::make, salalib/ngraph.cpp:243-259) and is then a member
 * of neighbourhoods in the local measures, although it has no row of its own and is never a BFS
 * source or counted.  Graphs made by vga_graph_build number ALL unfilled cells (G = cells - N) so
 * that shards built on different GPUs agree without communication.
 */
#ifndef VGA_B200_H
#define VGA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    VGA_OK = 0,
    VGA_ERR_INVALID = -1,     /* bad argument */
    VGA_ERR_NO_DEVICE = -2,   /* no CUDA device / driver: there is no CPU path */
    VGA_ERR_CUDA = -3,        /* CUDA runtime error, see vga_last_error() */
    VGA_ERR_UNSUPPORTED = -4, /* input outside the supported subset (see header comment) */
    VGA_ERR_CAPACITY = -5,    /* a fixed-capacity device structure overflowed */
    VGA_ERR_CANCELLED = -6    /* the cancel callback returned non-zero */
} vga_status;

typedef struct vga_ctx vga_ctx;
typedef struct vga_graph vga_graph;

/* Flat image of what PointMap holds when sparkGraph2 starts (after blockLines + makePoints):
 * Point::m_state and Point::m_lines for every cell (salalib/point.h:32-64). */
typedef struct {
    int32_t cols, rows;       /* PointMap::m_cols, m_rows */
    double spacing;           /* m_spacing */
    double bl_x, bl_y;        /* m_bottom_left = centre of cell (0,0) */
    double maxdist;           /* sparkGraph2's maxdist, -1.0 = unlimited */
    const uint16_t *state;    /* [cols*rows] index x*rows+y; FILLED = 0x0002 */
    const uint32_t *line_off; /* [cols*rows+1] offsets into lines */
    const double *lines;      /* 5 doubles per cropped segment: bl.x, bl.y, tr.x, tr.y, parity */
} vga_grid;

/* Progress / cancel callbacks (Communicator::CommPostMessage / IsCancelled, genlib/comm.h).
 * progress(user, done, total); cancel(user) != 0 aborts with VGA_ERR_CANCELLED. */
typedef void (*vga_progress_fn)(void *user, int64_t done, int64_t total);
typedef int (*vga_cancel_fn)(void *user);

/* Stage timings of the most recent call on a context, CUDA-event measured (milliseconds). */
typedef struct {
    double h2d_ms;     /* host->device copies of inputs */
    double kernel_ms;  /* all kernels of the call */
    double d2h_ms;     /* device->host copies of results */
    double main_kernel_ms;   /* the dominant kernel(s) only: BFS level kernels / sieve passes / local */
    int64_t launches;  /* kernels launched by the call */
    int64_t main_launches;
    double algo_bytes; /* algorithmic bytes of the dominant kernel(s), DESIGN.md formula, rows in the format the
                          kernels read (vga_global: 4 bytes per pyramid-node id of a run-length row) */
    double algo_bytes_csr; /* vga_global: the same model with 4-byte CSR entries (SURVEY.md 8d as written), else 0 */
    double prep_ms;    /* vga_global: one-time derivation of the BFS row lists of the graph (first call only) */
    int64_t batch_words; /* vga_global: 64-bit words per vertex and batch used (a batch = 64 * words sources) */
} vga_timing;

const char *vga_last_error(void);
const char *vga_version(void);
/* number of visible CUDA devices; 0 when there is no driver/device (never an error) */
int vga_device_count(void);

int vga_ctx_create(int device, vga_ctx **out);
void vga_ctx_destroy(vga_ctx *ctx);
int vga_ctx_set_callbacks(vga_ctx *ctx, vga_progress_fn progress, vga_cancel_fn cancel, void *user);
/* tuning knobs, also readable from the environment (VGA_BFS_MODE, VGA_BFS_CHUNK, ...);
 * unknown keys are VGA_ERR_INVALID */
int vga_ctx_set_option(vga_ctx *ctx, const char *key, int64_t value);
int vga_ctx_timing(const vga_ctx *ctx, vga_timing *out);
/* wait for the context's stream + error check */
int vga_ctx_sync(vga_ctx *ctx);

/* ---- construction (makegraph) ------------------------------------------------------------- */

/* Build the visibility graph rows of the sources with ordinals [src_begin, src_end)
 * (src_end < 0 = all).  The graph always knows all N cells. */
int vga_graph_build(vga_ctx *ctx, const vga_grid *grid, int64_t src_begin, int64_t src_end, vga_graph **out);

/* Two-step variant for measurements with inputs resident in HBM: upload once, build many times. */
typedef struct vga_dgrid vga_dgrid;
int vga_grid_upload(vga_ctx *ctx, const vga_grid *grid, vga_dgrid **out);
void vga_dgrid_free(vga_dgrid *g);
int vga_graph_build_resident(vga_ctx *ctx, const vga_dgrid *grid, int64_t src_begin, int64_t src_end,
                             vga_graph **out);

/* Adopt an adjacency that already exists (e.g. flattened from the Nodes of a loaded .graph, or the
 * concatenation of per-rank shards).  n_cells filled cells, n_ghosts ghost vertices, rows sorted or
 * not; col values < n_cells + n_ghosts.  bin may be NULL. */
int vga_graph_from_csr(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, const uint64_t *rowptr,
                       const uint32_t *col, const uint8_t *bin, vga_graph **out);

void vga_graph_free(vga_graph *g);

int64_t vga_graph_num_cells(const vga_graph *g);  /* N */
int64_t vga_graph_num_ghosts(const vga_graph *g); /* G */
int64_t vga_graph_num_edges(const vga_graph *g);  /* iterated adjacency entries held */
int64_t vga_graph_src_begin(const vga_graph *g);
int64_t vga_graph_src_end(const vga_graph *g);

/* Rows src_begin..src_end, each sorted by column ordinal (= PixelRef x-major order, ghosts last).
 * rowptr has (src_end-src_begin+1) entries starting at 0.  accepted[e] = 1 for cells the sieve
 * accepted (they count towards Connectivity and the moments), 0 for diagonal first..last run
 * fill-ins.  Any output pointer may be NULL. */
int vga_graph_csr(const vga_graph *g, uint64_t *rowptr, uint32_t *col, uint8_t *bin, uint8_t *accepted);
/* Packed PixelRef ((x<<16)+(y&0xffff)) of every vertex: N cells then G ghosts. */
int vga_graph_cell_refs(const vga_graph *g, int32_t *ref);
/* Attach packed PixelRefs (N cells, then G ghosts) to an adopted graph so that the analyses can form
 * spatially coherent source batches.  Optional: results never depend on it, only speed does. */
int vga_graph_set_cell_refs(vga_graph *g, const int32_t *ref, int64_t count);
/* Context-filled cells (GUI semi-fill: FILLED | CONTEXTFILLED, salalib/pointdata.cpp:435-441) that are not "even"
 * (PixelRef::iseven) are counted but NOT expanded by a radius-limited vga_global (vgavisualglobal.cpp:108-110) and by
 * vga_step_depth beyond level 0 (vgavisualglobaldepth.cpp:53).  flags[v] != 0 marks such a cell (N bytes, x-major
 * ordinals); NULL clears the marks.  vga_global with radius -1 ignores them, as the reference does.  Skipping those
 * cells as SOURCES is the caller's business (the host layer and the shims do not write their rows). */
int vga_graph_set_noexpand(vga_graph *g, const uint8_t *flags);
/* Per source row: what sparkPixel2 stores besides the pixel lists.  connectivity = neighbourhood_size,
 * sum_d / sum_d2 = total_dist / total_dist_sqr (double running sums in reference order),
 * far_bin_dists [rows*32] floats, bin_count [rows*32] (accepted pixels per bin),
 * grid_connections [rows] bytes.  Any pointer may be NULL. */
int vga_graph_node_stats(const vga_graph *g, int32_t *connectivity, double *sum_d, double *sum_d2,
                         float *far_bin_dists, int32_t *bin_count, uint8_t *grid_connections);

/* ---- analysis ----------------------------------------------------------------------------- */

/* All-sources BFS for sources [src_begin, src_end) over the (complete) graph.  radius -1 = n.
 * Outputs (host): total_nodes[k], total_depth[k], dist[k*max_levels] level histogram (level 0 =
 * the source itself), *levels_used = 1 + deepest non-empty level over these sources.  If the
 * histogram needs more than max_levels columns the call fails with VGA_ERR_CAPACITY and sets
 * *levels_used to the required value. */
int vga_global(vga_ctx *ctx, const vga_graph *g, int radius, int64_t src_begin, int64_t src_end,
               int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels,
               int32_t *levels_used);

/* The same for an explicit list of source ordinals (any order; the library forms spatially coherent batches itself).
 * Outputs are in list order. */
int vga_global_sources(vga_ctx *ctx, const vga_graph *g, int radius, const int64_t *sources, int64_t n_sources,
                       int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels,
                       int32_t *levels_used);
/* The order in which vga_global batches all N sources (order[i] = ordinal of the i-th source; consecutive groups of
 * 64*W entries are the spatially compact batches).  Lets a caller pick representative subsets of whole batches. */
int vga_graph_batch_order(vga_ctx *ctx, const vga_graph *g, int32_t *order);

/* Sizes of the row lists the BFS reads (derived on first use): runs and pyramid-node ids of the out-rows and in-rows. */
int vga_graph_list_sizes(vga_ctx *ctx, const vga_graph *g, int64_t *out_runs, int64_t *out_nodes, int64_t *in_runs,
                         int64_t *in_nodes);

/* Formula stage (host, FP64 -> float exactly as AttributeRow::setValue stores them); -1 sentinels
 * as in the reference.  Any output may be NULL. */
int vga_global_attributes(int64_t n, const int32_t *total_nodes, const int64_t *total_depth, const int32_t *dist,
                          int32_t max_levels, float *node_count, float *mean_depth, float *integ_hh,
                          float *integ_pv, float *integ_tk, float *entropy, float *rel_entropy);

/* Local measures for cells [src_begin, src_end): cluster = sum_u |iter(N(u)) n N(v)|, k = |N(v)|,
 * total = |U_u iter(N(u))|, control = float32 running sum of 1/retro_size(u) in PixelRef order. */
int vga_local(vga_ctx *ctx, const vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster,
              int32_t *k, int32_t *total, float *control);
int vga_local_attributes(int64_t n, const int64_t *cluster, const int32_t *k, const int32_t *total,
                         const float *control, float *clustering, float *control_out, float *controllability);

/* Visual step depth (VGAVisualGlobalDepth::run,
 * salalib/vgamodules/vgavisualglobaldepth.cpp:23-75): BFS from the SET of cells `sources` (x-major
 * ordinals, the map's selection); depth[v] for all N cells = level at which v is first reached, -1 if
 * never.  No merge links / context fill. */
int vga_step_depth(vga_ctx *ctx, const vga_graph *g, const int64_t *sources, int64_t n_sources, int32_t *depth);

/* Metric / angular VGA (VGAMetric::run salalib/vgamodules/vgametric.cpp:25-136, VGAAngular::run
 * salalib/vgamodules/vgaangular.cpp:22-133, extractMetric / extractAngular salalib/ngraph.cpp:67-87, 329-365): per
 * source a shortest-path search over the iterated adjacency with float32 keys (path length in cells / cumulated turn in
 * units of 90 degrees), pops in the order of the reference's std::set<(key, pixel)>, float32 sums in pop order.
 * blocked_adjacent[v] != 0 (N bytes, x-major ordinals) for cells that are BLOCKED or have a BLOCKED cell among their
 * eight neighbours (Point::blocked / PointMap::blockedAdjacent, salalib/pointdata.cpp:1016-1066): only those and the
 * source are expanded.  The graph needs cell coordinates (vga_graph_build gives them, else vga_graph_set_cell_refs).
 * sources = ordinals (NULL = all N cells, n_sources ignored); outputs in list order, exactly the values
 * AttributeRow::setValue receives: vga_metric -> "Metric Mean Shortest-Path Angle", "Metric Mean Shortest-Path
 * Distance", "Metric Mean Straight-Line Distance", "Metric Node Count"; vga_angular -> "Angular Mean Depth", "Angular
 * Total Depth", "Angular Node Count".  radius -1.0 = n (metric: in map units, compared with key * spacing; angular: in
 * units of 90 degrees).  Any output may be NULL.  *angle_unsafe = number of turn-angle evaluations whose float32
 * rounding could depend on the last bits of acos (0 = the angle sums are guaranteed bit-equal to a glibc host; distances
 * and node counts of vga_metric never depend on it).  merge_partner[v] (N ints, NULL = no merge links) = ordinal of the
 * cell v is merged with (Point::m_merge), -1 = none, symmetric: when v is finalised its partner is expanded from the same
 * key and finalised without being counted (vgametric.cpp:96-104, vgaangular.cpp:91-99); g is the PLAIN adjacency, not the
 * contracted one of vga_global. */
int vga_metric(vga_ctx *ctx, const vga_graph *g, const uint8_t *blocked_adjacent, const int32_t *merge_partner, double spacing,
               double radius,
               const int64_t *sources, int64_t n_sources, float *mean_angle, float *mean_path_dist, float *mean_line_dist,
               float *node_count, int64_t *angle_unsafe);
int vga_angular(vga_ctx *ctx, const vga_graph *g, const uint8_t *blocked_adjacent, const int32_t *merge_partner, double radius,
                const int64_t *sources,
                int64_t n_sources, float *mean_depth, float *total_depth, float *node_count, int64_t *angle_unsafe);

/* ---- device-resident access for multi-GPU plumbing (pointers are CUDA device pointers) ------ */

/* Device pointers of the sorted shard rows (valid until vga_graph_free): rowptr (u64, local,
 * rows+1), packed adjacency entries (u32: col<<6 | accepted<<5 | bin), number of entries. */
int vga_graph_device_rows(const vga_graph *g, const uint64_t **d_rowptr, const uint32_t **d_adj, int64_t *n_entries);
/* Adopt device-resident packed rows for all N cells (e.g. after an NCCL all-gather of shards). */
int vga_graph_from_device_rows(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, const uint64_t *d_rowptr,
                               const uint32_t *d_adj, int64_t n_entries, vga_graph **out);

/* Run-length form of the shard's rows: what the BFS reads and what a multi-GPU run exchanges (8 bytes per run instead of
 * 4 bytes per entry: 1.3 GB instead of 22 GB at 10^6 cells).  runptr (u64, local, rows+1), runs = pairs of u32 (first
 * column ordinal, length) sorted by first ordinal; no run straddles N, the ghost columns (>= N) are the last runs of a
 * row.  Valid until vga_graph_free. */
int vga_graph_device_runs(vga_ctx *ctx, const vga_graph *g, const uint64_t **d_runptr, const void **d_runs, int64_t *n_runs);
/* Adopt device-resident run-length rows of all N cells (e.g. the all-gathered shards).  The graph serves vga_global /
 * vga_global_sources only (no entries, bins or statistics).  d_degree (u32 [N], entries per row, may be NULL) only
 * feeds the CSR byte model of vga_timing. */
int vga_graph_from_device_runs(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, const uint64_t *d_runptr, const void *d_runs,
                               int64_t n_runs, const uint32_t *d_degree, vga_graph **out);

/* The same without the staging copy: the library allocates the run-length rows of a BFS-only graph and hands out
 * WRITABLE device pointers (runptr u64 [N+1], runs 2 x u32 [n_runs], degree u32 [N], zeroed); the caller fills them
 * (e.g. NCCL broadcasts of every rank's shard straight into their final place) and then calls vga_graph_runs_commit. */
int vga_graph_runs_alloc(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, int64_t n_runs, vga_graph **out,
                         uint64_t **d_runptr, void **d_runs, uint32_t **d_degree);
int vga_graph_runs_commit(vga_graph *g);
/* entries per row (u32) of the rows this graph holds, device-resident (valid until vga_graph_free) */
int vga_graph_device_degrees(vga_ctx *ctx, const vga_graph *g, const uint32_t **d_degree);

#ifdef __cplusplus
}
#endif
#endif /* VGA_B200_H */
