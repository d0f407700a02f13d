/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the depthmapX visibility-graph hot path.
 *
 * Plain C restatement of the reference algorithm, written from the behaviour of
 *   salalib/pointdata.cpp:1246-1565   (sparkGraph2 / sparkPixel2 / sieve2)
 *   salalib/sparksieve2.cpp:33-173    (gap list, block, collectgarbage, tanify, testblock)
 *   genlib/p2dpoly.cpp:247-363,626-667 (intersect_region, Line ctor, intersect_line, crop)
 *   salalib/pointdata.h:353-367,432-520 (depixelate, regionate, whichbin)
 *   salalib/ngraph.cpp:27-58,234-304,392-416 (Node::make / Bin::make / iteration order)
 *   salalib/vgamodules/vgavisualglobal.cpp:23-240, vgavisuallocal.cpp:23-117
 *   genlib/pafmath.h:61-80 (log2, dvalue, pvalue, teklinteg)
 *
 * PARITY IS PINNED: tests/test_oracle_vs_reference.py checks every output of this file against
 * the unmodified reference compiled into oracle/_ref/libdmxref.so, and tests/golden/ holds
 * fixtures generated from that library (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libvga_b200.so) never does.
 */
#ifndef VGA_ORACLE_H
#define VGA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int32_t cols, rows;
    double spacing, bl_x, bl_y; /* bl = centre of cell (0,0) */
    double maxdist;             /* -1.0 = unlimited */
    const uint16_t *state;      /* cols*rows, index x*rows+y, Point::m_state flags */
    const uint32_t *line_off;   /* cols*rows+1 */
    const double *lines;        /* 5 per segment: bl.x, bl.y, tr.x, tr.y, parity */
} vgao_grid;

typedef struct vgao_graph vgao_graph;

/* makegraph: returns a handle holding, per filled cell (x-major ordinal v):
 *   acc row  : accepted targets in reference order (q, depth, gap, ind) with bin ids,
 *   iter row : the adjacency the reference iterates (Node::first/next): bins 0..31, each bin
 *              sorted as Bin::make stores it, diagonal bins expanded first..last,
 *   far_bin_dists[32], connectivity, first/second moments (float, as setValue stores them),
 *   bin node counts (uint16, as stored), grid-connection byte. */
vgao_graph *vgao_makegraph(const vgao_grid *g);
/* same, restricted to the x-major source ordinals [src_begin, src_end) */
vgao_graph *vgao_makegraph_range(const vgao_grid *g, int64_t src_begin, int64_t src_end);
/* adjacency given directly (iterated order irrelevant), refs are packed PixelRef ints */
vgao_graph *vgao_graph_from_edges(const vgao_grid *g, const uint64_t *rowptr, const int32_t *ref);
void vgao_graph_free(vgao_graph *gr);

int64_t vgao_num_cells(const vgao_graph *gr);   /* N = filled cells */
int64_t vgao_num_acc(const vgao_graph *gr);     /* sum Connectivity */
int64_t vgao_num_iter(const vgao_graph *gr);    /* sum iterated row sizes */
void vgao_cell_refs(const vgao_graph *gr, int32_t *ref /*N*/);  /* packed PixelRef per ordinal */
void vgao_acc_rows(const vgao_graph *gr, uint64_t *rowptr, int32_t *ref, uint8_t *bin);
void vgao_iter_rows(const vgao_graph *gr, uint64_t *rowptr, int32_t *ref, uint8_t *bin);
void vgao_node_attrs(const vgao_graph *gr, float *connectivity, float *first_moment, float *second_moment,
                     float *far_bin_dists /*N*32*/, uint16_t *bin_count /*N*32*/, uint8_t *gridconn /*N*/);

/* global BFS over the iterated adjacency; radius -1 = n.  dist: N*maxl ints, zero padded;
 * nlevels[v] = distribution.size() of the reference (may include a trailing 0). */
int vgao_global(const vgao_graph *gr, int radius, int64_t src_begin, int64_t src_end, int32_t *total_nodes,
                int64_t *total_depth, int32_t *dist, int32_t maxl, int32_t *nlevels);
/* the seven output columns exactly as row.setValue stores them (float) */
void vgao_global_formulas(int64_t n, const int32_t *total_nodes, const int64_t *total_depth, const int32_t *dist,
                          int32_t maxl, const int32_t *nlevels, float *node_count, float *mean_depth,
                          float *integ_hh, float *integ_pv, float *integ_tk, float *entropy, float *rel_entropy);

/* local measures: integers + float32 control in sorted order, then the three columns */
int vgao_local(const vgao_graph *gr, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k,
               int32_t *total, float *control);
void vgao_local_formulas(int64_t n, const int64_t *cluster, const int32_t *k, const int32_t *total,
                         const float *control, float *clustering, float *control_out, float *controllability);

/* vgao_global / vgao_local over an adjacency given as CSR of vertex ordinals (entry = col[e] >> shift; ordinals >= n are
 * ghost cells), for sampled sources / cells of the full-size configurations; no copy of the adjacency; callers may
 * split the samples over threads.  ref = packed PixelRef of every vertex (n cells, then ghosts). */
int vgao_global_csr(int64_t n, const uint64_t *rowptr, const uint32_t *col, int shift, int radius, const int64_t *src,
                    int64_t nsrc, int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t maxl);
int vgao_local_csr(int64_t n, int64_t nv, const uint64_t *rowptr, const uint32_t *col, int shift, const int32_t *ref,
                   const int64_t *cells, int64_t ncells, int64_t *cluster, int32_t *k, int32_t *total, float *control);

/* visual step depth from a set of source cells (x-major ordinals): depth[N], -1 = not reached
 * (salalib/vgamodules/vgavisualglobaldepth.cpp:23-75) */
int vgao_step_depth(const vgao_graph *gr, const int32_t *src, int64_t nsrc, int32_t *depth);

/* known-answer access to the sieve for salaTest/testsparksieve.cpp:21-83:
 * centre (cx,cy), octant q, nlines segments as 4 doubles (x1,y1,x2,y2); after block+collectgarbage
 * returns the number of gaps and writes up to cap (start,end) pairs. */
int vgao_sieve_kat(double cx, double cy, int q, const double *segs, int nsegs, double *gaps, int cap);

/* metric / angular VGA (row f4): the four "Metric ..." columns (Mean Shortest-Path Angle, Mean Shortest-Path Distance,
 * Mean Straight-Line Distance, Node Count) and the three "Angular ..." columns (Mean Depth, Total Depth, Node Count)
 * as row.setValue stores them; radius -1.0 = n; spacing = PointMap::m_spacing; partner = merge links as ordinals
 * ([N], -1 = none) or NULL. */
int vgao_metric(const vgao_graph *gr, const int32_t *partner, double spacing, double radius, int64_t src_begin, int64_t src_end,
                float *mspa, float *mspl, float *msld, float *count);
int vgao_angular(const vgao_graph *gr, const int32_t *partner, double radius, int64_t src_begin, int64_t src_end,
                 float *mean_depth, float *total_depth, float *count);

#ifdef __cplusplus
}
#endif
#endif
