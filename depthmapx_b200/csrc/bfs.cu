// All-sources BFS of VGA visibility analysis (VGAVisualGlobal::run + extractUnseen,
// salalib/vgamodules/vgavisualglobal.cpp:66-130, 218-240) as a bit-parallel multi-source BFS.
//
// B = 64*W sources form a batch (W = 1, 2 or 4 machine words); a batch carries W consecutive 64-bit
// words per vertex for each of visited / frontier / next, so one adjacency entry serves B sources and
// the per-vertex state of a batch is one 8/16/32-byte vector (a full 32-byte sector for W = 4).
// Batches are formed from spatially compact groups of sources (floods of 64 cells over the direct
// neighbour links, seeded along 8x8 tiles in Morton order) because cells that are close and not
// separated by a wall see almost the same set: their BFS levels coincide for most vertices, so a
// vertex joins the batch's frontier about once instead of once per distinct level.
// Many batches (a "chunk") advance level by level together: grid = (vertex tiles, batches), one
// launch per level serves the whole chunk and the host reads one flag per level.
//
// Per level and batch the step is direction-optimising:
//   push  (top-down)  every vertex with a non-zero frontier vector streams its adjacency row with
//                     coalesced 32-bit loads and ORs its words into `next` of unvisited targets
//                     (atomicOr on L2-resident words, filtered by a vector read of `visited`);
//   pull  (bottom-up) every vertex that still misses some source bit streams its in-row
//                     (transpose adjacency) and ORs the frontier vectors of its in-neighbours
//                     (8-lane groups, redux.or), leaving the row as soon as every missing bit is found.
// A pull step wastes a full row scan on every vertex that cannot be reached yet.  A cheap coarse pass
// removes most of that: one extra bit-parallel BFS in which bit j stands for ALL sources of group j
// (bfs_group*64 consecutive sources) gives lo[j][w] = min over the group of level(w), a lower bound
// for every batch of the group; the pull step skips w while lo[w] > level + 1.
// `update` folds `next` into visited/frontier and counts the new vertices per source with
// ballot + popc (bit b of lane l's word -> source b), i.e. the reference's distribution[level].
//
// The reference's `extents` run short-circuit is a pure optimisation; the result
// is the plain level structure computed here.  Ghost vertices (unfilled cells inside a diagonal
// run) are never counted or expanded by the reference (p.filled() test, :104) and are dropped.
#include <cub/cub.cuh>

#include <algorithm>
#include <memory>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "pyramid.cuh"
#include "vga_dev.cuh"

namespace vga {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TPB = 256;
constexpr int PYR_LEVELS_DEV = 28;  // pyramid levels passed to the kernels (n < 2^26 vertices, cabi.cu limit)

typedef unsigned long long u64;

struct BfsDev {
    int64_t n;  // filled vertices
    const uint64_t *rowptr;
    const uint32_t *adj;  // packed col<<6|..
    const uint64_t *t_rowptr;
    const uint32_t *t_col;
    u64 *visited, *frontier, *next;  // [batches][n][W]
    const u64 *valid;                // [batches*W] valid source bits of each word
    int *active;                     // [batches] 1 while the batch's frontier is non-empty
    int *mode;                       // [batches] 0 push, 1 pull
    u64 *stats;                      // [batches][NSTAT]: frontier edges, pull cost, new vertices, open vertices, run cost
    int *any;                        // [1] bit 0: some batch still active, bit 1: the level just folded added vertices
    const uint8_t *lvl_in;           // [groups][n] coarse lower-bound level, or nullptr
    uint8_t *lvl_out;                // [coarse batches*64][n] written by the coarse pass, or nullptr
    int group;                       // batches per coarse group
    // pyramid pull (bfs_pull = 1), nullptr / 0 otherwise
    u64 *pyr;                        // [batches][pyr_total][W]: OR-pyramid levels >= 1 of the frontier
    int64_t pyr_total;               // nodes per batch
    const uint64_t *t_runptr;        // [n+1]
    const uint2 *t_runs;             // (first ordinal, length) runs of the sorted in-rows
    const uint64_t *t_costptr;       // [n+1] prefix sums of the pyramid loads of a full in-row scan
    // pyramid push (bfs_push = 1), nullptr otherwise
    u64 *npyr;                       // [batches][pyr_total][W]: range-OR update nodes of `next` (levels >= 1), all zero
                                     // between levels
    const uint64_t *f_runptr;        // [n+1] runs of the (sorted) out-rows, ghost columns excluded
    const uint2 *f_runs;
    const uint64_t *f_costptr;       // [n+1] prefix sums of the pyramid nodes an out-row touches
    const uint32_t *f_nodes;         // bfs_pyr_nodes = 1: per row the ids of the pyramid nodes tiling its runs (id < n: leaf =
    const uint32_t *t_nodes;         // the vertex itself, id >= n: inner node id - n); row offsets = f_costptr / t_costptr
    int push_force;                  // bfs_push = 2: every top-down step uses the pyramid (tests)
    int pyr_weight;                  // percent: cost of one pyramid-push node relative to one adjacency entry (bfs_pyr_cost)
    const uint8_t *noexpand;         // [n] or nullptr: vertices that are counted but never join the frontier (context-filled,
                                     // not even cells under a radius limit)
    int pyr_levels;
    int64_t pyr_off[PYR_LEVELS_DEV];
    int64_t pyr_cnt[PYR_LEVELS_DEV];
};
constexpr int NSTAT = 8;  // per-batch statistics: frontier edges, pull cost, new vertices, open vertices, frontier run cost,
                          // frontier run count

template <int W> __device__ __forceinline__ void ldw(const u64 *p, u64 (&o)[W]) {
    if constexpr (W == 1) {
        o[0] = p[0];
    } else {
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(p + j);
            o[j] = t.x;
            o[j + 1] = t.y;
        }
    }
}
template <int W> __device__ __forceinline__ void stw(u64 *p, const u64 (&o)[W]) {
    if constexpr (W == 1) {
        p[0] = o[0];
    } else {
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            ulonglong2 t;
            t.x = o[j];
            t.y = o[j + 1];
            *reinterpret_cast<ulonglong2 *>(p + j) = t;
        }
    }
}


// source i of the ordered list -> bit (i & 63) of word (i >> 6); word wi lives in batch wi / W, slot wi % W
template <int W> __global__ void k_init(BfsDev d, const int32_t *src, int64_t nsrc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsrc) return;
    int64_t wi = i >> 6;
    int64_t b = wi / W;
    int j = (int)(wi % W);
    int64_t v = src[i];
    u64 w = 1ULL << (i & 63);
    d.visited[(b * d.n + v) * W + j] = w;  // a vertex is a source in exactly one word: one writer
    d.frontier[(b * d.n + v) * W + j] = w;
}

// With the pyramid push available, the step from level 0 (the sources' own rows) also goes through the pyramid when that
// is cheaper for the batch: one warp per batch sums the row lengths and the run costs of its sources.
template <int W>
__global__ void k_init_mode(BfsDev d, const int32_t *src, int64_t nsrc, int nbatch) {
    const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= nbatch) return;
    const int64_t s0 = (int64_t)b * 64 * W, s1 = s0 + 64 * W < nsrc ? s0 + 64 * W : nsrc;
    unsigned long long fe = 0, frc = 0;
    for (int64_t i = s0 + lane; i < s1; i += 32) {
        const int64_t v = src[i];
        fe += d.rowptr[v + 1] - d.rowptr[v];
        frc += d.f_costptr[v + 1] - d.f_costptr[v];
    }
    for (int o = 16; o > 0; o >>= 1) {
        fe += __shfl_down_sync(FULL, fe, o);
        frc += __shfl_down_sync(FULL, frc, o);
    }
    if (lane == 0 && (d.push_force || (frc + 2 * (u64)d.n) * (u64)d.pyr_weight < fe * 100ULL)) d.mode[b] = 2;
}

// coarse pass (W = 1): the sources of group j (`per_group` consecutive sources) all carry bit j%64 of
// coarse batch j/64, so one bit-parallel BFS yields 64 group lower bounds at once.
__global__ void k_init_coarse(BfsDev d, const int32_t *src, int64_t nsrc, int per_group) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsrc) return;
    int64_t grp = i / per_group;
    int64_t cb = grp >> 6;
    u64 w = 1ULL << (grp & 63);
    int64_t v = src[i];
    d.visited[cb * d.n + v] = w;
    d.frontier[cb * d.n + v] = w;
    d.lvl_out[grp * d.n + v] = 0;
}

// top-down step.  U = adjacency entries per lane and round: with U > 1 a lane issues U index loads, then U `visited`
// gathers, then the atomics, so U dependent load chains are in flight per lane instead of one (the kernel is bound by
// the latency of that chain, profiles/README.md).  U = 1 is the validated default; U = 4 is opt-in (bfs_push_unroll).
template <int W, int U> __global__ void __launch_bounds__(TPB) k_push(BfsDev d) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 0) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    const u64 *vis = d.visited + (int64_t)b * d.n * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t u = base + threadIdx.x;
        u64 f[W];
#pragma unroll
        for (int j = 0; j < W; j++) f[j] = 0ULL;
        if (u < d.n) ldw<W>(fr + u * W, f);
        u64 anyf = 0ULL;
#pragma unroll
        for (int j = 0; j < W; j++) anyf |= f[j];
        // every lane fetches its own row bounds up front (coalesced); they are broadcast below
        uint64_t my0 = 0, my1 = 0;
        if (anyf != 0ULL) {
            my0 = d.rowptr[u];
            my1 = d.rowptr[u + 1];
        }
        unsigned m = __ballot_sync(FULL, anyf != 0ULL);
        while (m) {
            int src_lane = __ffs(m) - 1;
            m &= m - 1;
            u64 fw[W];
#pragma unroll
            for (int j = 0; j < W; j++) fw[j] = __shfl_sync(FULL, f[j], src_lane);
            uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            if constexpr (U == 1) {
                for (uint64_t e = e0 + lane; e < e1; e += 32) {
                    uint32_t c = d.adj[e] >> 6;
                    VGA_COUNT(push_entries, 1);
                    if (c < (uint32_t)d.n) {
                        u64 vv[W];
                        ldw<W>(vis + (int64_t)c * W, vv);
#pragma unroll
                        for (int j = 0; j < W; j++) {
                            u64 add = fw[j] & ~vv[j];
                            if (add) {
                                atomicOr(&nx[(int64_t)c * W + j], add);
                                VGA_COUNT(push_atomics, 1);
                            }
                        }
                    }
                }
            } else {
                for (uint64_t e = e0 + lane; e < e1; e += 32 * U) {
                    uint32_t c[U];
                    u64 vv[U][W];
#pragma unroll
                    for (int i = 0; i < U; i++) {
                        const uint64_t ee = e + 32 * (uint64_t)i;
                        c[i] = ee < e1 ? d.adj[ee] >> 6 : 0xffffffffu;  // ghosts and the tail fail the c < n test
                    }
#pragma unroll
                    for (int i = 0; i < U; i++) {
#pragma unroll
                        for (int j = 0; j < W; j++) vv[i][j] = ~0ULL;
                        if (c[i] < (uint32_t)d.n) ldw<W>(vis + (int64_t)c[i] * W, vv[i]);
                    }
#pragma unroll
                    for (int i = 0; i < U; i++) {
#pragma unroll
                        for (int j = 0; j < W; j++) {
                            u64 add = fw[j] & ~vv[i][j];
                            if (add) atomicOr(&nx[(int64_t)c[i] * W + j], add);
                        }
                    }
                }
            }
        }
    }
}

// bottom-up step with early exit; `level` = level of the current frontier
template <int W> __global__ void __launch_bounds__(TPB) k_pull(BfsDev d, int level) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 1) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    const u64 *vis = d.visited + (int64_t)b * d.n * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    const uint8_t *lvl = d.lvl_in ? d.lvl_in + (int64_t)(b / d.group) * d.n : nullptr;
    u64 valid[W];
#pragma unroll
    for (int j = 0; j < W; j++) valid[j] = d.valid[(int64_t)b * W + j];
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t w = base + threadIdx.x;
        u64 need[W];
#pragma unroll
        for (int j = 0; j < W; j++) need[j] = 0ULL;
        u64 anyneed = 0ULL;
        if (w < d.n && !(lvl && (int)lvl[w] > level + 1)) {  // lvl > level+1: cannot be reached yet
            u64 vv[W];
            ldw<W>(vis + w * W, vv);
#pragma unroll
            for (int j = 0; j < W; j++) {
                need[j] = valid[j] & ~vv[j];
                anyneed |= need[j];
            }
        }
        uint64_t my0 = 0, my1 = 0;
        if (anyneed != 0ULL) {
            my0 = d.t_rowptr[w];
            my1 = d.t_rowptr[w + 1];
        }
        // Candidates of this warp are served by 4 groups of 8 lanes, each group scanning a different
        // vertex's in-row: 4 rows in flight per warp hide the dependent index->word latency, and a
        // row is left as soon as the in-neighbours seen so far cover the missing bits.
        const unsigned m = __ballot_sync(FULL, anyneed != 0ULL);
        const int ncand = __popc(m);
        const int grp = lane >> 3, gl = lane & 7;
        const unsigned gmask = 0xffu << (grp * 8);
        for (int r = 0; r * 4 < ncand; r++) {
            const int k = r * 4 + grp;
            const bool has = k < ncand;
            const int src_lane = has ? (int)__fns(m, 0, k + 1) : 0;
            u64 nd[W];
#pragma unroll
            for (int j = 0; j < W; j++) nd[j] = __shfl_sync(FULL, need[j], src_lane);
            const uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            if (has) {
                u64 acc[W];
#pragma unroll
                for (int j = 0; j < W; j++) acc[j] = 0ULL;
                for (uint64_t e = e0; e < e1; e += 16) {
                    const uint64_t ea = e + gl, eb = e + 8 + gl;
                    u64 g0[W], g1[W];
#pragma unroll
                    for (int j = 0; j < W; j++) g0[j] = g1[j] = 0ULL;
                    uint32_t ca = 0, cb = 0;
                    if (ea < e1) ca = d.t_col[ea];
                    if (eb < e1) cb = d.t_col[eb];
                    VGA_COUNT(pull_entries, (ea < e1) + (eb < e1));
                    if (ea < e1) ldw<W>(fr + (int64_t)ca * W, g0);
                    if (eb < e1) ldw<W>(fr + (int64_t)cb * W, g1);
                    bool done = true;
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        u64 gj = g0[j] | g1[j];
                        unsigned lo = __reduce_or_sync(gmask, (unsigned)gj);
                        unsigned hi = __reduce_or_sync(gmask, (unsigned)(gj >> 32));
                        acc[j] |= ((u64)hi << 32) | lo;
                        done = done && ((acc[j] & nd[j]) == nd[j]);
                    }
                    if (done) break;
                }
                if (gl == 0) {
                    const int64_t ww = (base + (threadIdx.x & ~31)) + src_lane;
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        u64 nw = acc[j] & nd[j];
                        if (nw) nx[ww * W + j] = nw;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---- pyramid push (bfs_push = 1, see pyramid.cuh) -----------------------------------------------------------

// top-down step over run-length out-rows: every frontier vertex ORs its words into the pyramid nodes that tile each run
// of its row (at most 2 per level); single vertices (level-0 nodes) go straight to `next`, filtered by `visited`.
// A plain read skips the atomic when the node already holds the bits.
template <int W> __global__ void __launch_bounds__(TPB) k_push_pyr(BfsDev d) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 2) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    const u64 *vis = d.visited + (int64_t)b * d.n * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    u64 *np = d.npyr + (int64_t)b * d.pyr_total * W;
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t u = base + threadIdx.x;
        u64 f[W];
#pragma unroll
        for (int j = 0; j < W; j++) f[j] = 0ULL;
        if (u < d.n) ldw<W>(fr + u * W, f);
        u64 anyf = 0ULL;
#pragma unroll
        for (int j = 0; j < W; j++) anyf |= f[j];
        uint64_t my0 = 0, my1 = 0;
        if (anyf != 0ULL) {
            my0 = d.f_runptr[u];
            my1 = d.f_runptr[u + 1];
        }
        unsigned m = __ballot_sync(FULL, anyf != 0ULL);
        while (m) {
            int src_lane = __ffs(m) - 1;
            m &= m - 1;
            u64 fw[W];
#pragma unroll
            for (int j = 0; j < W; j++) fw[j] = __shfl_sync(FULL, f[j], src_lane);
            uint64_t r0 = __shfl_sync(FULL, my0, src_lane), r1 = __shfl_sync(FULL, my1, src_lane);
            for (uint64_t r = r0 + lane; r < r1; r += 32) {
                const uint2 run = d.f_runs[r];
                VGA_COUNT(ppush_runs, 1);
                pyr_decompose(run.x, run.y, [&](int kk, uint32_t i) {
                    u64 *p = kk == 0 ? nx + (int64_t)i * W : np + (d.pyr_off[kk] + (int64_t)i) * W;
                    u64 cur[W];
                    VGA_COUNT(ppush_nodes, 1);
                    ldw<W>(kk == 0 ? vis + (int64_t)i * W : p, cur);  // leaf: reached sources; node: bits already there
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        u64 add = fw[j] & ~cur[j];
                        if (add) {
                            atomicOr(&p[j], add);
                            VGA_COUNT(ppush_atomics, 1);
                        }
                    }
                });
            }
        }
    }
}

// down pass of the pyramid push, three levels per launch, top chunk first: level k+3 .. k+1 -> level k (`next` for k = 0)
template <int W> __global__ void __launch_bounds__(TPB) k_pyr_down(BfsDev d, int k) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 2) return;
    u64 *np = d.npyr + (int64_t)b * d.pyr_total * W;
    u64 *dst = k == 0 ? d.next + (int64_t)b * d.n * W : np + d.pyr_off[k] * W;
    u64 *s1 = k + 1 < d.pyr_levels ? np + d.pyr_off[k + 1] * W : nullptr;
    u64 *s2 = k + 2 < d.pyr_levels ? np + d.pyr_off[k + 2] * W : nullptr;
    u64 *s3 = k + 3 < d.pyr_levels ? np + d.pyr_off[k + 3] * W : nullptr;
    const int64_t c0 = d.pyr_cnt[k];
    const int64_t c1 = s1 ? d.pyr_cnt[k + 1] : 0, c2 = s2 ? d.pyr_cnt[k + 2] : 0, c3 = s3 ? d.pyr_cnt[k + 3] : 0;
    const int64_t groups = (c0 + 7) / 8;
    for (int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x; t < groups; t += (int64_t)gridDim.x * TPB) {
        VGA_COUNT(pyr_down_groups, 1);
        pyr_down_group<W>(dst, c0, s1, c1, s2, c2, s3, c3, t);
    }
}

// ---- node-id lists (bfs_pyr_nodes = 1): the pyramid steps with the inner loops of the entry kernels ------------------
// A row's runs are replaced by the ids of the pyramid nodes that tile them, so the top-down and bottom-up steps need no
// per-run decomposition: they are k_push / k_pull over id lists, where an id < n addresses the vertex's own word and an
// id >= n the inner node id - n of the batch's pyramid.

// ids of the nodes of every run, written at the run's offset (exclusive scan of the per-run node counts)
__global__ void k_emit_nodes(const uint2 *runs, int64_t nruns, const u64 *node_off, uint32_t n, BfsDev d, uint32_t *out) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nruns) return;
    const uint2 run = runs[r];
    uint64_t o = node_off[r];
    pyr_decompose(run.x, run.y, [&](int k, uint32_t i) { out[o++] = k == 0 ? i : n + (uint32_t)d.pyr_off[k] + i; });
}
__global__ void k_run_costs(const uint2 *runs, int64_t nruns, u64 *cost) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nruns) cost[r] = (u64)pyr_cost(runs[r].x, runs[r].y);
}

// top-down step over node-id lists: k_push with the target word itself as the read-before-atomic filter (k_update drops
// words that only repeat reached sources, as for k_push_pyr)
template <int W> __global__ void __launch_bounds__(TPB) k_push_nodes(BfsDev d) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 2) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    u64 *np = d.npyr + (int64_t)b * d.pyr_total * W;
    const uint32_t n = (uint32_t)d.n;
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t u = base + threadIdx.x;
        u64 f[W];
#pragma unroll
        for (int j = 0; j < W; j++) f[j] = 0ULL;
        if (u < d.n) ldw<W>(fr + u * W, f);
        u64 anyf = 0ULL;
#pragma unroll
        for (int j = 0; j < W; j++) anyf |= f[j];
        uint64_t my0 = 0, my1 = 0;
        if (anyf != 0ULL) {
            my0 = d.f_costptr[u];
            my1 = d.f_costptr[u + 1];
        }
        unsigned m = __ballot_sync(FULL, anyf != 0ULL);
        while (m) {
            int src_lane = __ffs(m) - 1;
            m &= m - 1;
            u64 fw[W];
#pragma unroll
            for (int j = 0; j < W; j++) fw[j] = __shfl_sync(FULL, f[j], src_lane);
            uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            for (uint64_t e = e0 + lane; e < e1; e += 32) {
                const uint32_t c = d.f_nodes[e];
                VGA_COUNT(npush_nodes, 1);
                u64 *p = c < n ? nx + (int64_t)c * W : np + (int64_t)(c - n) * W;
                u64 cur[W];
                ldw<W>(p, cur);
#pragma unroll
                for (int j = 0; j < W; j++) {
                    u64 add = fw[j] & ~cur[j];
                    if (add) {
                        atomicOr(&p[j], add);
                        VGA_COUNT(npush_atomics, 1);
                    }
                }
            }
        }
    }
}

// bottom-up step over node-id lists: k_pull with the frontier words of leaves and the frontier pyramid's inner nodes
template <int W> __global__ void __launch_bounds__(TPB) k_pull_nodes(BfsDev d, int level) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 1) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    const u64 *vis = d.visited + (int64_t)b * d.n * W;
    const u64 *pyr = d.pyr + (int64_t)b * d.pyr_total * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    const uint8_t *lvl = d.lvl_in ? d.lvl_in + (int64_t)(b / d.group) * d.n : nullptr;
    const uint32_t n = (uint32_t)d.n;
    u64 valid[W];
#pragma unroll
    for (int j = 0; j < W; j++) valid[j] = d.valid[(int64_t)b * W + j];
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t w = base + threadIdx.x;
        u64 need[W];
#pragma unroll
        for (int j = 0; j < W; j++) need[j] = 0ULL;
        u64 anyneed = 0ULL;
        if (w < d.n && !(lvl && (int)lvl[w] > level + 1)) {
            u64 vv[W];
            ldw<W>(vis + w * W, vv);
#pragma unroll
            for (int j = 0; j < W; j++) {
                need[j] = valid[j] & ~vv[j];
                anyneed |= need[j];
            }
        }
        uint64_t my0 = 0, my1 = 0;
        if (anyneed != 0ULL) {
            my0 = d.t_costptr[w];
            my1 = d.t_costptr[w + 1];
        }
        const unsigned m = __ballot_sync(FULL, anyneed != 0ULL);
        const int ncand = __popc(m);
        const int grp = lane >> 3, gl = lane & 7;
        const unsigned gmask = 0xffu << (grp * 8);
        for (int r = 0; r * 4 < ncand; r++) {
            const int k = r * 4 + grp;
            const bool has = k < ncand;
            const int src_lane = has ? (int)__fns(m, 0, k + 1) : 0;
            u64 nd[W];
#pragma unroll
            for (int j = 0; j < W; j++) nd[j] = __shfl_sync(FULL, need[j], src_lane);
            const uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            if (has) {
                u64 acc[W];
#pragma unroll
                for (int j = 0; j < W; j++) acc[j] = 0ULL;
                for (uint64_t e = e0; e < e1; e += 16) {
                    const uint64_t ea = e + gl, eb = e + 8 + gl;
                    u64 g0[W], g1[W];
#pragma unroll
                    for (int j = 0; j < W; j++) g0[j] = g1[j] = 0ULL;
                    uint32_t ca = 0, cb = 0;
                    if (ea < e1) ca = d.t_nodes[ea];
                    if (eb < e1) cb = d.t_nodes[eb];
                    VGA_COUNT(npull_nodes, (ea < e1) + (eb < e1));
                    if (ea < e1) ldw<W>(ca < n ? fr + (int64_t)ca * W : pyr + (int64_t)(ca - n) * W, g0);
                    if (eb < e1) ldw<W>(cb < n ? fr + (int64_t)cb * W : pyr + (int64_t)(cb - n) * W, g1);
                    bool done = true;
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        u64 gj = g0[j] | g1[j];
                        unsigned lo = __reduce_or_sync(gmask, (unsigned)gj);
                        unsigned hi = __reduce_or_sync(gmask, (unsigned)(gj >> 32));
                        acc[j] |= ((u64)hi << 32) | lo;
                        done = done && ((acc[j] & nd[j]) == nd[j]);
                    }
                    if (done) break;
                }
                if (gl == 0) {
                    const int64_t ww = (base + (threadIdx.x & ~31)) + src_lane;
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        u64 nw = acc[j] & nd[j];
                        if (nw) nx[ww * W + j] = nw;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// ---- pyramid pull (bfs_pull = 1, see pyramid.cuh) -----------------------------------------------------------

// Three pyramid levels per launch for every batch whose next step is a pull: level k (the frontier for k = 0) ->
// levels k+1 .. k+3.  One work item per aligned group of 8 level-k nodes.
template <int W> __global__ void __launch_bounds__(TPB) k_pyr_build(BfsDev d, int k) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 1) return;
    u64 *pyr = d.pyr + (int64_t)b * d.pyr_total * W;
    const u64 *src = k == 0 ? d.frontier + (int64_t)b * d.n * W : pyr + d.pyr_off[k] * W;
    u64 *d1 = k + 1 < d.pyr_levels ? pyr + d.pyr_off[k + 1] * W : nullptr;
    u64 *d2 = k + 2 < d.pyr_levels ? pyr + d.pyr_off[k + 2] * W : nullptr;
    u64 *d3 = k + 3 < d.pyr_levels ? pyr + d.pyr_off[k + 3] * W : nullptr;
    const int64_t c0 = d.pyr_cnt[k];
    const int64_t c1 = d1 ? d.pyr_cnt[k + 1] : 0, c2 = d2 ? d.pyr_cnt[k + 2] : 0, c3 = d3 ? d.pyr_cnt[k + 3] : 0;
    const int64_t groups = (c0 + 7) / 8;
    for (int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x; t < groups; t += (int64_t)gridDim.x * TPB) {
        VGA_COUNT(pyr_build_groups, 1);
        pyr_build_group<W>(src, c0, d1, c1, d2, c2, d3, c3, t);
    }
}

// bottom-up step over run-length in-rows: candidates as in k_pull (4 groups of 8 lanes per warp, one vertex per
// group); every lane of a group takes one run of the in-row per round and answers it with a range-OR query over the
// pyramid (at most 2 nodes per level), the group ORs its 8 answers and leaves the row once the missing bits are found.
template <int W> __global__ void __launch_bounds__(TPB) k_pull_pyr(BfsDev d, int level) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 1) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    const u64 *vis = d.visited + (int64_t)b * d.n * W;
    const u64 *pyr = d.pyr + (int64_t)b * d.pyr_total * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    const uint8_t *lvl = d.lvl_in ? d.lvl_in + (int64_t)(b / d.group) * d.n : nullptr;
    u64 valid[W];
#pragma unroll
    for (int j = 0; j < W; j++) valid[j] = d.valid[(int64_t)b * W + j];
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t w = base + threadIdx.x;
        u64 need[W];
#pragma unroll
        for (int j = 0; j < W; j++) need[j] = 0ULL;
        u64 anyneed = 0ULL;
        if (w < d.n && !(lvl && (int)lvl[w] > level + 1)) {
            u64 vv[W];
            ldw<W>(vis + w * W, vv);
#pragma unroll
            for (int j = 0; j < W; j++) {
                need[j] = valid[j] & ~vv[j];
                anyneed |= need[j];
            }
        }
        uint64_t my0 = 0, my1 = 0;
        if (anyneed != 0ULL) {
            my0 = d.t_runptr[w];
            my1 = d.t_runptr[w + 1];
        }
        const unsigned m = __ballot_sync(FULL, anyneed != 0ULL);
        const int ncand = __popc(m);
        const int grp = lane >> 3, gl = lane & 7;
        const unsigned gmask = 0xffu << (grp * 8);
        for (int r = 0; r * 4 < ncand; r++) {
            const int k = r * 4 + grp;
            const bool has = k < ncand;
            const int src_lane = has ? (int)__fns(m, 0, k + 1) : 0;
            u64 nd[W];
#pragma unroll
            for (int j = 0; j < W; j++) nd[j] = __shfl_sync(FULL, need[j], src_lane);
            const uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            if (has) {
                u64 acc[W];
#pragma unroll
                for (int j = 0; j < W; j++) acc[j] = 0ULL;
                for (uint64_t e = e0; e < e1; e += 8) {
                    u64 g[W];
#pragma unroll
                    for (int j = 0; j < W; j++) g[j] = 0ULL;
                    if (e + gl < e1) {
                        const uint2 run = d.t_runs[e + gl];
                        VGA_COUNT(ppull_runs, 1);
                        pyr_decompose(run.x, run.y, [&](int kk, uint32_t i) {
                            VGA_COUNT(ppull_nodes, 1);
                            const u64 *p = kk == 0 ? fr + (int64_t)i * W : pyr + (d.pyr_off[kk] + (int64_t)i) * W;
                            u64 t[W];
                            ldw<W>(p, t);
#pragma unroll
                            for (int j = 0; j < W; j++) g[j] |= t[j];
                        });
                    }
                    bool done = true;
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        unsigned lo = __reduce_or_sync(gmask, (unsigned)g[j]);
                        unsigned hi = __reduce_or_sync(gmask, (unsigned)(g[j] >> 32));
                        acc[j] |= ((u64)hi << 32) | lo;
                        done = done && ((acc[j] & nd[j]) == nd[j]);
                    }
                    if (done) break;
                }
                if (gl == 0) {
                    const int64_t ww = (base + (threadIdx.x & ~31)) + src_lane;
#pragma unroll
                    for (int j = 0; j < W; j++) {
                        u64 nw = acc[j] & nd[j];
                        if (nw) nx[ww * W + j] = nw;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// fold next into visited/frontier, count new vertices per source, gather direction statistics;
// `level_next` = level of the vertices being added
template <int W>
__global__ void __launch_bounds__(TPB) k_update(BfsDev d, int32_t *counts /*[batches*W][64] of level_next*/, int level_next) {
    const int b = blockIdx.y;
    if (!d.active[b]) return;
    __shared__ int s_cnt[W * 64];
    __shared__ u64 s_stat[6];
    for (int i = threadIdx.x; i < W * 64; i += TPB) s_cnt[i] = 0;
    if (threadIdx.x < 6) s_stat[threadIdx.x] = 0ULL;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    u64 *fr = d.frontier + (int64_t)b * d.n * W;
    u64 *vis = d.visited + (int64_t)b * d.n * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    const uint8_t *lvl = d.lvl_in ? d.lvl_in + (int64_t)(b / d.group) * d.n : nullptr;
    uint8_t *lout = d.lvl_out ? d.lvl_out + (int64_t)b * 64 * d.n : nullptr;
    u64 valid[W];
#pragma unroll
    for (int j = 0; j < W; j++) valid[j] = d.valid[(int64_t)b * W + j];
    int cnt[W][2];  // lane l counts source bits l and l+32 of each word
#pragma unroll
    for (int j = 0; j < W; j++) cnt[j][0] = cnt[j][1] = 0;
    u64 f_edges = 0, u_edges = 0, n_new = 0, n_open = 0, f_runs = 0, f_nrun = 0;
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t v = base + threadIdx.x;
        u64 nw[W];
#pragma unroll
        for (int j = 0; j < W; j++) nw[j] = 0ULL;
        if (v < d.n) {
            VGA_COUNT(update_words, 1);
            u64 vv[W], xx[W];
            ldw<W>(vis + v * W, vv);
            ldw<W>(nx + v * W, xx);
            u64 anynew = 0ULL, anyneed = 0ULL;
#pragma unroll
            for (int j = 0; j < W; j++) {
                nw[j] = xx[j] & ~vv[j];
                vv[j] |= nw[j];
                anynew |= nw[j];
                anyneed |= valid[j] & ~vv[j];
            }
            if (anynew) {
                stw<W>(vis + v * W, vv);
                u64 zero[W];
#pragma unroll
                for (int j = 0; j < W; j++) zero[j] = 0ULL;
                stw<W>(nx + v * W, zero);
                f_edges += d.rowptr[v + 1] - d.rowptr[v];
                if (d.f_costptr) {
                    f_runs += d.f_costptr[v + 1] - d.f_costptr[v];
                    f_nrun += d.f_runptr[v + 1] - d.f_runptr[v];
                }
                n_new += 1;
                if (lout) {
                    // coarse pass (W == 1): bit j of coarse batch b is group b*64+j
                    u64 bits = nw[0];
                    while (bits) {
                        int j = __ffsll((long long)bits) - 1;
                        bits &= bits - 1;
                        lout[(int64_t)j * d.n + v] = (uint8_t)min(level_next, 254);
                    }
                }
            }
            else if (d.npyr) {
                // the pyramid push does not filter by `visited`: drop words that only repeat reached sources
                u64 anyx = 0ULL;
#pragma unroll
                for (int j = 0; j < W; j++) anyx |= xx[j];
                if (anyx) {
                    u64 zero[W];
#pragma unroll
                    for (int j = 0; j < W; j++) zero[j] = 0ULL;
                    stw<W>(nx + v * W, zero);
                }
            }
            if (d.noexpand && d.noexpand[v]) {
                // reached and counted, but not expanded (vgavisualglobal.cpp:108-110)
                u64 zero[W];
#pragma unroll
                for (int j = 0; j < W; j++) zero[j] = 0ULL;
                stw<W>(fr + v * W, zero);
            } else {
                stw<W>(fr + v * W, nw);
            }
            if (anyneed != 0ULL) n_open += 1;  // still unreached by some source of the batch
            // in-edges (or, for the pyramid pull, pyramid loads) the next pull step would have to consider
            const uint64_t *cost = d.t_costptr ? d.t_costptr : d.t_rowptr;
            if (cost && anyneed != 0ULL && (!lvl || (int)lvl[v] <= level_next + 1)) u_edges += cost[v + 1] - cost[v];
        }
#pragma unroll
        for (int j = 0; j < W; j++) {
            // new vertices per source: column sums of the warp's 32 x 64 bit matrix.  Few non-empty columns (late
            // levels): one ballot per column; many (a coherent batch reaching a vertex with most of its sources at
            // once): transpose the two 32 x 32 halves and popc.
            const unsigned lo = (unsigned)nw[j], hi = (unsigned)(nw[j] >> 32);
            unsigned lo_any = __reduce_or_sync(FULL, lo);
            unsigned hi_any = __reduce_or_sync(FULL, hi);
            if (__popc(lo_any) >= DENSE_COLUMNS) {
                cnt[j][0] += __popc(warp_transpose32(lo, lane));
            } else {
                while (lo_any) {
                    int bit = __ffs(lo_any) - 1;
                    lo_any &= lo_any - 1;
                    int c = __popc(__ballot_sync(FULL, (lo >> bit) & 1u));
                    if (lane == bit) cnt[j][0] += c;
                }
            }
            if (__popc(hi_any) >= DENSE_COLUMNS) {
                cnt[j][1] += __popc(warp_transpose32(hi, lane));
            } else {
                while (hi_any) {
                    int bit = __ffs(hi_any) - 1;
                    hi_any &= hi_any - 1;
                    int c = __popc(__ballot_sync(FULL, (hi >> bit) & 1u));
                    if (lane == bit) cnt[j][1] += c;
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < W; j++) {
        if (cnt[j][0]) atomicAdd(&s_cnt[j * 64 + lane], cnt[j][0]);
        if (cnt[j][1]) atomicAdd(&s_cnt[j * 64 + lane + 32], cnt[j][1]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        f_edges += __shfl_down_sync(FULL, f_edges, o);
        u_edges += __shfl_down_sync(FULL, u_edges, o);
        n_new += __shfl_down_sync(FULL, n_new, o);
        n_open += __shfl_down_sync(FULL, n_open, o);
        f_runs += __shfl_down_sync(FULL, f_runs, o);
        f_nrun += __shfl_down_sync(FULL, f_nrun, o);
    }
    if (lane == 0) {
        if (f_edges) atomicAdd(&s_stat[0], f_edges);
        if (u_edges) atomicAdd(&s_stat[1], u_edges);
        if (n_new) atomicAdd(&s_stat[2], n_new);
        if (n_open) atomicAdd(&s_stat[3], n_open);
        if (f_runs) atomicAdd(&s_stat[4], f_runs);
        if (f_nrun) atomicAdd(&s_stat[5], f_nrun);
    }
    __syncthreads();
    if (counts)
        for (int i = threadIdx.x; i < W * 64; i += TPB)
            if (s_cnt[i]) atomicAdd(&counts[(int64_t)b * W * 64 + i], s_cnt[i]);
    if (threadIdx.x < 6 && s_stat[threadIdx.x]) atomicAdd(&d.stats[b * NSTAT + threadIdx.x], s_stat[threadIdx.x]);
}

// per batch: retire empty batches, choose the next step's direction, reset statistics
__global__ void k_decide(BfsDev d, int nb, int bfs_mode, int64_t alpha, int64_t beta, u64 *work /*[4]*/) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    if (!d.active[b]) return;
    u64 fe = d.stats[b * NSTAT + 0], ue = d.stats[b * NSTAT + 1], nn = d.stats[b * NSTAT + 2], open = d.stats[b * NSTAT + 3];
    const u64 frc = d.stats[b * NSTAT + 4], fnr = d.stats[b * NSTAT + 5];
    for (int i = 0; i < 6; i++) d.stats[b * NSTAT + i] = 0;
    // Retire the batch when nothing new was reached -- or when every vertex has been reached by every source of the
    // batch: expanding the last frontier could not find anything (on connected plans that last, useless expansion of
    // the largest frontier was 10-20 % of the top-down work).
    if (nn != 0) atomicOr(d.any, 2);  // bit 1: this level added vertices (the histogram has one more level)
    if (nn == 0 || open == 0) {
        d.active[b] = 0;
        return;
    }
    atomicOr(d.any, 1);  // bit 0: some batch goes on
    // step of the next level: 0 = push over adjacency entries, 1 = pull, 2 = push as range-OR updates over the runs of the
    // out-rows (its cost: the pyramid nodes the frontier rows touch plus the down pass over ~2n nodes)
    int m = 0;
    if (bfs_mode == 1)
        m = 1;
    else if (bfs_mode == 2)
        m = (fe * (u64)alpha > ue * (u64)beta) ? 1 : 0;
    const u64 pyr_cost_w = (frc + 2 * (u64)d.n) * (u64)d.pyr_weight;  // in hundredths of an adjacency entry
    if (m == 0 && d.npyr && (d.push_force || pyr_cost_w < fe * 100ULL)) m = 2;
    if (m == 1 && d.npyr && bfs_mode == 2 && pyr_cost_w * (u64)alpha < ue * (u64)beta * 100ULL) m = 2;
    d.mode[b] = m;
    if (work) {
        // work[0]: adjacency entries of the vertices that expand next (sum of deg over U_l),
        // work[1]: vertices newly reached (|U_l| summed), work[2]: in-edges offered to the pull step
        atomicAdd(&work[0], fe);
        atomicAdd(&work[1], nn);
        if (m == 1) atomicAdd(&work[2], ue);
        atomicAdd(&work[3], fnr);  // runs of the out-rows of the vertices that expand next (run-length model)
    }
}

// ---- transpose (in-edge lists of filled vertices), needed by the pull step --------------------
__global__ void k_indeg(const uint32_t *adj, uint64_t n_entries, uint32_t n, u64 *indeg) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    uint32_t c = adj[e] >> 6;
    if (c < n) atomicAdd(&indeg[c], 1ULL);
}
__global__ void k_scatter_t(const uint64_t *rowptr, const uint32_t *adj, int64_t n, u64 *cursor, uint32_t *t_col) {
    int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;  // one warp per source row
    int lane = threadIdx.x & 31;
    if (u >= n) return;
    for (uint64_t e = rowptr[u] + lane; e < rowptr[u + 1]; e += 32) {
        uint32_t c = adj[e] >> 6;
        if (c < (uint32_t)n) {
            u64 p = atomicAdd(&cursor[c], 1ULL);
            t_col[p] = (uint32_t)u;
        }
    }
}

// bit k of gc[v] = the k-th of the 8 grid neighbours (E, NE, N, NW, W, SW, S, SE) is in v's row;
// rows are sorted by column, so each lookup is a binary search.  Used to grow source batches that
// do not cross walls.
__global__ void k_neighbour_bits(int64_t n, const uint64_t *rowptr, const uint32_t *adj, const int32_t *refs,
                                 const int32_t *ord_of, int cols, int rows, uint8_t *gc) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    uint32_t r = (uint32_t)refs[v];
    int x = (int)(r >> 16), y = (int)(r & 0xffff);
    uint64_t e0 = rowptr[v], e1 = rowptr[v + 1];
    uint8_t out = 0;
    for (int k = 0; k < 8; k++) {
        int nx = x + dx[k], ny = y + dy[k];
        if (nx < 0 || nx >= cols || ny < 0 || ny >= rows) continue;
        int32_t w = ord_of[(int64_t)nx * rows + ny];
        if (w < 0) continue;
        uint64_t lo = e0, hi = e1;  // first entry with col >= w
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if ((adj[mid] >> 6) < (uint32_t)w) lo = mid + 1; else hi = mid;
        }
        if (lo < e1 && (adj[lo] >> 6) == (uint32_t)w) out |= (uint8_t)(1 << k);
    }
    gc[v] = out;
}

// ---- run-length rows (bfs_pull / bfs_push = 1): one warp per vertex over a SORTED row --------------------------------
// entry e of row v holds ordinal arr[e] >> shift; ordinals >= limit (ghost columns, which sort last) are not part of
// any run.  A run starts at a valid entry that is the first of the row or does not continue the previous ordinal.
__global__ void k_count_runs(const uint64_t *rowptr, const uint32_t *arr, int shift, uint32_t limit, int64_t n, u64 *count,
                             uint32_t *nvalid) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint64_t e0 = rowptr[v], e1 = rowptr[v + 1];
    unsigned c = 0, nv = 0;
    for (uint64_t e = e0 + lane; e < e1; e += 32) {
        const uint32_t x = arr[e] >> shift;
        if (x >= limit) continue;
        nv++;
        c += (e == e0 || x != (arr[e - 1] >> shift) + 1u) ? 1u : 0u;
    }
    c = __reduce_add_sync(FULL, c);
    nv = __reduce_add_sync(FULL, nv);
    if (lane == 0) {
        count[v] = c;
        nvalid[v] = nv;
    }
}
// first_off[r] = offset (within its row) of the entry that starts run r
__global__ void k_mark_runs(const uint64_t *rowptr, const uint32_t *arr, int shift, const uint32_t *nvalid, int64_t n,
                            const uint64_t *runptr, uint32_t *first_off) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint64_t e0 = rowptr[v], e1 = e0 + nvalid[v];
    uint64_t out = runptr[v];
    for (uint64_t eb = e0; eb < e1; eb += 32) {
        const uint64_t e = eb + lane;
        const bool start = e < e1 && (e == e0 || (arr[e] >> shift) != (arr[e - 1] >> shift) + 1u);
        const unsigned mask = __ballot_sync(FULL, start);
        if (start) first_off[out + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(e - e0);
        out += __popc(mask);
    }
}
// runs[r] = (first ordinal, length); cost[v] = pyramid nodes a full pass over the row of v touches
__global__ void k_emit_runs(const uint64_t *rowptr, const uint32_t *arr, int shift, const uint32_t *nvalid, int64_t n,
                            const uint64_t *runptr, const uint32_t *first_off, uint2 *runs, u64 *cost) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint64_t e0 = rowptr[v];
    const uint32_t nv = nvalid[v];
    const uint64_t r0 = runptr[v], r1 = runptr[v + 1];
    unsigned c = 0;
    for (uint64_t r = r0 + lane; r < r1; r += 32) {
        const uint32_t off = first_off[r];
        const uint32_t nxt = r + 1 < r1 ? first_off[r + 1] : nv;
        const uint2 run = make_uint2(arr[e0 + off] >> shift, nxt - off);
        runs[r] = run;
        c += (unsigned)pyr_cost(run.x, run.y);
    }
    c = __reduce_add_sync(FULL, c);
    if (lane == 0) cost[v] = c;
}

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

inline uint32_t morton2(uint32_t x, uint32_t y) {
    auto spread = [](uint32_t v) {
        v &= 0xffff;
        v = (v | (v << 8)) & 0x00ff00ff;
        v = (v | (v << 4)) & 0x0f0f0f0f;
        v = (v | (v << 2)) & 0x33333333;
        v = (v | (v << 1)) & 0x55555555;
        return v;
    };
    return spread(x) | (spread(y) << 1);
}

// Level loop over a set of batches of W words.  nlev = number of levels with data (>= 1).
template <int W>
int run_levels(vga_ctx *ctx, BfsDev &d, int64_t nb, int radius, int bfs_mode, DevBuf<int32_t> *counts, int *lcap,
               int64_t counts_stride, u64 *work, StageTimer &mt, int *nlev_out) {
    cudaStream_t st = ctx->stream;
    Timing &tm = ctx->timing;
    const unsigned xblocks = (unsigned)std::min<int64_t>((d.n + TPB - 1) / TPB, 4096);
    int level = 0, nlev = 1;
    while (radius == -1 || level < radius) {
        dim3 grid(xblocks, (unsigned)nb);
        if (ctx->opt.bfs_push_unroll == 4)
            k_push<W, 4><<<grid, TPB, 0, st>>>(d);
        else
            k_push<W, 1><<<grid, TPB, 0, st>>>(d);
        tm.launches++;
        tm.main_launches++;
        if (d.npyr) {
            if (d.f_nodes)
                k_push_nodes<W><<<grid, TPB, 0, st>>>(d);
            else
                k_push_pyr<W><<<grid, TPB, 0, st>>>(d);
            tm.launches++;
            tm.main_launches++;
            int kmax = 0;
            while (kmax + 4 < d.pyr_levels) kmax += 3;
            for (int k = kmax; k >= 0; k -= 3) {
                const int64_t groups = (d.pyr_cnt[k] + 7) / 8;
                dim3 pgrid((unsigned)std::min<int64_t>((groups + TPB - 1) / TPB, 4096), (unsigned)nb);
                k_pyr_down<W><<<pgrid, TPB, 0, st>>>(d, k);
                tm.launches++;
                tm.main_launches++;
            }
        }
        if (bfs_mode != 0 && level > 0) {
            if (d.pyr) {
                for (int k = 0; k + 1 < d.pyr_levels; k += 3) {
                    const int64_t groups = (d.pyr_cnt[k] + 7) / 8;
                    dim3 pgrid((unsigned)std::min<int64_t>((groups + TPB - 1) / TPB, 4096), (unsigned)nb);
                    k_pyr_build<W><<<pgrid, TPB, 0, st>>>(d, k);
                    tm.launches++;
                    tm.main_launches++;
                }
                if (d.t_nodes)
                    k_pull_nodes<W><<<grid, TPB, 0, st>>>(d, level);
                else
                    k_pull_pyr<W><<<grid, TPB, 0, st>>>(d, level);
            } else {
                k_pull<W><<<grid, TPB, 0, st>>>(d, level);
            }
            tm.launches++;
            tm.main_launches++;
        }
        if (counts && level + 1 >= *lcap) {
            mt.stop();
            DevBuf<int32_t> bigger;
            VGA_TRY(bigger.alloc_zero((size_t)(*lcap) * 2 * counts_stride, st));
            VGA_CUDA(cudaMemcpyAsync(bigger.p, counts->p, sizeof(int32_t) * (size_t)(*lcap) * counts_stride,
                                     cudaMemcpyDeviceToDevice, st));
            VGA_CUDA(cudaStreamSynchronize(st));
            *counts = std::move(bigger);
            *lcap *= 2;
            mt.start();
        }
        VGA_CUDA(cudaMemsetAsync(d.any, 0, sizeof(int), st));
        k_update<W><<<grid, TPB, 0, st>>>(d, counts ? counts->p + (size_t)(level + 1) * counts_stride : nullptr, level + 1);
        k_decide<<<blocks_for(nb, 128), 128, 0, st>>>(d, (int)nb, bfs_mode, ctx->opt.pull_alpha, ctx->opt.pull_beta, work);
        tm.launches += 2;
        tm.main_launches += 2;
        int h_any = 0;
        VGA_CUDA(cudaMemcpyAsync(&h_any, d.any, sizeof(int), cudaMemcpyDeviceToHost, st));
        VGA_CUDA(cudaStreamSynchronize(st));
        if (h_any & 2) nlev = level + 2;
        if (!(h_any & 1)) break;
        level++;
    }
    *nlev_out = nlev;
    return VGA_OK;
}

// Source order for batching (see header comment).  order[i] = ordinal of the i-th source.
int source_order(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, std::vector<int32_t> &order) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n, nsrc = src_end - src_begin;
    order.resize((size_t)nsrc);
    std::iota(order.begin(), order.end(), (int32_t)src_begin);
    if (ctx->opt.bfs_order == 0 || (int64_t)g->h_refs.size() < n) return VGA_OK;
    // key = Morton code of the 8x8 tile, then x and y inside the tile (30 bits), index in the low word; a stable LSD radix
    // sort over the key bytes replaces std::sort (1.8 ms -> 0.3 ms for the 65,536 sources of C2, once per call)
    std::vector<uint64_t> key((size_t)nsrc), tmp((size_t)nsrc);
    for (int64_t i = 0; i < nsrc; i++) {
        uint32_t r = (uint32_t)g->h_refs[(size_t)(src_begin + i)];
        uint32_t x = r >> 16, y = r & 0xffff;
        key[(size_t)i] = ((uint64_t)morton2(x >> 3, y >> 3) << 38) | ((uint64_t)(x & 7) << 35) | ((uint64_t)(y & 7) << 32) |
                         (uint64_t)(uint32_t)i;
    }
    for (int shift = 32; shift < 64; shift += 8) {
        size_t cnt[257] = {0};
        for (uint64_t k : key) cnt[((k >> shift) & 0xff) + 1]++;
        for (int b = 0; b < 256; b++) cnt[b + 1] += cnt[b];
        for (uint64_t k : key) tmp[cnt[(k >> shift) & 0xff]++] = k;
        key.swap(tmp);
    }
    for (int64_t i = 0; i < nsrc; i++) order[(size_t)i] = (int32_t)(src_begin + (int64_t)(key[(size_t)i] & 0xffffffffu));
    if (ctx->opt.bfs_order < 2) return VGA_OK;
    // Wall-respecting clusters: a raw 8x8 tile often straddles a wall, which puts cells of two rooms
    // (very different level structure) into one batch.  Grow each cluster as a flood of up to 64 cells
    // over the direct-neighbour links (the 8 grid neighbours that are in the cell's row, i.e. actually
    // visible), seeded in Morton order.
    int maxx = 0, maxy = 0;
    for (int64_t v = 0; v < n; v++) {
        uint32_t r = (uint32_t)g->h_refs[(size_t)v];
        maxx = std::max(maxx, (int)(r >> 16));
        maxy = std::max(maxy, (int)(r & 0xffff));
    }
    const int64_t cols = maxx + 1, rows = maxy + 1;
    std::vector<int32_t> ord_of((size_t)(cols * rows), -1);
    for (int64_t v = 0; v < n; v++) {
        uint32_t r = (uint32_t)g->h_refs[(size_t)v];
        ord_of[(size_t)((int64_t)(r >> 16) * rows + (r & 0xffff))] = (int32_t)v;
    }
    std::vector<uint8_t> gc((size_t)n);
    {
        DevBuf<int32_t> d_refs, d_ord;
        DevBuf<uint8_t> d_gc;
        VGA_TRY(d_refs.alloc((size_t)n));
        VGA_TRY(d_ord.alloc((size_t)(cols * rows)));
        VGA_TRY(d_gc.alloc((size_t)n));
        VGA_CUDA(cudaMemcpyAsync(d_refs.p, g->h_refs.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
        VGA_CUDA(cudaMemcpyAsync(d_ord.p, ord_of.data(), sizeof(int32_t) * cols * rows, cudaMemcpyHostToDevice, st));
        k_neighbour_bits<<<blocks_for(n, 256), 256, 0, st>>>(n, g->rowptr.p, g->adj.p, d_refs.p, d_ord.p, (int)cols, (int)rows,
                                                            d_gc.p);
        ctx->timing.launches++;
        VGA_CUDA(cudaMemcpyAsync(gc.data(), d_gc.p, (size_t)n, cudaMemcpyDeviceToHost, st));
        VGA_CUDA(cudaStreamSynchronize(st));
    }
    static const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    static const int dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    std::vector<uint8_t> taken((size_t)n, 0);
    std::vector<int32_t> clustered, queue;
    clustered.reserve((size_t)nsrc);
    for (int64_t i = 0; i < nsrc; i++) {
        const int32_t seed = order[(size_t)i];
        if (taken[(size_t)seed]) continue;
        queue.clear();
        queue.push_back(seed);
        taken[(size_t)seed] = 1;
        size_t head = 0;
        int count = 0;
        while (head < queue.size() && count < 64) {
            const int32_t v = queue[head++];
            clustered.push_back(v);
            count++;
            const uint32_t r = (uint32_t)g->h_refs[(size_t)v];
            const int x = (int)(r >> 16), y = (int)(r & 0xffff);
            for (int k = 0; k < 8; k++) {
                if (!(gc[(size_t)v] & (1 << k))) continue;
                const int nx = x + dx[k], ny = y + dy[k];
                if (nx < 0 || nx >= cols || ny < 0 || ny >= rows) continue;
                const int32_t w = ord_of[(size_t)((int64_t)nx * rows + ny)];
                if (w < (int32_t)src_begin || w >= (int32_t)src_end || taken[(size_t)w]) continue;
                taken[(size_t)w] = 1;
                queue.push_back(w);
            }
        }
        for (; head < queue.size(); head++) taken[(size_t)queue[head]] = 0;  // not placed: free again
    }
    if ((int64_t)clustered.size() == nsrc) order.swap(clustered);
    return VGA_OK;
}

template <int W>
int run_global_w(vga_ctx *ctx, vga_graph *g, int radius, int64_t src_begin, int64_t src_end, int32_t *total_nodes,
                 int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    const int64_t nsrc = src_end - src_begin;
    Timing &tm = ctx->timing;
    const int bfs_mode = (int)ctx->opt.bfs_mode;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);
    StageTimer dt(ctx, 4, &tm.d2h_ms);

    const bool dbg = std::getenv("VGA_DEBUG_TIMING") != nullptr;
    auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double w0 = wall();
    kt.start();
    if (bfs_mode != 0) VGA_TRY(ensure_transpose(ctx, g));
    PyrLayout pl = pyr_layout(n);
    const bool pyr_pull = bfs_mode != 0 && ctx->opt.bfs_pull == 1 && pl.levels <= PYR_LEVELS_DEV && n > 1;
    if (pyr_pull) VGA_TRY(ensure_runs(ctx, g));
    const bool pyr_push = ctx->opt.bfs_push >= 1 && pl.levels <= PYR_LEVELS_DEV && n > 1;
    if (pyr_push) VGA_TRY(ensure_fwd_runs(ctx, g));
    const bool node_lists = ctx->opt.bfs_pyr_nodes == 1 && (pyr_push || pyr_pull);
    if (node_lists) VGA_TRY(ensure_node_lists(ctx, g, pyr_push, pyr_pull));
    kt.stop();
    const double w1 = wall();

    std::vector<int32_t> order;
    VGA_TRY(source_order(ctx, g, src_begin, src_end, order));
    const double w2 = wall();
    DevBuf<int32_t> d_order;
    VGA_TRY(d_order.alloc((size_t)nsrc));
    VGA_CUDA(cudaMemcpyAsync(d_order.p, order.data(), sizeof(int32_t) * nsrc, cudaMemcpyHostToDevice, st));

    const int64_t nwords = (nsrc + 63) / 64;      // 64-source words
    const int64_t nbatch = (nwords + W - 1) / W;  // batches of W words
    // coarse groups: `group` batches each (bfs_group is given in 64-source words)
    const int group = (int)std::max<int64_t>(1, ctx->opt.bfs_group / W);
    const bool coarse = bfs_mode != 0 && ctx->opt.bfs_coarse != 0;
    // chunk size: 3*W words per (batch, vertex) + the coarse pass; below ~40% of free memory and 48 GB
    size_t free_b = 0, total_b = 0;
    VGA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    int64_t budget = (int64_t)std::min<size_t>((size_t)(free_b * 0.4), (size_t)48 << 30);
    int64_t chunk = ctx->opt.bfs_chunk > 0 ? std::max<int64_t>(1, ctx->opt.bfs_chunk / W)
                                           : std::max<int64_t>(1, budget / ((26 + (pyr_pull ? 9 : 0) + (pyr_push ? 9 : 0)) * W * std::max<int64_t>(n, 1)));
    chunk = std::min<int64_t>(chunk, nbatch);
    chunk = std::min<int64_t>(chunk, 65535);
    if (chunk > group) chunk -= chunk % group;  // whole groups per chunk
    const int64_t max_groups = (chunk + group - 1) / group;
    const int64_t max_cbatch = (max_groups + 63) / 64;  // coarse batches: 64 groups per word

    DevBuf<u64> valid, stats, work, c_visited, c_frontier, c_next, c_valid, c_stats;
    DevBuf<int> active, mode, any, c_active, c_mode;
    DevBuf<int32_t> counts;
    DevBuf<uint8_t> lvl;
    int lcap = 16;
    // the three big state arrays live in the context's workspace and are reused by later calls
    struct {
        u64 *p;
    } visited, frontier, next;
    VGA_TRY(ctx->ws.get("bfs_visited", sizeof(u64) * (size_t)chunk * n * W, (void **)&visited.p));
    VGA_TRY(ctx->ws.get("bfs_frontier", sizeof(u64) * (size_t)chunk * n * W, (void **)&frontier.p));
    VGA_TRY(ctx->ws.get("bfs_next", sizeof(u64) * (size_t)chunk * n * W, (void **)&next.p));
    u64 *pyr_p = nullptr;
    if (pyr_pull) VGA_TRY(ctx->ws.get("bfs_pyr", sizeof(u64) * (size_t)chunk * (size_t)pl.total * W, (void **)&pyr_p));
    u64 *npyr_p = nullptr;
    if (pyr_push) {
        VGA_TRY(ctx->ws.get("bfs_npyr", sizeof(u64) * (size_t)chunk * (size_t)pl.total * W, (void **)&npyr_p));
        // all zero between levels: the down pass clears what the push wrote
        VGA_CUDA(cudaMemsetAsync(npyr_p, 0, sizeof(u64) * (size_t)chunk * (size_t)pl.total * W, st));
    }
    VGA_TRY(valid.alloc((size_t)chunk * W));
    VGA_TRY(stats.alloc((size_t)chunk * NSTAT));
    VGA_TRY(work.alloc_zero(4, st));
    VGA_TRY(active.alloc((size_t)chunk));
    VGA_TRY(mode.alloc((size_t)chunk));
    VGA_TRY(any.alloc(1));
    VGA_TRY(counts.alloc((size_t)lcap * chunk * W * 64));
    if (coarse) {
        VGA_TRY(c_visited.alloc((size_t)max_cbatch * n));
        VGA_TRY(c_frontier.alloc((size_t)max_cbatch * n));
        VGA_TRY(c_next.alloc((size_t)max_cbatch * n));
        VGA_TRY(c_valid.alloc((size_t)max_cbatch));
        VGA_TRY(c_stats.alloc((size_t)max_cbatch * NSTAT));
        VGA_TRY(c_active.alloc((size_t)max_cbatch));
        VGA_TRY(c_mode.alloc((size_t)max_cbatch));
        VGA_TRY(lvl.alloc((size_t)max_cbatch * 64 * n));
    }

    BfsDev d = {};  // every pointer of an unused option must be null
    d.n = n;
    d.rowptr = g->rowptr.p;
    d.adj = g->adj.p;
    d.t_rowptr = g->has_transpose ? g->t_rowptr.p : nullptr;
    d.t_col = g->has_transpose ? g->t_col.p : nullptr;
    d.visited = visited.p;
    d.frontier = frontier.p;
    d.next = next.p;
    d.valid = valid.p;
    d.active = active.p;
    d.mode = mode.p;
    d.stats = stats.p;
    d.any = any.p;
    d.lvl_in = coarse ? lvl.p : nullptr;
    d.lvl_out = nullptr;
    d.group = group;
    d.pyr = pyr_p;
    d.pyr_total = pl.total;
    d.t_runptr = pyr_pull ? g->t_runptr.p : nullptr;
    d.t_runs = pyr_pull ? g->t_runs.p : nullptr;
    d.t_costptr = pyr_pull ? g->t_costptr.p : nullptr;
    d.noexpand = (radius != -1 && g->noexpand.p) ? g->noexpand.p : nullptr;
    d.npyr = npyr_p;
    d.push_force = ctx->opt.bfs_push == 2 ? 1 : 0;
    d.pyr_weight = (int)std::max<int64_t>(1, ctx->opt.bfs_pyr_cost);
    d.f_nodes = (node_lists && pyr_push) ? g->f_nodes.p : nullptr;
    d.t_nodes = (node_lists && pyr_pull) ? g->t_nodes.p : nullptr;
    d.f_runptr = pyr_push ? g->f_runptr.p : nullptr;
    d.f_runs = pyr_push ? g->f_runs.p : nullptr;
    d.f_costptr = pyr_push ? g->f_costptr.p : nullptr;
    d.pyr_levels = pl.levels;
    for (int k = 0; k < PYR_LEVELS_DEV; k++) {
        d.pyr_off[k] = pl.off[k];
        d.pyr_cnt[k] = pl.cnt[k];
    }

    int deepest = 0;
    std::vector<int32_t> h_counts;
    std::vector<u64> h_valid;
    std::vector<int> ones;
    const int64_t cstride = chunk * W * 64;  // counts per level

    for (int64_t b0 = 0; b0 < nbatch; b0 += chunk) {
        const int64_t cb = std::min<int64_t>(chunk, nbatch - b0);
        const int64_t first = b0 * W * 64;  // index into the ordered source list
        const int64_t cs = std::min<int64_t>(cb * W * 64, nsrc - first);
        if (ctx->cancel && ctx->cancel(ctx->user)) {
            set_error("cancelled");
            return VGA_ERR_CANCELLED;
        }
        kt.start();
        // ---- coarse pass: group lower bounds, 64 groups per bit-parallel batch (W = 1, push only)
        if (coarse) {
            const int64_t ngroups = (cb + group - 1) / group;
            const int64_t ng = (ngroups + 63) / 64;
            BfsDev c = d;
            c.visited = c_visited.p;
            c.frontier = c_frontier.p;
            c.next = c_next.p;
            c.valid = c_valid.p;
            c.active = c_active.p;
            c.mode = c_mode.p;
            c.stats = c_stats.p;
            c.lvl_in = nullptr;
            c.lvl_out = lvl.p;
            c.group = 1;
            c.pyr = nullptr;
            c.t_costptr = nullptr;
            c.npyr = nullptr;
            c.f_costptr = nullptr;
            c.f_nodes = nullptr;
            c.t_nodes = nullptr;
            h_valid.assign((size_t)ng, ~0ULL);
            if (ngroups & 63) h_valid[(size_t)ng - 1] = (1ULL << (ngroups & 63)) - 1ULL;
            ones.assign((size_t)ng, 1);
            VGA_CUDA(cudaMemcpyAsync(c_valid.p, h_valid.data(), sizeof(u64) * ng, cudaMemcpyHostToDevice, st));
            VGA_CUDA(cudaMemcpyAsync(c_active.p, ones.data(), sizeof(int) * ng, cudaMemcpyHostToDevice, st));
            VGA_CUDA(cudaMemsetAsync(c_visited.p, 0, sizeof(u64) * (size_t)ng * n, st));
            VGA_CUDA(cudaMemsetAsync(c_frontier.p, 0, sizeof(u64) * (size_t)ng * n, st));
            VGA_CUDA(cudaMemsetAsync(c_next.p, 0, sizeof(u64) * (size_t)ng * n, st));
            VGA_CUDA(cudaMemsetAsync(c_stats.p, 0, sizeof(u64) * (size_t)ng * NSTAT, st));
            VGA_CUDA(cudaMemsetAsync(c_mode.p, 0, sizeof(int) * (size_t)ng, st));
            VGA_CUDA(cudaMemsetAsync(lvl.p, 0xff, (size_t)ng * 64 * n, st));
            VGA_CUDA(cudaStreamSynchronize(st));  // host staging vectors are reused below
            k_init_coarse<<<blocks_for(cs, 256), 256, 0, st>>>(c, d_order.p + first, cs, 64 * W * group);
            tm.launches++;
            mt.start();
            int cl = 0;
            VGA_TRY(run_levels<1>(ctx, c, ng, radius, 0, nullptr, nullptr, 0, nullptr, mt, &cl));
            mt.stop();
        }
        // ---- exact pass
        {
            const int64_t words_here = (cs + 63) / 64;
            h_valid.assign((size_t)(cb * W), 0ULL);
            for (int64_t wi = 0; wi < words_here; wi++) h_valid[(size_t)wi] = ~0ULL;
            if (cs & 63) h_valid[(size_t)words_here - 1] = (1ULL << (cs & 63)) - 1ULL;
        }
        ones.assign((size_t)cb, 1);
        VGA_CUDA(cudaMemcpyAsync(valid.p, h_valid.data(), sizeof(u64) * cb * W, cudaMemcpyHostToDevice, st));
        VGA_CUDA(cudaMemcpyAsync(active.p, ones.data(), sizeof(int) * cb, cudaMemcpyHostToDevice, st));
        VGA_CUDA(cudaMemsetAsync(visited.p, 0, sizeof(u64) * (size_t)cb * n * W, st));
        VGA_CUDA(cudaMemsetAsync(frontier.p, 0, sizeof(u64) * (size_t)cb * n * W, st));
        VGA_CUDA(cudaMemsetAsync(next.p, 0, sizeof(u64) * (size_t)cb * n * W, st));
        VGA_CUDA(cudaMemsetAsync(stats.p, 0, sizeof(u64) * (size_t)cb * NSTAT, st));
        VGA_CUDA(cudaMemsetAsync(mode.p, 0, sizeof(int) * (size_t)cb, st));  // level 0 always pushes
        VGA_CUDA(cudaMemsetAsync(counts.p, 0, sizeof(int32_t) * (size_t)lcap * cstride, st));
        VGA_CUDA(cudaStreamSynchronize(st));
        k_init<W><<<blocks_for(cs, 256), 256, 0, st>>>(d, d_order.p + first, cs);
        tm.launches++;
        if (d.npyr) {
            k_init_mode<W><<<blocks_for(cb * 32, 256), 256, 0, st>>>(d, d_order.p + first, cs, (int)cb);
            tm.launches++;
        }
        mt.start();
        int nlev = 1;
        VGA_TRY(run_levels<W>(ctx, d, cb, radius, bfs_mode, &counts, &lcap, cstride, work.p, mt, &nlev));
        mt.stop();
        kt.stop();
        VGA_CUDA(cudaGetLastError());
        deepest = std::max(deepest, nlev);

        // results of this chunk: level histogram -> host
        dt.start();
        h_counts.resize((size_t)nlev * cstride);
        VGA_CUDA(cudaMemcpyAsync(h_counts.data(), counts.p, sizeof(int32_t) * (size_t)nlev * cstride, cudaMemcpyDeviceToHost, st));
        dt.stop();
        for (int64_t i = 0; i < cs; i++) {
            const int64_t o = (int64_t)order[(size_t)(first + i)] - src_begin;
            const int64_t wi = i >> 6;  // word index within the chunk = b*W + j
            const int bit = (int)(i & 63);
            int64_t tn = 1, td = 0;
            if (dist && max_levels > 0) {
                for (int l = 0; l < max_levels; l++) dist[o * max_levels + l] = 0;
                dist[o * max_levels] = 1;
            }
            for (int l = 1; l < nlev; l++) {
                int32_t c = h_counts[(size_t)l * cstride + (size_t)wi * 64 + bit];
                tn += c;
                td += (int64_t)l * c;
                if (dist && l < max_levels) dist[o * max_levels + l] = c;
            }
            if (total_nodes) total_nodes[o] = (int32_t)tn;
            if (total_depth) total_depth[o] = td;
        }
        if (ctx->progress) ctx->progress(ctx->user, std::min<int64_t>(nsrc, (b0 + cb) * W * 64), nsrc);
    }
    if (levels_used) *levels_used = deepest;
    if (dbg)
        fprintf(stderr, "[vga_global] transpose %.2f ms, source order %.2f ms, alloc+levels+results %.2f ms (level kernels %.2f)\n",
                w1 - w0, w2 - w1, wall() - w2, tm.main_kernel_ms);
    {
        u64 hw[4] = {0, 0, 0, 0};
        VGA_CUDA(cudaMemcpy(hw, work.p, sizeof(hw), cudaMemcpyDeviceToHost));
        uint64_t rp[2] = {0, 0};
        VGA_CUDA(cudaMemcpy(&rp[0], g->rowptr.p + src_begin, sizeof(uint64_t), cudaMemcpyDeviceToHost));
        VGA_CUDA(cudaMemcpy(&rp[1], g->rowptr.p + src_end, sizeof(uint64_t), cudaMemcpyDeviceToHost));
        // algorithmic bytes (SURVEY.md §8d with B = 64*W sources per batch, CSR rows of 4-byte entries), per
        // batch and level: rows of the expanding vertices + one frontier vector (B/8 bytes) read per
        // expanding vertex + one visited/next read-modify-write (2 vectors) per newly reached vertex;
        // level 0 expands the sources.
        const double src_edges = (double)(rp[1] - rp[0]);
        const double vec = 8.0 * W;
        tm.algo_bytes = 4.0 * ((double)hw[0] + src_edges) + vec * ((double)hw[1] + (double)nsrc) + 2.0 * vec * (double)hw[1];
        // the same model for the run-length rows of bfs_push = 1: 8 bytes per run (first ordinal, length) instead of 4 per
        // entry; the sources' own rows counted from their run counts
        tm.algo_bytes_runs = 0.0;
        if (pyr_push) {
            uint64_t fr[2] = {0, 0};
            VGA_CUDA(cudaMemcpy(&fr[0], g->f_runptr.p + src_begin, sizeof(uint64_t), cudaMemcpyDeviceToHost));
            VGA_CUDA(cudaMemcpy(&fr[1], g->f_runptr.p + src_end, sizeof(uint64_t), cudaMemcpyDeviceToHost));
            tm.algo_bytes_runs = 8.0 * ((double)hw[3] + (double)(fr[1] - fr[0])) + vec * ((double)hw[1] + (double)nsrc) +
                                 2.0 * vec * (double)hw[1];
        }
    }
    if (dist && deepest > max_levels) {
        set_error("vga_global: level histogram needs " + std::to_string(deepest) + " columns");
        return VGA_ERR_CAPACITY;
    }
    return VGA_OK;
}

}  // namespace

int batch_source_order(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, std::vector<int32_t> &order) {
    return source_order(ctx, g, src_begin, src_end, order);
}

int ensure_transpose(vga_ctx *ctx, vga_graph *g) {
    if (g->has_transpose) return VGA_OK;
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    DevBuf<u64> indeg, cursor;
    VGA_TRY(indeg.alloc_zero((size_t)n + 1, st));
    VGA_TRY(g->t_rowptr.alloc((size_t)n + 1));
    if (g->entries > 0) {
        k_indeg<<<blocks_for(g->entries, 256), 256, 0, st>>>(g->adj.p, (uint64_t)g->entries, (uint32_t)n, indeg.p);
        ctx->timing.launches++;
    }
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, indeg.p, (u64 *)g->t_rowptr.p, (int)(n + 1), st);
    DevBuf<unsigned char> tmp;
    VGA_TRY(tmp.alloc(tb + 16));
    cub::DeviceScan::ExclusiveSum(tmp.p, tb, indeg.p, (u64 *)g->t_rowptr.p, (int)(n + 1), st);
    ctx->timing.launches++;
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, g->t_rowptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    g->t_entries = (int64_t)total;
    VGA_TRY(g->t_col.alloc((size_t)total + 1));
    VGA_TRY(cursor.alloc((size_t)n + 1));
    VGA_CUDA(cudaMemcpyAsync(cursor.p, g->t_rowptr.p, sizeof(u64) * (n + 1), cudaMemcpyDeviceToDevice, st));
    if (n > 0) {
        k_scatter_t<<<blocks_for(n * 32, 256), 256, 0, st>>>(g->rowptr.p, g->adj.p, n, cursor.p, g->t_col.p);
        ctx->timing.launches++;
    }
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    g->has_transpose = true;
    return VGA_OK;
}

// Run-length form of sorted rows: one run per maximal stretch of consecutive ordinals (ghost columns excluded).
static int build_runs(vga_ctx *ctx, int64_t n, const uint64_t *rowptr, const uint32_t *arr, int shift, DevBuf<uint64_t> &runptr,
                      DevBuf<uint2> &runs, DevBuf<uint64_t> &costptr, int64_t *nruns) {
    cudaStream_t st = ctx->stream;
    DevBuf<u64> count, cost;
    DevBuf<uint32_t> nvalid, first_off;
    DevBuf<unsigned char> tmp;
    VGA_TRY(count.alloc_zero((size_t)n + 1, st));
    VGA_TRY(cost.alloc_zero((size_t)n + 1, st));
    VGA_TRY(nvalid.alloc_zero((size_t)n + 1, st));
    VGA_TRY(runptr.alloc((size_t)n + 1));
    VGA_TRY(costptr.alloc((size_t)n + 1));
    if (n > 0) {
        k_count_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(rowptr, arr, shift, (uint32_t)n, n, count.p, nvalid.p);
        ctx->timing.launches++;
    }
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, count.p, (u64 *)runptr.p, (int)(n + 1), st);
    VGA_TRY(tmp.alloc(tb + 16));
    VGA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, count.p, (u64 *)runptr.p, (int)(n + 1), st));
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, runptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    *nruns = (int64_t)total;
    VGA_TRY(first_off.alloc((size_t)total + 1));
    VGA_TRY(runs.alloc((size_t)total + 1));
    if (n > 0) {
        k_mark_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(rowptr, arr, shift, nvalid.p, n, runptr.p, first_off.p);
        k_emit_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(rowptr, arr, shift, nvalid.p, n, runptr.p, first_off.p, runs.p, cost.p);
        ctx->timing.launches += 2;
    }
    VGA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cost.p, (u64 *)costptr.p, (int)(n + 1), st));
    ctx->timing.launches += 2;
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// in-rows for the pyramid pull: sort every in-row first.  Built once per graph.
int ensure_runs(vga_ctx *ctx, vga_graph *g) {
    if (g->has_runs) return VGA_OK;
    VGA_TRY(ensure_transpose(ctx, g));
    if (g->t_entries > 0) {
        DevBuf<uint32_t> sorted;
        VGA_TRY(sorted.alloc((size_t)g->t_entries + 1));
        VGA_TRY(sort_segments_u32(ctx, g->t_col.p, sorted.p, g->t_entries, g->n, g->t_rowptr.p));
        g->t_col = std::move(sorted);
    }
    VGA_TRY(build_runs(ctx, g->n, g->t_rowptr.p, g->t_col.p, 0, g->t_runptr, g->t_runs, g->t_costptr, &g->t_nruns));
    g->has_runs = true;
    return VGA_OK;
}

// out-rows for the pyramid push (rows are kept sorted by column; packed entries col << 6 | ...)
int ensure_fwd_runs(vga_ctx *ctx, vga_graph *g) {
    if (g->has_fwd_runs) return VGA_OK;
    VGA_TRY(build_runs(ctx, g->n, g->rowptr.p, g->adj.p, 6, g->f_runptr, g->f_runs, g->f_costptr, &g->f_nruns));
    g->has_fwd_runs = true;
    return VGA_OK;
}

// node-id lists from runs: per-run node counts -> exclusive scan -> ids written at the run's offset
static int build_nodes(vga_ctx *ctx, int64_t n, const DevBuf<uint2> &runs, int64_t nruns, DevBuf<uint32_t> &nodes) {
    cudaStream_t st = ctx->stream;
    const PyrLayout pl = pyr_layout(n);
    BfsDev d0;
    memset(&d0, 0, sizeof(d0));
    d0.pyr_levels = pl.levels;
    for (int k = 0; k < PYR_LEVELS_DEV; k++) {
        d0.pyr_off[k] = pl.off[k];
        d0.pyr_cnt[k] = pl.cnt[k];
    }
    DevBuf<u64> cost, off;
    DevBuf<unsigned char> tmp;
    VGA_TRY(cost.alloc_zero((size_t)nruns + 1, st));
    VGA_TRY(off.alloc((size_t)nruns + 1));
    if (nruns > 0) {
        k_run_costs<<<blocks_for(nruns, 256), 256, 0, st>>>(runs.p, nruns, cost.p);
        ctx->timing.launches++;
    }
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, cost.p, off.p, (int)(nruns + 1), st);
    VGA_TRY(tmp.alloc(tb + 16));
    VGA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cost.p, off.p, (int)(nruns + 1), st));
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, off.p + nruns, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_TRY(nodes.alloc((size_t)total + 1));
    if (nruns > 0) {
        k_emit_nodes<<<blocks_for(nruns, 256), 256, 0, st>>>(runs.p, nruns, off.p, (uint32_t)n, d0, nodes.p);
        ctx->timing.launches += 2;
    }
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

int ensure_node_lists(vga_ctx *ctx, vga_graph *g, bool fwd, bool transposed) {
    if (fwd && !g->has_f_nodes) {
        VGA_TRY(ensure_fwd_runs(ctx, g));
        VGA_TRY(build_nodes(ctx, g->n, g->f_runs, g->f_nruns, g->f_nodes));
        g->has_f_nodes = true;
    }
    if (transposed && !g->has_t_nodes) {
        VGA_TRY(ensure_runs(ctx, g));
        VGA_TRY(build_nodes(ctx, g->n, g->t_runs, g->t_nruns, g->t_nodes));
        g->has_t_nodes = true;
    }
    return VGA_OK;
}

int run_global(vga_ctx *ctx, vga_graph *g, int radius, int64_t src_begin, int64_t src_end, int32_t *total_nodes,
               int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used) {
    const int64_t n = g->n;
    if (g->src_begin != 0 || g->src_end != n) {
        set_error("vga_global: the graph must hold the rows of all cells (gather the shards first)");
        return VGA_ERR_INVALID;
    }
    if (src_end < 0 || src_end > n) src_end = n;
    if (src_begin < 0) src_begin = 0;
    if (levels_used) *levels_used = 0;
    if (src_end <= src_begin) return VGA_OK;
    int words = (int)ctx->opt.bfs_words;
    if (words <= 0) {
        // auto (measured, profiles/): while the adjacency is L2 resident one word per batch is as fast as
        // any; once rows stream from HBM two words halve that traffic; very long rows (large open halls)
        // amortise best over four
        const double avg_deg = n > 0 ? (double)g->entries / (double)n : 0.0;
        if ((double)g->entries * 4.0 <= 256e6)
            words = 1;
        else if (avg_deg >= 8192.0)
            words = 4;
        else
            words = 2;
    }
    if ((src_end - src_begin) <= 64) words = 1;
    switch (words) {
    case 4: return run_global_w<4>(ctx, g, radius, src_begin, src_end, total_nodes, total_depth, dist, max_levels, levels_used);
    case 2: return run_global_w<2>(ctx, g, radius, src_begin, src_end, total_nodes, total_depth, dist, max_levels, levels_used);
    default: return run_global_w<1>(ctx, g, radius, src_begin, src_end, total_nodes, total_depth, dist, max_levels, levels_used);
    }
}

}  // namespace vga
