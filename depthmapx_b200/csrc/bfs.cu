// All-sources BFS of VGA visibility analysis (VGAVisualGlobal::run + extractUnseen,
// salalib/vgamodules/vgavisualglobal.cpp:66-130, 218-240) as a bit-parallel multi-source BFS over
// run-length rows and OR-pyramids.
//
// B = 64*W sources form a batch (W = 1, 2, 4 or 8 machine words); a batch carries W consecutive 64-bit words per
// vertex for each of visited / frontier / next.  Batches are spatially compact groups of sources (floods of 64 cells
// over the direct neighbour links, seeded along 8x8 tiles in Morton order): cells that are close and not separated by
// a wall see almost the same set, so a vertex joins the batch's frontier about once instead of once per level.
// Many batches (a "chunk") advance level by level together: grid = (vertex tiles, batches).
//
// Row format.  A row of a grid visibility graph is the union of a few runs of consecutive ordinals (a cell sees
// contiguous vertical spans; ordinals number the filled cells x-major): ~9 entries per run on office plans, ~35 on
// urban plans.  With a pyramid  P_0 = words of the vertices,  P_k[i] = P_{k-1}[2i] | P_{k-1}[2i+1]  a run is tiled by
// at most 2 nodes per level (pyramid.cuh); a row is stored as the list of the ids of those nodes (4 bytes each,
// measured 1.7-2.6 per run).  The adjacency ENTRIES are never read by the BFS (measured on B200, profiles/: 2.9x
// faster at C2, 16x on a C4 slice, 9.6x on a C5 slice, 13x at C1 than the entry-streaming kernels of round 1).
//
// Per level and batch the step is direction-optimising:
//   push  (top-down)   every vertex with a non-zero frontier vector ORs it into the nodes of its out-row (the node's own
//                      word is the read-before-atomic filter), inner nodes live in a pyramid of `next`; a down pass
//                      (k_pyr_down, three levels per launch) then spreads the inner nodes to the leaves;
//   pull  (bottom-up)  a pyramid of the frontier is built (k_pyr_build); every vertex that still misses some source bit
//                      ORs the nodes of its in-row (8-lane groups, redux.or) and leaves the row as soon as every
//                      missing bit is found.
// `update` folds `next` into visited/frontier and counts the new vertices per source (ballot + popc or a warp
// bit-matrix transposition), i.e. the reference's distribution[level]; `decide` retires finished batches and picks
// the direction of the next step from the node counts of the frontier's out-rows and of the open vertices' in-rows.
//
// In-rows are obtained without an entry-sized transpose: a vertical run of column v starts at row u where v is covered
// by a run of row u but by none of row u-1 (and ends likewise), so the start / end events are the set differences of
// consecutive rows' run lists -- O(runs) work (k_trans_events), then every start finds its end inside its column.
//
// The reference's `extents` run short-circuit is a pure optimisation; the result is the plain level
// structure computed here.  Ghost vertices (unfilled cells inside a diagonal run) are never counted or expanded by
// the reference (p.filled() test, :104) and are not part of any run.
#include <cub/cub.cuh>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
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
    int64_t n;                       // filled vertices
    const uint64_t *f_ptr;           // [n+1] out-rows: offsets into f_nodes
    const uint32_t *f_nodes;         // node ids: id < n = the vertex's own word, id >= n = inner pyramid node id - n
    const uint32_t *f_split;         // [n] leading "new" nodes of every out-row (the others are covered by row u-1), or nullptr
    const uint64_t *t_ptr;           // [n+1] in-rows (nullptr in push-only runs)
    const uint32_t *t_nodes;
    const uint64_t *rowptr;          // [n+1] entry offsets of the CSR rows, or nullptr (statistics of the CSR byte model)
    const uint32_t *deg;             // [n] entries per row when the graph holds runs only, or nullptr
    u64 *visited, *frontier, *next;  // [batches][n][W]
    u64 *pyr;                        // [batches][pyr_total][W]: OR-pyramid levels >= 1 of the frontier (pull)
    u64 *npyr;                       // [batches][pyr_total][W]: inner nodes of `next` (push); all zero between levels
    // y-major alternative lists (nullptr when the graph has none): the same two pyramids over the y-major order of the
    // vertices; their level 0 is the shared x-major state, reached through perm_x (y-major rank -> ordinal)
    u64 *pyr_y, *npyr_y;
    const uint32_t *perm_x;
    const uint8_t *f_isy, *t_isy;    // [n] 1 = the row's list is the y-major one
    int64_t pyr_total;               // inner nodes per batch
    const u64 *valid;                // [batches*W] valid source bits of each word
    int *active;                     // [batches] 1 while the batch goes on
    int *mode;                       // [batches] 0 push, 1 pull
    u64 *stats;                      // [batches][NSTAT]
    int *any;                        // [1] bit 0: some batch still active, bit 1: the level just folded added vertices
    const uint8_t *noexpand;         // [n] or nullptr: vertices that are counted but never join the frontier
    int pyr_levels;
    int64_t pyr_off[PYR_LEVELS_DEV];
    int64_t pyr_cnt[PYR_LEVELS_DEV];
};
// per-batch statistics of a level: 0 CSR entries of the new frontier's rows, 1 in-row nodes of the open vertices,
// 2 new vertices, 3 open vertices, 4 out-row nodes of the new frontier, 5 the out-row nodes the next top-down step will
// visit (= 4 without f_split)
constexpr int NSTAT = 8;
constexpr int UPD_PLANES = 4, UPD_FLUSH = 15;  // k_update: bit-sliced counters hold 2^4 - 1 additions

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
// streaming variants (evict-first): arrays that are read or written once per level must not push the randomly
// accessed pyramids out of L2
template <int W> __device__ __forceinline__ void ldw_stream(const u64 *p, u64 (&o)[W]) {
    if constexpr (W == 1) {
        o[0] = __ldcs(p);
    } else {
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            ulonglong2 t = __ldcs(reinterpret_cast<const ulonglong2 *>(p + j));
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

// LW consecutive words of a vector (what one lane of the lane-cooperative kernels owns): 16-byte accesses for LW = 2
template <int LW> struct Slice {
    u64 w[LW];
};
template <int LW> __device__ __forceinline__ Slice<LW> ld_slice(const u64 *p) {
    Slice<LW> s;
    if constexpr (LW == 2) {
        const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(p);
        s.w[0] = t.x;
        s.w[1] = t.y;
    } else {
        s.w[0] = p[0];
    }
    return s;
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
    // a vertex is a source in exactly one word of a call unless the caller lists it twice: atomics keep both bits
    atomicOr(&d.visited[(b * d.n + v) * W + j], w);
    atomicOr(&d.frontier[(b * d.n + v) * W + j], w);
}

// top-down step over node-id lists: the frontier words of u are ORed into every node of its out-row; the node's own word
// is the read-before-atomic filter (k_update drops words that only repeat reached sources)
template <int W> __global__ void __launch_bounds__(TPB) k_push_nodes(BfsDev d) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 0) return;
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
        if (u < d.n) ldw_stream<W>(fr + u * W, f);
        u64 anyf = 0ULL;
#pragma unroll
        for (int j = 0; j < W; j++) anyf |= f[j];
        // every lane fetches its own row bounds up front (coalesced); they are broadcast below
        uint64_t my0 = 0, my1 = 0;
        if (anyf != 0ULL) {
            my0 = d.f_ptr[u];
            my1 = d.f_ptr[u + 1];
        }
        unsigned m = __ballot_sync(FULL, anyf != 0ULL);
        while (m) {
            int src_lane = __ffs(m) - 1;
            m &= m - 1;
            u64 fw[W];
#pragma unroll
            for (int j = 0; j < W; j++) fw[j] = __shfl_sync(FULL, f[j], src_lane);
            uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            // inner nodes of a y-major list live in the y-major pyramid
            u64 *npu = (d.f_isy && d.f_isy[base + (threadIdx.x & ~31) + src_lane]) ? d.npyr_y + (int64_t)b * d.pyr_total * W : np;
            for (uint64_t e = e0 + lane; e < e1; e += 32) {
                const uint32_t c = __ldcs(d.f_nodes + e);
                VGA_COUNT(npush_nodes, 1);
                u64 *p = c < n ? nx + (int64_t)c * W : npu + (int64_t)(c - n) * W;
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

// down pass of the push, three levels per launch, top chunk first: level k+3 .. k+1 -> level k (`next` for k = 0)
template <int W> __global__ void __launch_bounds__(TPB, W <= 4 ? 4 : 1) k_pyr_down(BfsDev d, int k, int ymaj) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 0) return;
    u64 *np = (ymaj ? d.npyr_y : d.npyr) + (int64_t)b * d.pyr_total * W;
    const uint32_t *leaf = (ymaj && k == 0) ? d.perm_x : nullptr;
    u64 *dst = k == 0 ? d.next + (int64_t)b * d.n * W : np + d.pyr_off[k] * W;
    u64 *s1 = k + 1 < d.pyr_levels ? np + d.pyr_off[k + 1] * W : nullptr;
    u64 *s2 = k + 2 < d.pyr_levels ? np + d.pyr_off[k + 2] * W : nullptr;
    u64 *s3 = k + 3 < d.pyr_levels ? np + d.pyr_off[k + 3] * W : nullptr;
    const int64_t c0 = d.pyr_cnt[k];
    const int64_t c1 = s1 ? d.pyr_cnt[k + 1] : 0, c2 = s2 ? d.pyr_cnt[k + 2] : 0, c3 = s3 ? d.pyr_cnt[k + 3] : 0;
    const int64_t groups = (c0 + 7) / 8;
    for (int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x; t < groups; t += (int64_t)gridDim.x * TPB) {
        VGA_COUNT(pyr_down_groups, 1);
        pyr_down_group<W>(dst, c0, s1, c1, s2, c2, s3, c3, t, leaf);
    }
}

// Three pyramid levels per launch for every batch whose next step is a pull: level k (the frontier for k = 0) ->
// levels k+1 .. k+3.  One work item per aligned group of 8 level-k nodes.
template <int W> __global__ void __launch_bounds__(TPB) k_pyr_build(BfsDev d, int k, int ymaj) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 1) return;
    u64 *pyr = (ymaj ? d.pyr_y : d.pyr) + (int64_t)b * d.pyr_total * W;
    const uint32_t *leaf = (ymaj && k == 0) ? d.perm_x : nullptr;
    const u64 *src = k == 0 ? d.frontier + (int64_t)b * d.n * W : pyr + d.pyr_off[k] * W;
    u64 *d1 = k + 1 < d.pyr_levels ? pyr + d.pyr_off[k + 1] * W : nullptr;
    u64 *d2 = k + 2 < d.pyr_levels ? pyr + d.pyr_off[k + 2] * W : nullptr;
    u64 *d3 = k + 3 < d.pyr_levels ? pyr + d.pyr_off[k + 3] * W : nullptr;
    const int64_t c0 = d.pyr_cnt[k];
    const int64_t c1 = d1 ? d.pyr_cnt[k + 1] : 0, c2 = d2 ? d.pyr_cnt[k + 2] : 0, c3 = d3 ? d.pyr_cnt[k + 3] : 0;
    const int64_t groups = (c0 + 7) / 8;
    for (int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x; t < groups; t += (int64_t)gridDim.x * TPB) {
        VGA_COUNT(pyr_build_groups, 1);
        pyr_build_group<W>(src, c0, d1, c1, d2, c2, d3, c3, t, leaf);
    }
}

// bottom-up step over node-id lists with early exit: candidates of a warp are served by 4 groups of 8 lanes, each
// group scanning a different vertex's in-row (4 rows in flight per warp hide the dependent index->word latency)
template <int W> __global__ void __launch_bounds__(TPB) k_pull_nodes(BfsDev d) {
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 1) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    const u64 *vis = d.visited + (int64_t)b * d.n * W;
    const u64 *pyr = d.pyr + (int64_t)b * d.pyr_total * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
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
        if (w < d.n) {
            u64 vv[W];
            ldw_stream<W>(vis + w * W, vv);
#pragma unroll
            for (int j = 0; j < W; j++) {
                need[j] = valid[j] & ~vv[j];
                anyneed |= need[j];
            }
        }
        uint64_t my0 = 0, my1 = 0;
        if (anyneed != 0ULL) {
            my0 = d.t_ptr[w];
            my1 = d.t_ptr[w + 1];
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
                const u64 *pyu = (d.t_isy && d.t_isy[(base + (threadIdx.x & ~31)) + src_lane]) ? d.pyr_y + (int64_t)b * d.pyr_total * W : pyr;
                for (uint64_t e = e0; e < e1; e += 16) {
                    const uint64_t ea = e + gl, eb = e + 8 + gl;
                    u64 g0[W], g1[W];
#pragma unroll
                    for (int j = 0; j < W; j++) g0[j] = g1[j] = 0ULL;
                    uint32_t ca = 0, cb = 0;
                    if (ea < e1) ca = __ldcs(d.t_nodes + ea);
                    if (eb < e1) cb = __ldcs(d.t_nodes + eb);
                    VGA_COUNT(npull_nodes, (ea < e1) + (eb < e1));
                    if (ea < e1) ldw<W>(ca < n ? fr + (int64_t)ca * W : pyu + (int64_t)(ca - n) * W, g0);
                    if (eb < e1) ldw<W>(cb < n ? fr + (int64_t)cb * W : pyu + (int64_t)(cb - n) * W, g1);
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

// ---- lane-cooperative variants (bfs_coop = 1, W >= 4) ---------------------------------------------------------------
// ncu on the heavy top-down launch of a 10^6-cell plan (profiles/r2_prof_k_push_nodes_C5slice_summary.txt): bound by L1
// data-pipe wavefronts (87 %) -- a warp-wide gather of 32 scattered 16-byte vectors costs one wavefront per vector, and a
// lane that owns a whole W-word vector issues W/2 such gathers, so wider batches bought nothing (266 / 301 / 437 ms for
// W = 2 / 4 / 8).  Here G = W/2 consecutive lanes share one node: each loads / updates its own 16-byte slice, so a node's
// vector arrives in ONE wavefront whatever W is, and the visits per source drop by G.
// PU = nodes per lane whose ids, then vectors, are requested before the first is used (the kernel stalls on the dependent
// id -> vector chain: 80 % long-scoreboard samples with one node in flight, profiles/r2_prof_k_push_nodes_coop_summary.txt)
template <int W, int G, int PU> __global__ void __launch_bounds__(TPB) k_push_nodes_coop(BfsDev d) {
    static_assert(W == 2 * G, "two words per lane");
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 0) return;
    const int lane = threadIdx.x & 31;
    const int g = lane % G, sub = lane / G;
    constexpr int NPI = 32 / G;  // nodes per warp iteration
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    u64 *np = d.npyr + (int64_t)b * d.pyr_total * W;
    const uint32_t n = (uint32_t)d.n;
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        const int64_t wbase = base + (threadIdx.x & ~31);
        int64_t u = base + threadIdx.x;
        u64 anyf = 0ULL;
        if (u < d.n) {
            u64 f[W];
            ldw_stream<W>(fr + u * W, f);
#pragma unroll
            for (int j = 0; j < W; j++) anyf |= f[j];
        }
        uint64_t my0 = 0, my1 = 0;
        if (anyf != 0ULL) {
            my0 = d.f_ptr[u];
            my1 = d.f_ptr[u + 1];
        }
        unsigned m = __ballot_sync(FULL, anyf != 0ULL);
        while (m) {
            int src_lane = __ffs(m) - 1;
            m &= m - 1;
            const uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            // this lane's slice of the vertex's frontier vector (just read by the warp: an L1 / L2 hit)
            const ulonglong2 myf = *reinterpret_cast<const ulonglong2 *>(fr + (wbase + src_lane) * W + 2 * g);
            if ((myf.x | myf.y) == 0ULL) continue;
            u64 *npu = (d.f_isy && d.f_isy[wbase + src_lane]) ? d.npyr_y + (int64_t)b * d.pyr_total * W : np;
            for (uint64_t e = e0 + sub; e < e1; e += NPI * PU) {
                uint32_t c[PU];
#pragma unroll
                for (int i = 0; i < PU; i++) c[i] = e + (uint64_t)i * NPI < e1 ? __ldcs(d.f_nodes + e + (uint64_t)i * NPI) : 0xffffffffu;
                u64 *p[PU];
                ulonglong2 cur[PU];
#pragma unroll
                for (int i = 0; i < PU; i++) {
                    p[i] = (c[i] < n ? nx + (int64_t)c[i] * W : npu + (int64_t)(c[i] - n) * W) + 2 * g;
                    cur[i] = make_ulonglong2(~0ULL, ~0ULL);
                    if (c[i] != 0xffffffffu) {
                        VGA_COUNT(npush_nodes, g == 0);
                        cur[i] = *reinterpret_cast<const ulonglong2 *>(p[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < PU; i++) {
                    const u64 a0 = myf.x & ~cur[i].x, a1 = myf.y & ~cur[i].y;
                    if (a0) atomicOr(&p[i][0], a0);
                    if (a1) atomicOr(&p[i][1], a1);
                }
            }
        }
    }
}


// ---- top-down step that leaves shared nodes to the previous vertex (bfs_delta = 1, default) -------------------------------
// Rows of consecutive ordinals (vertically adjacent cells) are nearly the same set: 93 % of the pyramid nodes of an out-row
// of an urban plan cover only cells that row u-1 holds as well, and with spatially coherent batches a vertex joins the
// frontier with the same source bits as its neighbour at the same level -- so most of what k_push_nodes ORs into a node has
// just been ORed there by the neighbour (ncu: 2 % of the visits add a bit).  Every out-row is stored as [new nodes |
// covered nodes] (k_emit_nodes / k_copy_chosen, f_split).  A vertex u pushes its frontier vector F[u] to its new nodes and
// only F[u] & ~F[u-1] to its covered nodes: every cell of a covered node is in row u-1, and vertex u-1 delivers F[u-1]
// to all of its row in this same launch (by the same rule, recursively down to a vertex whose predecessor holds no bit).
// A warp takes 32 consecutive vertices and walks the new nodes of all of them, plus the covered nodes of the few whose
// vector is not a subset of the predecessor's, as ONE flat list (segment found by binary search in a shared-memory prefix of
// the lengths), so short lists keep every lane busy.
// LW words per lane (2, or 1 for W = 1), G = W / LW lanes share a node (see k_push_nodes_coop), PU nodes per lane in flight.
template <int W, int PU> __global__ void __launch_bounds__(TPB) k_push_delta(BfsDev d) {
    constexpr int LW = W >= 2 ? 2 : 1;
    constexpr int G = W / LW;
    constexpr int NPI = 32 / G;  // nodes per warp and load
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 0) return;
    // per warp: 64 list segments (vertex l: segment 2l = its new nodes, 2l+1 = its covered nodes when it has bits of its own)
    __shared__ uint32_t s_pre[TPB / 32][65];
    __shared__ uint64_t s_ptr[TPB / 32][64];
    __shared__ uint8_t s_isy[TPB / 32][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int g = lane % G, sub = lane / G;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    u64 *np = d.npyr + (int64_t)b * d.pyr_total * W;
    u64 *npy = d.npyr_y ? d.npyr_y + (int64_t)b * d.pyr_total * W : np;
    const uint32_t n = (uint32_t)d.n;
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        const int64_t wbase = base + (threadIdx.x & ~31);
        const int64_t u = base + threadIdx.x;
        u64 f[W];
#pragma unroll
        for (int j = 0; j < W; j++) f[j] = 0ULL;
        if (u < d.n) ldw_stream<W>(fr + u * W, f);
        u64 anyf = 0ULL, anyeff = 0ULL;
#pragma unroll
        for (int j = 0; j < W; j++) {
            // the predecessor's vector: the lane below, or (lane 0) the last vertex of the previous tile
            u64 pf = __shfl_up_sync(FULL, f[j], 1);
            if (lane == 0) pf = (u > 0 && u < d.n) ? fr[(u - 1) * W + j] : 0ULL;
            anyf |= f[j];
            anyeff |= f[j] & ~pf;
        }
        if (__ballot_sync(FULL, anyf != 0ULL) == 0u) continue;
        uint64_t my0 = 0;
        uint32_t cnt_new = 0, cnt_cov = 0;
        if (anyf != 0ULL) {
            my0 = d.f_ptr[u];
            cnt_new = d.f_split[u];
            if (anyeff != 0ULL) cnt_cov = (uint32_t)(d.f_ptr[u + 1] - my0) - cnt_new;
        }
        // exclusive prefix of the segment lengths of the tile
        uint32_t inc = cnt_new + cnt_cov;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += t;
        }
        const uint32_t total = __shfl_sync(FULL, inc, 31);
        __syncwarp();  // the previous tile's readers are done
        s_pre[wid][2 * lane] = inc - cnt_new - cnt_cov;
        s_pre[wid][2 * lane + 1] = inc - cnt_cov;
        if (lane == 31) s_pre[wid][64] = total;
        s_ptr[wid][2 * lane] = my0;
        s_ptr[wid][2 * lane + 1] = my0 + cnt_new;
        s_isy[wid][lane] = (d.f_isy && anyf != 0ULL) ? d.f_isy[u] : (uint8_t)0;
        __syncwarp();
        // all segments of the tile as one flat list: short lists keep every lane busy
        int at = -1;  // segment of this lane's previous element: its elements come in ascending order
        for (uint32_t t0 = 0; t0 < total; t0 += NPI * PU) {
            uint32_t c[PU];
            int seg[PU];
            u64 *p[PU];
            Slice<LW> cur[PU];
#pragma unroll
            for (int i = 0; i < PU; i++) {
                const uint32_t t = t0 + (uint32_t)(i * NPI + sub);
                c[i] = 0xffffffffu;
                seg[i] = 0;
                p[i] = nullptr;
                if (t < total) {
                    // last segment with s_pre <= t (empty segments share their successor's prefix): binary search for the
                    // lane's first element, then a walk forward (segments are about as long as the lane's stride)
                    int lo;
                    if (at < 0) {
                        lo = 0;
                        int hi = 64;
                        while (hi - lo > 1) {
                            const int m = (lo + hi) >> 1;
                            if (s_pre[wid][m] <= t) lo = m; else hi = m;
                        }
                    } else {
                        lo = at;
                        while (s_pre[wid][lo + 1] <= t) lo++;  // s_pre[64] = total > t
                    }
                    at = lo;
                    seg[i] = lo;
                    c[i] = __ldcs(d.f_nodes + s_ptr[wid][lo] + (t - s_pre[wid][lo]));
                    p[i] = (c[i] < n ? nx + (int64_t)c[i] * W : (s_isy[wid][lo >> 1] ? npy : np) + (int64_t)(c[i] - n) * W) + LW * g;
                }
            }
#pragma unroll
            for (int i = 0; i < PU; i++) {
                if (c[i] != 0xffffffffu) {
                    VGA_COUNT(npush_nodes, g == 0);
                    cur[i] = ld_slice<LW>(p[i]);
                }
            }
#pragma unroll
            for (int i = 0; i < PU; i++) {
                if (c[i] != 0xffffffffu) {
                    // the owner's vector (and its predecessor's) were read by this warp a moment ago: L1 hits
                    const int64_t v = wbase + (seg[i] >> 1);
                    Slice<LW> fv = ld_slice<LW>(fr + v * W + LW * g);
                    if ((seg[i] & 1) && v > 0) {
                        const Slice<LW> pf = ld_slice<LW>(fr + (v - 1) * W + LW * g);
#pragma unroll
                        for (int j = 0; j < LW; j++) fv.w[j] &= ~pf.w[j];
                    }
#pragma unroll
                    for (int j = 0; j < LW; j++) {
                        const u64 a = fv.w[j] & ~cur[i].w[j];
                        if (a) {
                            atomicOr(&p[i][j], a);
                            VGA_COUNT(npush_atomics, 1);
                        }
                    }
                }
            }
        }
    }
}

// U = node loads per lane between two early-exit checks (2 or 4): a vertex whose missing bits arrive at a later level scans
// its whole in-row, and the group-wide OR + ballot per check is then pure overhead.  A lane whose own 16-byte slice is
// complete skips its loads (with W >= 4 the slices are 64-source clusters of their own, often complete at different levels).
template <int W, int G, int U> __global__ void __launch_bounds__(TPB) k_pull_nodes_coop(BfsDev d) {
    static_assert(W == 2 * G && G <= 8, "two words per lane, a node within an 8-lane group");
    const int b = blockIdx.y;
    if (!d.active[b] || d.mode[b] != 1) return;
    const int lane = threadIdx.x & 31;
    const u64 *fr = d.frontier + (int64_t)b * d.n * W;
    const u64 *vis = d.visited + (int64_t)b * d.n * W;
    const u64 *pyr = d.pyr + (int64_t)b * d.pyr_total * W;
    u64 *nx = d.next + (int64_t)b * d.n * W;
    const uint32_t n = (uint32_t)d.n;
    const int grp = lane >> 3, gl = lane & 7;
    const int g = gl % G, sub = gl / G;
    constexpr int NPG = 8 / G;  // nodes per 8-lane group and load
    const unsigned gmask = 0xffu << (grp * 8);
    const ulonglong2 valid = *reinterpret_cast<const ulonglong2 *>(d.valid + (int64_t)b * W + 2 * g);
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        const int64_t wbase = base + (threadIdx.x & ~31);
        int64_t w = base + threadIdx.x;
        u64 anyneed = 0ULL;
        if (w < d.n) {
            u64 vv[W];
            ldw_stream<W>(vis + w * W, vv);
#pragma unroll
            for (int j = 0; j < W; j++) anyneed |= d.valid[(int64_t)b * W + j] & ~vv[j];
        }
        uint64_t my0 = 0, my1 = 0;
        if (anyneed != 0ULL) {
            my0 = d.t_ptr[w];
            my1 = d.t_ptr[w + 1];
        }
        const unsigned m = __ballot_sync(FULL, anyneed != 0ULL);
        const int ncand = __popc(m);
        for (int r = 0; r * 4 < ncand; r++) {
            const int k = r * 4 + grp;
            const bool has = k < ncand;
            const int src_lane = has ? (int)__fns(m, 0, k + 1) : 0;
            const uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            if (has) {
                const int64_t ww = wbase + src_lane;
                const ulonglong2 vv = *reinterpret_cast<const ulonglong2 *>(vis + ww * W + 2 * g);
                const u64 nd0 = valid.x & ~vv.x, nd1 = valid.y & ~vv.y;
                u64 acc0 = 0ULL, acc1 = 0ULL;
                bool mine_done = (nd0 | nd1) == 0ULL;
                const u64 *pyu = (d.t_isy && d.t_isy[ww]) ? d.pyr_y + (int64_t)b * d.pyr_total * W : pyr;
                for (uint64_t e = e0; e < e1; e += U * NPG) {
                    u64 x0 = 0ULL, x1 = 0ULL;
                    if (!mine_done) {
                        uint32_t c[U];
#pragma unroll
                        for (int i = 0; i < U; i++) {
                            const uint64_t ei = e + (uint64_t)(i * NPG + sub);
                            c[i] = ei < e1 ? __ldcs(d.t_nodes + ei) : 0xffffffffu;
                        }
#pragma unroll
                        for (int i = 0; i < U; i++) {
                            if (c[i] != 0xffffffffu) {
                                VGA_COUNT(npull_nodes, g == 0);
                                const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(
                                    (c[i] < n ? fr + (int64_t)c[i] * W : pyu + (int64_t)(c[i] - n) * W) + 2 * g);
                                x0 |= t.x;
                                x1 |= t.y;
                            }
                        }
                    }
                    // OR over the lanes of the group that hold the same slice
#pragma unroll
                    for (int o = G; o < 8; o <<= 1) {
                        x0 |= __shfl_xor_sync(gmask, x0, o);
                        x1 |= __shfl_xor_sync(gmask, x1, o);
                    }
                    acc0 |= x0;
                    acc1 |= x1;
                    mine_done = ((acc0 & nd0) == nd0) && ((acc1 & nd1) == nd1);
                    if (__ballot_sync(gmask, !mine_done) == 0u) break;
                }
                if (sub == 0) {
                    const u64 n0 = acc0 & nd0, n1 = acc1 & nd1;
                    if (n0) nx[ww * W + 2 * g] = n0;
                    if (n1) nx[ww * W + 2 * g + 1] = n1;
                }
            }
            __syncwarp();
        }
    }
}

// fold next into visited/frontier, count new vertices per source, gather direction statistics;
// `level_next` = level of the vertices being added
template <int W>
__global__ void __launch_bounds__(TPB, W <= 4 ? 3 : 1) k_update(BfsDev d, int32_t *counts /*[batches*W][64] of level_next*/) {
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
    u64 valid[W];
#pragma unroll
    for (int j = 0; j < W; j++) valid[j] = d.valid[(int64_t)b * W + j];
    int cnt[W][2];  // lane l counts source bits l and l+32 of each word
#pragma unroll
    for (int j = 0; j < W; j++) cnt[j][0] = cnt[j][1] = 0;
    // New vertices per source = column sums of the bit matrix (vertices x sources).  Every thread first adds the words of
    // its own vertices into bit-sliced counters (plane k = bit k of 64 vertical counts, a ripple of half adders: ~2 AND/XOR
    // pairs per word) and only every UPD_FLUSH iterations the warp turns the planes into per-source counts (32 x 32
    // bit-matrix transposes + popc, planes that are zero across the warp skipped): one transpose per word and iteration
    // before, 4 planes per 15 iterations now.
    u64 plane[W][UPD_PLANES];
#pragma unroll
    for (int j = 0; j < W; j++)
#pragma unroll
        for (int k = 0; k < UPD_PLANES; k++) plane[j][k] = 0ULL;
    int pending = 0;
    auto flush = [&]() {
#pragma unroll
        for (int j = 0; j < W; j++)
#pragma unroll
            for (int k = 0; k < UPD_PLANES; k++) {
                const unsigned lo = (unsigned)plane[j][k], hi = (unsigned)(plane[j][k] >> 32);
                if (__any_sync(FULL, lo != 0u)) cnt[j][0] += __popc(warp_transpose32(lo, lane)) << k;
                if (__any_sync(FULL, hi != 0u)) cnt[j][1] += __popc(warp_transpose32(hi, lane)) << k;
                plane[j][k] = 0ULL;
            }
        pending = 0;
    };
    u64 f_edges = 0, u_nodes = 0, n_new = 0, n_open = 0, f_nodes = 0, f_visit = 0;
    for (int64_t base = (int64_t)blockIdx.x * TPB; base < d.n; base += (int64_t)gridDim.x * TPB) {
        int64_t v = base + threadIdx.x;
        u64 nw[W];
#pragma unroll
        for (int j = 0; j < W; j++) nw[j] = 0ULL;
        bool expands = false;
        if (v < d.n) {
            VGA_COUNT(update_words, 1);
            u64 vv[W], xx[W];
            ldw<W>(vis + v * W, vv);
            ldw<W>(nx + v * W, xx);
            u64 anynew = 0ULL, anyneed = 0ULL, anyx = 0ULL;
#pragma unroll
            for (int j = 0; j < W; j++) {
                nw[j] = xx[j] & ~vv[j];  // the push does not filter by `visited`
                vv[j] |= nw[j];
                anynew |= nw[j];
                anyx |= xx[j];
                anyneed |= valid[j] & ~vv[j];
            }
            u64 zero[W];
#pragma unroll
            for (int j = 0; j < W; j++) zero[j] = 0ULL;
            if (anyx) stw<W>(nx + v * W, zero);
            if (anynew) {
                stw<W>(vis + v * W, vv);
                n_new += 1;
                expands = !(d.noexpand && d.noexpand[v]);  // reached and counted, but not expanded (vgavisualglobal.cpp:108-110)
                if (expands) {
                    f_edges += d.rowptr ? d.rowptr[v + 1] - d.rowptr[v] : (d.deg ? d.deg[v] : 0);
                    f_nodes += d.f_ptr[v + 1] - d.f_ptr[v];
                }
            }
            stw<W>(fr + v * W, expands ? nw : zero);
            if (anyneed != 0ULL) {
                n_open += 1;  // still unreached by some source of the batch
                if (d.t_ptr) u_nodes += d.t_ptr[v + 1] - d.t_ptr[v];
            }
        }
        if (d.f_split) {
            // what k_push_delta will visit for v: its new nodes, and the covered ones only if it holds a source bit that the
            // vertex below it does not push (the first vertex of a warp tile is charged in full: 1/32 of the estimate)
            bool sub = lane != 0;
#pragma unroll
            for (int j = 0; j < W; j++) {
                const u64 mine = expands ? nw[j] : 0ULL;
                const u64 pf = __shfl_up_sync(FULL, mine, 1);
                if (mine & ~pf) sub = false;
            }
            if (expands) f_visit += sub ? (u64)d.f_split[v] : d.f_ptr[v + 1] - d.f_ptr[v];
        }
#pragma unroll
        for (int j = 0; j < W; j++) {
            u64 carry = nw[j];
#pragma unroll
            for (int k = 0; k < UPD_PLANES; k++) {
                const u64 t = plane[j][k] & carry;
                plane[j][k] ^= carry;
                carry = t;
            }
        }
        if (++pending == UPD_FLUSH) flush();  // uniform: every thread of the block makes the same number of iterations
    }
    if (pending) flush();
#pragma unroll
    for (int j = 0; j < W; j++) {
        if (cnt[j][0]) atomicAdd(&s_cnt[j * 64 + lane], cnt[j][0]);
        if (cnt[j][1]) atomicAdd(&s_cnt[j * 64 + lane + 32], cnt[j][1]);
    }
    for (int o = 16; o > 0; o >>= 1) {
        f_edges += __shfl_down_sync(FULL, f_edges, o);
        u_nodes += __shfl_down_sync(FULL, u_nodes, o);
        n_new += __shfl_down_sync(FULL, n_new, o);
        n_open += __shfl_down_sync(FULL, n_open, o);
        f_nodes += __shfl_down_sync(FULL, f_nodes, o);
        f_visit += __shfl_down_sync(FULL, f_visit, o);
    }
    if (lane == 0) {
        if (f_edges) atomicAdd(&s_stat[0], f_edges);
        if (u_nodes) atomicAdd(&s_stat[1], u_nodes);
        if (n_new) atomicAdd(&s_stat[2], n_new);
        if (n_open) atomicAdd(&s_stat[3], n_open);
        if (f_nodes) atomicAdd(&s_stat[4], f_nodes);
        if (f_visit) atomicAdd(&s_stat[5], f_visit);
    }
    __syncthreads();
    if (counts)
        for (int i = threadIdx.x; i < W * 64; i += TPB)
            if (s_cnt[i]) atomicAdd(&counts[(int64_t)b * W * 64 + i], s_cnt[i]);
    if (threadIdx.x < 6 && s_stat[threadIdx.x]) atomicAdd(&d.stats[b * NSTAT + threadIdx.x], s_stat[threadIdx.x]);
}

// per batch: retire finished batches, choose the next step's direction, reset statistics
__global__ void k_decide(BfsDev d, int nb, int bfs_mode, int64_t alpha, int64_t beta, int64_t wdelta, u64 *work /*[4]*/) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    if (!d.active[b]) return;
    const u64 fe = d.stats[b * NSTAT + 0], un = d.stats[b * NSTAT + 1], nn = d.stats[b * NSTAT + 2], open = d.stats[b * NSTAT + 3];
    const u64 fn = d.stats[b * NSTAT + 4];
    // nodes the next top-down step would visit; a visit of k_push_delta nearly always ends in an atomic and costs
    // `wdelta` times a node offered to the bottom-up step (measured, profiles/r2_level_timing_C5slice.log)
    const u64 fv = d.f_split ? d.stats[b * NSTAT + 5] * (u64)wdelta : fn;
    for (int i = 0; i < 6; i++) d.stats[b * NSTAT + i] = 0;
    // Retire the batch when nothing new was reached -- or when every vertex has been reached by every source of the
    // batch: expanding the last (largest) frontier could not find anything.
    if (nn != 0) atomicOr(d.any, 2);  // bit 1: this level added vertices (the histogram has one more level)
    if (nn == 0 || open == 0) {
        d.active[b] = 0;
        return;
    }
    atomicOr(d.any, 1);  // bit 0: some batch goes on
    // cost of the next step in node visits: top-down = the out-row nodes of the new frontier + the down pass over ~2n
    // nodes; bottom-up = the in-row nodes of the open vertices (an upper bound: rows are left early) + the build (~n)
    int m = 0;
    if (bfs_mode == 1)
        m = 1;
    else if (bfs_mode == 2)
        m = ((fv + 2 * (u64)d.n) * (u64)alpha > (un + (u64)d.n) * (u64)beta) ? 1 : 0;
    d.mode[b] = m;
    if (work) {
        // work[0]: adjacency entries of the vertices that expand next (SURVEY.md §8d CSR model, sum of deg over U_l),
        // work[1]: vertices newly reached (|U_l| summed), work[2]: in-row nodes offered to the pull steps,
        // work[3]: out-row nodes of the vertices that expand next (node-list model)
        atomicAdd(&work[0], fe);
        atomicAdd(&work[1], nn);
        if (m == 1) atomicAdd(&work[2], un);
        atomicAdd(&work[3], fn);
    }
}

// bit k of gc[v] = the k-th of the 8 grid neighbours (E, NE, N, NW, W, SW, S, SE) is in v's out-row; the row is a sorted
// list of runs, so each lookup is a binary search.  Used to grow source batches that do not cross walls.
__global__ void k_neighbour_bits(int64_t n, const uint64_t *runptr, const uint2 *runs, const int32_t *refs, const int32_t *ord_of,
                                 int cols, int rows, uint8_t *gc) {
    int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    uint32_t r = (uint32_t)refs[v];
    int x = (int)(r >> 16), y = (int)(r & 0xffff);
    uint64_t e0 = runptr[v], e1 = runptr[v + 1];
    uint8_t out = 0;
    for (int k = 0; k < 8; k++) {
        int nx = x + dx[k], ny = y + dy[k];
        if (nx < 0 || nx >= cols || ny < 0 || ny >= rows) continue;
        int32_t w = ord_of[(int64_t)nx * rows + ny];
        if (w < 0) continue;
        uint64_t lo = e0, hi = e1;  // first run that ends beyond w
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if (runs[mid].x + runs[mid].y <= (uint32_t)w) lo = mid + 1; else hi = mid;
        }
        if (lo < e1 && runs[lo].x <= (uint32_t)w) out |= (uint8_t)(1 << k);
    }
    gc[v] = out;
}

// ---- run-length rows from SORTED entry rows: one warp per vertex ----------------------------------------------------
// entry e of row v holds ordinal arr[e] >> shift.  A run starts at an entry that is the first of the row, does not
// continue the previous ordinal, or is the first ghost column (ordinal == brk = number of filled cells): runs never
// straddle the filled / ghost boundary, and the ghost runs (first >= brk) are the tail of every row.  The BFS skips
// them (ghosts are never counted or expanded), the local measures need them.
__global__ void k_count_runs(const uint64_t *rowptr, const uint32_t *arr, int shift, uint32_t brk, int64_t n, u64 *count) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint64_t e0 = rowptr[v], e1 = rowptr[v + 1];
    unsigned c = 0;
    for (uint64_t e = e0 + lane; e < e1; e += 32) {
        const uint32_t x = arr[e] >> shift;
        c += (e == e0 || x != (arr[e - 1] >> shift) + 1u || x == brk) ? 1u : 0u;
    }
    c = __reduce_add_sync(FULL, c);
    if (lane == 0) count[v] = c;
}
// first_off[r] = offset (within its row) of the entry that starts run r
__global__ void k_mark_runs(const uint64_t *rowptr, const uint32_t *arr, int shift, uint32_t brk, int64_t n,
                            const uint64_t *runptr, uint32_t *first_off) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint64_t e0 = rowptr[v], e1 = rowptr[v + 1];
    uint64_t out = runptr[v];
    for (uint64_t eb = e0; eb < e1; eb += 32) {
        const uint64_t e = eb + lane;
        const bool start = e < e1 && (e == e0 || (arr[e] >> shift) != (arr[e - 1] >> shift) + 1u || (arr[e] >> shift) == brk);
        const unsigned mask = __ballot_sync(FULL, start);
        if (start) first_off[out + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(e - e0);
        out += __popc(mask);
    }
}
// runs[r] = (first ordinal, length)
__global__ void k_emit_runs(const uint64_t *rowptr, const uint32_t *arr, int shift, int64_t n,
                            const uint64_t *runptr, const uint32_t *first_off, uint2 *runs) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint64_t e0 = rowptr[v];
    const uint32_t nv = (uint32_t)(rowptr[v + 1] - e0);
    const uint64_t r0 = runptr[v], r1 = runptr[v + 1];
    for (uint64_t r = r0 + lane; r < r1; r += 32) {
        const uint32_t off = first_off[r];
        const uint32_t nxt = r + 1 < r1 ? first_off[r + 1] : nv;
        runs[r] = make_uint2(arr[e0 + off] >> shift, nxt - off);
    }
}

// ---- transposition of run-length rows in O(runs) --------------------------------------------------------------------
// The columns of run [a, b) of a row that no run of the row `other` = [o0, o1) covers (both sorted by first ordinal).
template <typename Emit>
__device__ __forceinline__ void uncovered_columns(const uint2 *runs, uint64_t o0, uint64_t o1, uint32_t a, uint32_t b, Emit &&emit) {
    uint64_t lo = o0, hi = o1;  // first run of `other` that ends beyond a
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (runs[mid].x + runs[mid].y <= a) lo = mid + 1; else hi = mid;
    }
    uint32_t cur = a;
    while (cur < b) {
        const uint32_t nf = lo < o1 ? runs[lo].x : 0xffffffffu;
        if (nf <= cur) {
            const uint32_t e = runs[lo].x + runs[lo].y;
            cur = e < b ? e : b;
            lo++;
        } else {
            const uint32_t ge = nf < b ? nf : b;
            for (uint32_t v = cur; v < ge; v++) emit(v);
            cur = ge;
        }
    }
}
// One warp per row u, one lane per run.  A vertical run of column v STARTS at u when v is in a run of row u but in none
// of row u-1, and ENDS before u+1 when it is in none of row u+1.
//   MODE 0: count[v] += 1 per start.   MODE 1: start / end rows written into the column's slots (any order).
template <int MODE>
__global__ void k_trans_events(int64_t n, const uint64_t *runptr, const uint2 *runs, u64 *count, const uint64_t *t_runptr,
                               uint32_t *cur_s, uint32_t *cur_e, uint32_t *t_start, uint32_t *t_end) {
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    const uint64_t r0 = runptr[u], r1 = runptr[u + 1];
    const uint64_t p0 = u > 0 ? runptr[u - 1] : r0, p1 = r0;          // row u-1 (empty for u = 0)
    const uint64_t q0 = r1, q1 = u + 1 < n ? runptr[u + 2] : r1;      // row u+1 (empty for the last row)
    for (uint64_t r = r0 + lane; r < r1; r += 32) {
        const uint2 run = runs[r];
        if (run.x >= (uint32_t)n) continue;  // ghost columns have no in-row
        if (MODE == 0) {
            uncovered_columns(runs, p0, p1, run.x, run.x + run.y, [&](uint32_t v) { atomicAdd(&count[v], 1ULL); });
        } else {
            uncovered_columns(runs, p0, p1, run.x, run.x + run.y,
                              [&](uint32_t v) { t_start[t_runptr[v] + atomicAdd(&cur_s[v], 1u)] = (uint32_t)u; });
            uncovered_columns(runs, q0, q1, run.x, run.x + run.y,
                              [&](uint32_t v) { t_end[t_runptr[v] + atomicAdd(&cur_e[v], 1u)] = (uint32_t)u + 1u; });
        }
    }
}
// One warp per column: the vertical runs of a column are disjoint, so the end of the run that starts at s is the smallest
// end beyond s.  t_runs[slot of the start] = (s, end - s).
__global__ void k_trans_pair(int64_t n, const uint64_t *t_runptr, const uint32_t *t_start, const uint32_t *t_end, uint2 *t_runs) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const uint64_t b0 = t_runptr[v], b1 = t_runptr[v + 1];
    for (uint64_t i = b0 + lane; i < b1; i += 32) {
        const uint32_t s = t_start[i];
        uint32_t best = 0xffffffffu;
        for (uint64_t j = b0; j < b1; j++) {
            const uint32_t e = t_end[j];
            if (e > s && e < best) best = e;
        }
        t_runs[i] = make_uint2(s, best - s);
    }
}

// ---- node-id lists from runs -------------------------------------------------------------------------------------------
__global__ void k_run_costs(const uint2 *runs, int64_t nruns, uint32_t n, u64 *cost) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r < nruns) cost[r] = runs[r].x < n ? (u64)pyr_cost(runs[r].x, runs[r].y) : 0ULL;  // ghost runs: no nodes
}
// ids of the nodes of every run, written at the run's offset (exclusive scan of the per-run node counts); one warp per
// row, one lane per run.
// `leaf` (y-major lists): a level-0 node i is written as the vertex leaf[i] (its x-major ordinal) -- leaves are always the
// vertices' own words, only the inner nodes belong to the list's own pyramid.
// `cover` (out-rows): bit 31 of an id (NODE_COVERED) is set when every cell the node covers also belongs to the PREVIOUS
// row (row u-1 of the same run arrays, whose runs end where this row's begin): the top-down step then leaves the node
// to the previous vertex whenever that vertex pushes a superset of the source bits (k_push_delta).
constexpr uint32_t NODE_COVERED = 0x80000000u;
__global__ void k_emit_nodes(int64_t rows, const uint64_t *runptr, const uint2 *runs, const u64 *node_off, uint32_t n, BfsDev d,
                             const uint32_t *leaf, int cover, uint32_t *out) {
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= rows) return;
    const uint64_t r0 = runptr[u], r1 = runptr[u + 1];
    const uint64_t p0 = (cover && u > 0) ? runptr[u - 1] : r0;
    // end of the previous row's run that contains cell s (0 = none); ghost runs start at >= n and sort last
    auto covering_end = [&](uint32_t s) -> uint32_t {
        uint64_t a = p0, b = r0;  // first run of [p0, r0) that starts beyond s
        while (a < b) {
            const uint64_t m = (a + b) >> 1;
            if (runs[m].x <= s) a = m + 1; else b = m;
        }
        if (a == p0) return 0u;
        const uint2 pr = runs[a - 1];
        return pr.x + pr.y;
    };
    for (uint64_t r = r0 + lane; r < r1; r += 32) {
        const uint2 run = runs[r];
        if (run.x >= n) continue;
        uint64_t o = node_off[r];
        const uint32_t cend = p0 < r0 ? covering_end(run.x) : 0u;
        const bool whole = cend >= run.x + run.y;
        pyr_decompose(run.x, run.y, [&](int k, uint32_t i) {
            bool cov = whole;
            if (!whole && p0 < r0) {
                const uint32_t s = i << k, e = (i + 1u) << k;  // inside the run: no clipping at n
                cov = (s == run.x ? cend : covering_end(s)) >= e;
            }
            const uint32_t id = k == 0 ? (leaf ? leaf[i] : i) : n + (uint32_t)d.pyr_off[k] + i;
            out[o++] = cov ? (id | NODE_COVERED) : id;
        });
    }
}
// nodeptr[v] = node offset of the first run of row v (node_off has nruns + 1 entries)
__global__ void k_row_node_offsets(int64_t n, const uint64_t *runptr, const u64 *node_off, uint64_t *nodeptr) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v <= n) nodeptr[v] = node_off[runptr[v]];
}
__global__ void k_row_lengths(int64_t n, const uint64_t *rowptr, uint32_t *deg) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < n) deg[v] = (uint32_t)(rowptr[v + 1] - rowptr[v]);
}

// ---- y-major alternative rows ---------------------------------------------------------------------------------------
// Ordinals number the cells x-major, so a row is a list of VERTICAL spans: a cell of a corridor or street that runs along x
// sees a thousand columns and needs a thousand runs, while in y-major numbering the same set is a handful of horizontal
// spans (measured on sampled rows of the 10^6-cell plan: 276 pyramid nodes per row x-major, 307 y-major, 166 with the
// smaller of the two per row; 89 / 57 / 42 on the office plan).  Every row therefore also gets its y-major run list and
// the BFS walks whichever is shorter: inner nodes of a y-major list live in a second pyramid built over the y-major
// order of the vertices, leaves stay the vertices' own words (ids are translated to x-major ordinals when the list is
// written), and the level-0 passes of that pyramid go through the permutation (pyramid.cuh `leaf`).

// y-major rank of every filled cell: flag[y * cols + x] = 1, exclusive scan, gather
__global__ void k_mark_cells_ymajor(int64_t n, const int32_t *refs, int cols, u64 *flag) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t r = (uint32_t)refs[v];
    flag[(int64_t)(r & 0xffff) * cols + (r >> 16)] = 1ULL;
}
__global__ void k_perm_from_scan(int64_t n, const int32_t *refs, int cols, const u64 *scan, uint32_t *perm_y, uint32_t *perm_x) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const uint32_t r = (uint32_t)refs[v];
    const uint32_t y = (uint32_t)scan[(int64_t)(r & 0xffff) * cols + (r >> 16)];
    perm_y[v] = y;
    perm_x[y] = (uint32_t)v;
}

// One CTA per row u: the row's members (x-major runs, ghost runs skipped) are set as bits perm_y[w] of a bitmap of the n
// vertices in shared memory; the bitmap's runs are the row's y-major runs.  MODE 0 counts them, MODE 1 writes them
// (first y-major ordinal, length) in ascending order at out_ptr[u].  Only the word range [lo, hi] the row touches is
// cleared and scanned (perm_y grows along an x-major run, so its ends bound it); one warp per run sets the bits; the
// per-thread start / end counts are combined by shuffle scans.
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t *warp_tot /*[33]*/, uint32_t *total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t w = lane < nw ? warp_tot[lane] : 0u, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, winc, o);
            if (lane >= o) winc += t;
        }
        warp_tot[lane] = winc - w;
        if (lane == 31) warp_tot[32] = winc;
    }
    __syncthreads();
    const uint32_t r = warp_tot[wid] + inc - v;
    *total = warp_tot[32];
    __syncthreads();  // warp_tot may be reused
    return r;
}

template <int MODE>
__global__ void k_yruns(int64_t n, const uint64_t *runptr, const uint2 *runs, const uint32_t *perm_y, u64 *count,
                        const uint64_t *out_ptr, uint2 *out) {
    extern __shared__ __align__(16) uint32_t ysm[];
    const uint32_t words = (uint32_t)((n + 31) >> 5);
    uint32_t *bm = ysm;                 // [words + 1] (one zero word behind the end)
    uint32_t *wt = ysm + words + 1;     // [33] scan scratch, then [2] range
    uint32_t *range = wt + 33;
    uint32_t *run_first = range + 2, *run_off = run_first + blockDim.x;  // [blockDim.x] each: the chunk of runs in flight
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // the whole bitmap is zero between rows: every row clears the range it used
    for (uint32_t i = threadIdx.x; i <= words; i += blockDim.x) bm[i] = 0u;
    if (threadIdx.x == 0) {
        range[0] = 0xffffffffu;
        range[1] = 0u;
    }
    __syncthreads();
    for (int64_t u = blockIdx.x; u < n; u += gridDim.x) {
        const uint64_t r0 = runptr[u], r1 = runptr[u + 1];
        uint32_t lo = 0xffffffffu, hi = 0u;
        // blockDim.x runs at a time: their lengths are scanned and the CELLS are dealt out to the threads (binary search of
        // the cell index in the scanned offsets), so the two dependent loads (run, perm_y) are paid once per chunk
        for (uint64_t rb = r0; rb < r1; rb += blockDim.x) {
            const uint64_t r = rb + threadIdx.x;
            uint2 run = make_uint2(0u, 0u);
            if (r < r1) run = runs[r];
            if (run.x >= (uint32_t)n) run.y = 0u;  // ghost columns are not part of the BFS
            uint32_t cells = 0;
            const uint32_t off = block_excl_scan(run.y, wt, &cells);
            run_first[threadIdx.x] = run.x;
            run_off[threadIdx.x] = off;
            __syncthreads();
            const uint32_t nr = (uint32_t)min((uint64_t)blockDim.x, r1 - rb);
            for (uint32_t c = threadIdx.x; c < cells; c += blockDim.x) {
                uint32_t a = 0, b = nr;  // last run whose offset is <= c
                while (b - a > 1) {
                    const uint32_t m = (a + b) >> 1;
                    if (run_off[m] <= c) a = m; else b = m;
                }
                const uint32_t y = perm_y[run_first[a] + (c - run_off[a])];
                atomicOr(&bm[y >> 5], 1u << (y & 31));
                lo = min(lo, y >> 5);
                hi = max(hi, y >> 5);
            }
            __syncthreads();
        }
        lo = __reduce_min_sync(FULL, lo);
        hi = __reduce_max_sync(FULL, hi);
        if (lane == 0 && lo <= hi) {
            atomicMin(&range[0], lo);
            atomicMax(&range[1], hi);
        }
        __syncthreads();
        const uint32_t wlo = range[0], whi = range[1];
        const bool any = wlo <= whi;
        const uint32_t span = any ? whi - wlo + 1 : 0u;
        const uint32_t per = (span + blockDim.x - 1) / blockDim.x;
        // this thread's slice of words: starts = bits whose predecessor is clear, ends = bits whose successor is clear
        const uint32_t w0 = wlo + min(threadIdx.x * per, span), w1 = wlo + min((threadIdx.x + 1) * per, span);
        uint32_t ns = 0, ne = 0;
        for (uint32_t i = w0; i < w1; i++) {
            const uint32_t x = bm[i], prev = i ? bm[i - 1] >> 31 : 0u, next = bm[i + 1] & 1u;
            ns += __popc(x & ~((x << 1) | prev));
            if (MODE == 1) ne += __popc(x & ~((x >> 1) | (next << 31)));
        }
        uint32_t total = 0;
        uint32_t ks = block_excl_scan(ns, wt, &total);  // ends with a barrier: range[] and bm[] reads above are complete
        if (MODE == 0) {
            if (threadIdx.x == 0) count[u] = total;
        } else {
            uint32_t tot_e = 0;
            uint32_t ke = block_excl_scan(ne, wt, &tot_e);
            uint2 *o = out + out_ptr[u];
            for (uint32_t i = w0; i < w1; i++) {
                const uint32_t x = bm[i], prev = i ? bm[i - 1] >> 31 : 0u, next = bm[i + 1] & 1u;
                uint32_t st = x & ~((x << 1) | prev);
                while (st) {
                    const int b = __ffs(st) - 1;
                    st &= st - 1;
                    o[ks++].x = (i << 5) + (uint32_t)b;
                }
                uint32_t en = x & ~((x >> 1) | (next << 31));
                while (en) {
                    const int b = __ffs(en) - 1;
                    en &= en - 1;
                    o[ke++].y = (i << 5) + (uint32_t)b;  // position of the last bit, turned into a length below
                }
            }
            __syncthreads();
            for (uint32_t k = threadIdx.x; k < total; k += blockDim.x) o[k].y = o[k].y - o[k].x + 1u;
        }
        __syncthreads();
        // leave the bitmap zero for the next row
        for (uint32_t i = wlo + threadIdx.x; any && i <= whi; i += blockDim.x) bm[i] = 0u;
        if (threadIdx.x == 0) {
            range[0] = 0xffffffffu;
            range[1] = 0u;
        }
        __syncthreads();
    }
}

// rows handed over in x-major row order -> the same lists in y-major row order (row perm_y[u] = list of u): sizes, then copy
__global__ void k_row_sizes_permuted(int64_t n, const uint64_t *ptr, const uint32_t *perm_y, u64 *size_y) {
    const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (u < n) size_y[perm_y[u]] = ptr[u + 1] - ptr[u];
}
__global__ void k_rows_copy_permuted(int64_t n, const uint64_t *ptr, const uint2 *runs, const uint32_t *perm_y, const uint64_t *ptr_y,
                                     uint2 *runs_y) {
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    const uint64_t s = ptr[u], e = ptr[u + 1], d = ptr_y[perm_y[u]];
    for (uint64_t i = s + lane; i < e; i += 32) runs_y[d + (i - s)] = runs[i];
}

// per row the shorter of its x-major node list (list index = row) and its y-major one (list index = ymap ? ymap[row] : row)
__global__ void k_choose_lists(int64_t n, const uint64_t *xptr, const uint64_t *yptr, const uint32_t *ymap, u64 *size, uint8_t *isy) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int64_t yv = ymap ? (int64_t)ymap[v] : v;
    const u64 cx = xptr[v + 1] - xptr[v], cy = yptr[yv + 1] - yptr[yv];
    const bool y = cy * 10 < cx * 9;  // the y-major pyramid has to pay for its own build / down pass: take it for >= 10 %
    isy[v] = y ? 1 : 0;
    size[v] = y ? cy : cx;
}
// copies the chosen list of every row; nodes flagged NODE_COVERED go behind the others (flag stripped) and split[v] = the
// number of unflagged ("new") nodes at the front.  yptr == nullptr: every row takes its x list.
__global__ void k_copy_chosen(int64_t n, const uint64_t *xptr, const uint32_t *xnodes, const uint64_t *yptr, const uint32_t *ynodes,
                              const uint32_t *ymap, const uint8_t *isy, const uint64_t *optr, uint32_t *onodes, uint32_t *split) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= n) return;
    const int64_t yv = ymap ? (int64_t)ymap[v] : v;
    const uint32_t *src = (yptr && isy[v]) ? ynodes + yptr[yv] : xnodes + xptr[v];
    const uint64_t cnt = optr[v + 1] - optr[v], d = optr[v];
    const unsigned below = (1u << lane) - 1u;
    uint64_t at = 0;
    for (int pass = 0; pass < 2; pass++) {
        for (uint64_t i0 = 0; i0 < cnt; i0 += 32) {
            const uint64_t i = i0 + lane;
            const uint32_t id = i < cnt ? src[i] : 0u;
            const bool mine = i < cnt && ((id & NODE_COVERED) != 0u) == (pass == 1);
            const unsigned m = __ballot_sync(FULL, mine);
            if (mine) onodes[d + at + __popc(m & below)] = id & ~NODE_COVERED;
            at += __popc(m);
        }
        if (pass == 0 && split && lane == 0) split[v] = (uint32_t)at;
    }
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

int exclusive_sum_u64(vga_ctx *ctx, const u64 *in, u64 *out, int64_t count) {
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, in, out, (int)count, ctx->stream);
    DevBuf<unsigned char> tmp;
    VGA_TRY(tmp.alloc(tb + 16));
    VGA_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, in, out, (int)count, ctx->stream));
    ctx->timing.launches += 2;
    return VGA_OK;
}

// Level loop over a set of batches of W words.  nlev = number of levels with data (>= 1).
template <int W>
int run_levels(vga_ctx *ctx, BfsDev &d, int64_t nb, int radius, int bfs_mode, DevBuf<int32_t> *counts, int *lcap,
               int64_t counts_stride, u64 *work, StageTimer &mt, int *nlev_out) {
    cudaStream_t st = ctx->stream;
    Timing &tm = ctx->timing;
    const unsigned xblocks = (unsigned)std::min<int64_t>((d.n + TPB - 1) / TPB, 4096);
    int kmax = 0;
    while (kmax + 4 < d.pyr_levels) kmax += 3;
    int level = 0, nlev = 1;
    // VGA_LEVEL_TIMING=1 (development aid): device time of every kernel group of every level on stderr
    const bool lt = std::getenv("VGA_LEVEL_TIMING") != nullptr;
    cudaEvent_t le[6] = {};
    if (lt)
        for (auto &e : le) cudaEventCreate(&e);
    while (radius == -1 || level < radius) {
        dim3 grid(xblocks, (unsigned)nb);
        if (lt) cudaEventRecord(le[0], st);
        if (bfs_mode != 1 || level == 0) {
            if (d.f_split) {
                if (ctx->opt.bfs_delta_unroll >= 4)
                    k_push_delta<W, 4><<<grid, TPB, 0, st>>>(d);
                else if (ctx->opt.bfs_delta_unroll >= 2)
                    k_push_delta<W, 2><<<grid, TPB, 0, st>>>(d);
                else
                    k_push_delta<W, 1><<<grid, TPB, 0, st>>>(d);
            } else if constexpr (W >= 4 && W <= 16) {
                if (ctx->opt.bfs_coop && ctx->opt.bfs_push_unroll >= 4)
                    k_push_nodes_coop<W, W / 2, 4><<<grid, TPB, 0, st>>>(d);
                else if (ctx->opt.bfs_coop && ctx->opt.bfs_push_unroll >= 2)
                    k_push_nodes_coop<W, W / 2, 2><<<grid, TPB, 0, st>>>(d);
                else if (ctx->opt.bfs_coop)
                    k_push_nodes_coop<W, W / 2, 1><<<grid, TPB, 0, st>>>(d);
                else
                    k_push_nodes<W><<<grid, TPB, 0, st>>>(d);
            } else {
                k_push_nodes<W><<<grid, TPB, 0, st>>>(d);
            }
            tm.launches++;
            tm.main_launches++;
            if (lt) cudaEventRecord(le[1], st);
            for (int ymaj = 0; ymaj < (d.npyr_y ? 2 : 1); ymaj++)
                for (int k = kmax; k >= 0; k -= 3) {
                    const int64_t groups = (d.pyr_cnt[k] + 7) / 8;
                    dim3 pgrid((unsigned)std::min<int64_t>((groups + TPB - 1) / TPB, 4096), (unsigned)nb);
                    k_pyr_down<W><<<pgrid, TPB, 0, st>>>(d, k, ymaj);
                    tm.launches++;
                    tm.main_launches++;
                }
        }
        if (lt && !(bfs_mode != 1 || level == 0)) cudaEventRecord(le[1], st);
        if (lt) cudaEventRecord(le[2], st);
        if (bfs_mode != 0 && level > 0) {
            for (int ymaj = 0; ymaj < (d.pyr_y ? 2 : 1); ymaj++)
                for (int k = 0; k + 1 < d.pyr_levels; k += 3) {
                    const int64_t groups = (d.pyr_cnt[k] + 7) / 8;
                    dim3 pgrid((unsigned)std::min<int64_t>((groups + TPB - 1) / TPB, 4096), (unsigned)nb);
                    k_pyr_build<W><<<pgrid, TPB, 0, st>>>(d, k, ymaj);
                    tm.launches++;
                    tm.main_launches++;
                }
            if (lt) cudaEventRecord(le[3], st);
            if constexpr (W >= 4 && W <= 16) {
                if (ctx->opt.bfs_coop && ctx->opt.bfs_pull_unroll == 4)
                    k_pull_nodes_coop<W, W / 2, 4><<<grid, TPB, 0, st>>>(d);
                else if (ctx->opt.bfs_coop)
                    k_pull_nodes_coop<W, W / 2, 2><<<grid, TPB, 0, st>>>(d);
                else
                    k_pull_nodes<W><<<grid, TPB, 0, st>>>(d);
            } else {
                k_pull_nodes<W><<<grid, TPB, 0, st>>>(d);
            }
            tm.launches++;
            tm.main_launches++;
        }
        if (lt && !(bfs_mode != 0 && level > 0)) cudaEventRecord(le[3], st);
        if (lt) cudaEventRecord(le[4], st);
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
        // ~15 vertices per thread (one flush of the bit-sliced counters) as long as that leaves >= 16 CTAs per SM
        const int64_t ublocks_all = (d.n + TPB - 1) / TPB;
        const int64_t ublocks = std::min<int64_t>(ublocks_all, std::max<int64_t>((ublocks_all + UPD_FLUSH - 1) / UPD_FLUSH,
                                                                                 ((int64_t)ctx->sm_count * 16 + nb - 1) / nb));
        k_update<W><<<dim3((unsigned)ublocks, (unsigned)nb), TPB, 0, st>>>(d, counts ? counts->p + (size_t)(level + 1) * counts_stride : nullptr);
        // measured on a C5 slice of 16,384 sources with both list kinds (profiles/r2_pull_alpha_hybrid.log): level kernels
        // 75.9 / 69.1 / 70.0 / 82.6 ms for alpha 1 / 2 / 3 / 5 -- the in-lists shrink more than the out-lists and the down
        // pass runs over two pyramids; x-major lists only: alpha 1 is best.  With k_push_delta: alpha 1 and the visits weighted
        // (bfs_delta_weight, see k_decide)
        double lsum[NSTAT] = {};
        if (lt) {
            std::vector<u64> hs((size_t)nb * NSTAT);
            cudaMemcpyAsync(hs.data(), d.stats, sizeof(u64) * nb * NSTAT, cudaMemcpyDeviceToHost, st);
            cudaStreamSynchronize(st);
            for (int64_t i = 0; i < nb; i++)
                for (int k = 0; k < NSTAT; k++) lsum[k] += (double)hs[(size_t)(i * NSTAT + k)];
        }
        k_decide<<<blocks_for(nb, 128), 128, 0, st>>>(d, (int)nb, bfs_mode,
                                                       ctx->opt.pull_alpha > 0 ? ctx->opt.pull_alpha : ((d.npyr_y && !d.f_split) ? 2 : 1),
                                                       ctx->opt.pull_beta, std::max<int64_t>(1, ctx->opt.bfs_delta_weight), work);
        tm.launches += 2;
        tm.main_launches += 2;
        if (lt) cudaEventRecord(le[5], st);
        int h_any = 0;
        VGA_CUDA(cudaMemcpyAsync(&h_any, d.any, sizeof(int), cudaMemcpyDeviceToHost, st));
        VGA_CUDA(cudaStreamSynchronize(st));
        if (lt) {
            float ms[5];
            for (int i = 0; i < 5; i++) cudaEventElapsedTime(&ms[i], le[i], le[i + 1]);
            std::vector<int> hm((size_t)nb), ha((size_t)nb);
            cudaMemcpy(hm.data(), d.mode, sizeof(int) * nb, cudaMemcpyDeviceToHost);
            cudaMemcpy(ha.data(), d.active, sizeof(int) * nb, cudaMemcpyDeviceToHost);
            int np = 0, nl = 0;
            for (int64_t i = 0; i < nb; i++)
                if (ha[(size_t)i]) (hm[(size_t)i] ? nl : np)++;
            fprintf(stderr, "[level %d] push %.3f down %.3f build %.3f pull %.3f update %.3f ms; next step: %d batches push, %d pull; "
                    "per batch: new vertices %.0f, open %.0f, out-row nodes %.0f (to visit %.0f), in-row nodes of open %.0f\n",
                    level, ms[0], ms[1], ms[2], ms[3], ms[4], np, nl, lsum[2] / nb, lsum[3] / nb, lsum[4] / nb, lsum[5] / nb, lsum[1] / nb);
        }
        if (h_any & 2) nlev = level + 2;
        if (!(h_any & 1)) break;
        level++;
    }
    if (lt)
        for (auto &e : le) cudaEventDestroy(e);
    *nlev_out = nlev;
    return VGA_OK;
}

// Spatially coherent order of a list of sources for batching (see header comment).  order[i] = ordinal of the i-th
// source; positions[i] = index of that source in the caller's list.
int source_order(vga_ctx *ctx, vga_graph *g, const int64_t *sources, int64_t src_begin, int64_t nsrc, std::vector<int32_t> &order,
                 std::vector<int64_t> &position) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    order.resize((size_t)nsrc);
    position.resize((size_t)nsrc);
    for (int64_t i = 0; i < nsrc; i++) {
        order[(size_t)i] = (int32_t)(sources ? sources[i] : src_begin + i);
        position[(size_t)i] = i;
    }
    if (ctx->opt.bfs_order == 0 || (int64_t)g->h_refs.size() < n || nsrc == 0) return VGA_OK;
    const bool whole = !sources && src_begin == 0 && nsrc == n;
    if (whole && (int64_t)g->h_order.size() == n) {  // cached
        order = g->h_order;
        for (int64_t i = 0; i < nsrc; i++) position[(size_t)i] = order[(size_t)i];
        return VGA_OK;
    }
    // key = Morton code of the 8x8 tile, then x and y inside the tile (30 bits), list index in the low word; a stable
    // LSD radix sort over the key bytes
    std::vector<uint64_t> key((size_t)nsrc), tmp((size_t)nsrc);
    for (int64_t i = 0; i < nsrc; i++) {
        uint32_t r = (uint32_t)g->h_refs[(size_t)order[(size_t)i]];
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
    std::vector<int32_t> sorted((size_t)nsrc);
    std::vector<int64_t> spos((size_t)nsrc);
    for (int64_t i = 0; i < nsrc; i++) {
        const int64_t li = (int64_t)(key[(size_t)i] & 0xffffffffu);
        sorted[(size_t)i] = order[(size_t)li];
        spos[(size_t)i] = li;
    }
    order.swap(sorted);
    position.swap(spos);
    if (ctx->opt.bfs_order < 2 || !g->has_fwd_runs) return VGA_OK;
    // Wall-respecting clusters: a raw 8x8 tile often straddles a wall, which puts cells of two rooms (very different
    // level structure) into one batch.  Grow each cluster as a flood of up to 64 cells over the direct-neighbour
    // links (the 8 grid neighbours that are in the cell's row, i.e. actually visible), seeded in Morton order.
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
        k_neighbour_bits<<<blocks_for(n, 256), 256, 0, st>>>(n, g->f_runptr.p, g->f_runs.p, d_refs.p, d_ord.p, (int)cols, (int)rows,
                                                            d_gc.p);
        ctx->timing.launches++;
        VGA_CUDA(cudaMemcpyAsync(gc.data(), d_gc.p, (size_t)n, cudaMemcpyDeviceToHost, st));
        VGA_CUDA(cudaStreamSynchronize(st));
    }
    static const int dx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    static const int dy[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    // state per vertex: -1 not in the list, else index of its (first) list entry; taken marks
    std::vector<int64_t> listed((size_t)n, -1);
    for (int64_t i = nsrc - 1; i >= 0; i--) listed[(size_t)order[(size_t)i]] = i;
    std::vector<uint8_t> taken((size_t)n, 0);
    std::vector<int32_t> clustered, queue;
    std::vector<int64_t> cpos;
    clustered.reserve((size_t)nsrc);
    cpos.reserve((size_t)nsrc);
    bool duplicates = false;
    for (int64_t i = 0; i < nsrc; i++) {
        const int32_t seed = order[(size_t)i];
        if (listed[(size_t)seed] != i) {
            duplicates = true;  // a source listed twice: keep the plain Morton order
            break;
        }
        if (taken[(size_t)seed]) continue;
        queue.clear();
        queue.push_back(seed);
        taken[(size_t)seed] = 1;
        size_t head = 0;
        int count = 0;
        while (head < queue.size() && count < 64) {
            const int32_t v = queue[head++];
            clustered.push_back(v);
            cpos.push_back(position[(size_t)listed[(size_t)v]]);
            count++;
            const uint32_t r = (uint32_t)g->h_refs[(size_t)v];
            const int x = (int)(r >> 16), y = (int)(r & 0xffff);
            for (int k = 0; k < 8; k++) {
                if (!(gc[(size_t)v] & (1 << k))) continue;
                const int nx = x + dx[k], ny = y + dy[k];
                if (nx < 0 || nx >= cols || ny < 0 || ny >= rows) continue;
                const int32_t w = ord_of[(size_t)((int64_t)nx * rows + ny)];
                if (w < 0 || listed[(size_t)w] < 0 || taken[(size_t)w]) continue;
                taken[(size_t)w] = 1;
                queue.push_back(w);
            }
        }
        for (; head < queue.size(); head++) taken[(size_t)queue[head]] = 0;  // not placed: free again
    }
    if (!duplicates && (int64_t)clustered.size() == nsrc) {
        order.swap(clustered);
        position.swap(cpos);
    }
    if (whole) g->h_order = order;
    return VGA_OK;
}

template <int W>
int run_global_w(vga_ctx *ctx, vga_graph *g, int radius, const int64_t *sources, int64_t src_begin, int64_t nsrc,
                 int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    Timing &tm = ctx->timing;
    const int bfs_mode = (int)ctx->opt.bfs_mode;
    tm.batch_words = W;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);
    StageTimer dt(ctx, 4, &tm.d2h_ms);
    StageTimer pt(ctx, 6, &tm.prep_ms);

    const bool dbg = std::getenv("VGA_DEBUG_TIMING") != nullptr;
    auto wall = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double w0 = wall();
    pt.start();
    VGA_TRY(ensure_bfs_lists(ctx, g, bfs_mode != 0));
    pt.stop();
    const PyrLayout pl = pyr_layout(n);
    const double w1 = wall();

    std::vector<int32_t> order;
    std::vector<int64_t> position;
    VGA_TRY(source_order(ctx, g, sources, src_begin, nsrc, order, position));
    const double w2 = wall();
    DevBuf<int32_t> d_order;
    VGA_TRY(d_order.alloc((size_t)nsrc));
    VGA_CUDA(cudaMemcpyAsync(d_order.p, order.data(), sizeof(int32_t) * nsrc, cudaMemcpyHostToDevice, st));

    const int64_t nwords = (nsrc + 63) / 64;      // 64-source words
    const int64_t nbatch = (nwords + W - 1) / W;  // batches of W words
    // chunk size: per (batch, vertex) 3 state words + 2 pyramids (~n inner nodes each) = 5*W words + the histogram;
    // below ~40% of free memory and 48 GB
    size_t free_b = 0, total_b = 0;
    VGA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const int64_t budget = (int64_t)std::min<size_t>((size_t)(free_b * 0.4), (size_t)48 << 30);
    int64_t chunk = ctx->opt.bfs_chunk > 0 ? std::max<int64_t>(1, ctx->opt.bfs_chunk / W)
                                           : std::max<int64_t>(1, budget / ((g->has_y ? 60 : 44) * W * std::max<int64_t>(n, 1)));
    chunk = std::min<int64_t>(chunk, nbatch);
    chunk = std::min<int64_t>(chunk, 65535);

    DevBuf<u64> valid, stats, work;
    DevBuf<int> active, mode, any;
    DevBuf<int32_t> counts;
    int lcap = 16;
    // the big state arrays live in the context's workspace and are reused by later calls
    u64 *visited = nullptr, *frontier = nullptr, *next = nullptr, *pyr_p = nullptr, *npyr_p = nullptr;
    VGA_TRY(ctx->ws.get("bfs_visited", sizeof(u64) * (size_t)chunk * n * W, (void **)&visited));
    VGA_TRY(ctx->ws.get("bfs_frontier", sizeof(u64) * (size_t)chunk * n * W, (void **)&frontier));
    VGA_TRY(ctx->ws.get("bfs_next", sizeof(u64) * (size_t)chunk * n * W, (void **)&next));
    if (bfs_mode != 0) VGA_TRY(ctx->ws.get("bfs_pyr", sizeof(u64) * (size_t)chunk * (size_t)std::max<int64_t>(pl.total, 1) * W, (void **)&pyr_p));
    VGA_TRY(ctx->ws.get("bfs_npyr", sizeof(u64) * (size_t)chunk * (size_t)std::max<int64_t>(pl.total, 1) * W, (void **)&npyr_p));
    // all zero between levels: the down pass clears what the push wrote
    VGA_CUDA(cudaMemsetAsync(npyr_p, 0, sizeof(u64) * (size_t)chunk * (size_t)std::max<int64_t>(pl.total, 1) * W, st));
    u64 *pyr_y = nullptr, *npyr_y = nullptr;
    if (g->has_y) {
        if (bfs_mode != 0) VGA_TRY(ctx->ws.get("bfs_pyr_y", sizeof(u64) * (size_t)chunk * (size_t)std::max<int64_t>(pl.total, 1) * W, (void **)&pyr_y));
        VGA_TRY(ctx->ws.get("bfs_npyr_y", sizeof(u64) * (size_t)chunk * (size_t)std::max<int64_t>(pl.total, 1) * W, (void **)&npyr_y));
        VGA_CUDA(cudaMemsetAsync(npyr_y, 0, sizeof(u64) * (size_t)chunk * (size_t)std::max<int64_t>(pl.total, 1) * W, st));
    }
    VGA_TRY(valid.alloc((size_t)chunk * W));
    VGA_TRY(stats.alloc((size_t)chunk * NSTAT));
    VGA_TRY(work.alloc_zero(4, st));
    VGA_TRY(active.alloc((size_t)chunk));
    VGA_TRY(mode.alloc((size_t)chunk));
    VGA_TRY(any.alloc(1));
    VGA_TRY(counts.alloc((size_t)lcap * chunk * W * 64));

    BfsDev d = {};  // every pointer of an unused option must be null
    d.n = n;
    d.f_ptr = g->f_nodeptr.p;
    d.f_nodes = g->f_nodes.p;
    d.f_split = (ctx->opt.bfs_delta && g->f_split.p) ? g->f_split.p : nullptr;
    d.t_ptr = bfs_mode != 0 ? g->t_nodeptr.p : nullptr;
    d.t_nodes = bfs_mode != 0 ? g->t_nodes.p : nullptr;
    d.rowptr = !g->runs_only ? g->rowptr.p : nullptr;
    d.deg = g->runs_only ? g->deg.p : nullptr;
    d.visited = visited;
    d.frontier = frontier;
    d.next = next;
    d.pyr = pyr_p;
    d.npyr = npyr_p;
    d.pyr_y = pyr_y;
    d.npyr_y = npyr_y;
    d.perm_x = g->has_y ? g->perm_x.p : nullptr;
    d.f_isy = g->has_y ? g->f_isy.p : nullptr;
    d.t_isy = (g->has_y && bfs_mode != 0) ? g->t_isy.p : nullptr;
    d.pyr_total = pl.total;
    d.valid = valid.p;
    d.active = active.p;
    d.mode = mode.p;
    d.stats = stats.p;
    d.any = any.p;
    d.noexpand = (radius != -1 && g->noexpand.p) ? g->noexpand.p : nullptr;
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
    const double w3 = wall();
    double w_unpack = 0.0;

    for (int64_t b0 = 0; b0 < nbatch; b0 += chunk) {
        const int64_t cb = std::min<int64_t>(chunk, nbatch - b0);
        const int64_t first = b0 * W * 64;  // index into the ordered source list
        const int64_t cs = std::min<int64_t>(cb * W * 64, nsrc - first);
        if (ctx->cancel && ctx->cancel(ctx->user)) {
            set_error("cancelled");
            return VGA_ERR_CANCELLED;
        }
        kt.start();
        {
            const int64_t words_here = (cs + 63) / 64;
            h_valid.assign((size_t)(cb * W), 0ULL);
            for (int64_t wi = 0; wi < words_here; wi++) h_valid[(size_t)wi] = ~0ULL;
            if (cs & 63) h_valid[(size_t)words_here - 1] = (1ULL << (cs & 63)) - 1ULL;
        }
        ones.assign((size_t)cb, 1);
        VGA_CUDA(cudaMemcpyAsync(valid.p, h_valid.data(), sizeof(u64) * cb * W, cudaMemcpyHostToDevice, st));
        VGA_CUDA(cudaMemcpyAsync(active.p, ones.data(), sizeof(int) * cb, cudaMemcpyHostToDevice, st));
        VGA_CUDA(cudaMemsetAsync(visited, 0, sizeof(u64) * (size_t)cb * n * W, st));
        VGA_CUDA(cudaMemsetAsync(frontier, 0, sizeof(u64) * (size_t)cb * n * W, st));
        VGA_CUDA(cudaMemsetAsync(next, 0, sizeof(u64) * (size_t)cb * n * W, st));
        VGA_CUDA(cudaMemsetAsync(stats.p, 0, sizeof(u64) * (size_t)cb * NSTAT, st));
        VGA_CUDA(cudaMemsetAsync(mode.p, 0, sizeof(int) * (size_t)cb, st));  // level 0 always pushes
        VGA_CUDA(cudaMemsetAsync(counts.p, 0, sizeof(int32_t) * (size_t)lcap * cstride, st));
        VGA_CUDA(cudaStreamSynchronize(st));  // the host staging vectors are reused by the next chunk
        k_init<W><<<blocks_for(cs, 256), 256, 0, st>>>(d, d_order.p + first, cs);
        tm.launches++;
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
        const double wu0 = wall();
        for (int64_t i = 0; i < cs; i++) {
            const int64_t o = position[(size_t)(first + i)];
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
        w_unpack += wall() - wu0;
        if (ctx->progress) ctx->progress(ctx->user, std::min<int64_t>(nsrc, (b0 + cb) * W * 64), nsrc);
    }
    if (levels_used) *levels_used = deepest;
    if (dbg)
        fprintf(stderr, "[vga_global] W=%d chunk=%lld: row lists %.2f ms, source order %.2f ms, state set-up %.2f ms, chunks %.2f ms "
                "(level kernels %.2f, result unpack on the host %.2f)\n", W, (long long)chunk, w1 - w0, w2 - w1, w3 - w2, wall() - w3,
                tm.main_kernel_ms, w_unpack);
    {
        u64 hw[4] = {0, 0, 0, 0};
        VGA_CUDA(cudaMemcpy(hw, work.p, sizeof(hw), cudaMemcpyDeviceToHost));
        // rows of the sources themselves (level 0 expands them): entries and nodes
        double src_edges = 0, src_nodes = 0;
        {
            std::vector<uint64_t> fp;
            if (!sources) {
                uint64_t a[2] = {0, 0}, b[2] = {0, 0};
                VGA_CUDA(cudaMemcpy(&a[0], g->f_nodeptr.p + src_begin, sizeof(uint64_t), cudaMemcpyDeviceToHost));
                VGA_CUDA(cudaMemcpy(&a[1], g->f_nodeptr.p + src_begin + nsrc, sizeof(uint64_t), cudaMemcpyDeviceToHost));
                src_nodes = (double)(a[1] - a[0]);
                if (!g->runs_only) {
                    VGA_CUDA(cudaMemcpy(&b[0], g->rowptr.p + src_begin, sizeof(uint64_t), cudaMemcpyDeviceToHost));
                    VGA_CUDA(cudaMemcpy(&b[1], g->rowptr.p + src_begin + nsrc, sizeof(uint64_t), cudaMemcpyDeviceToHost));
                    src_edges = (double)(b[1] - b[0]);
                }
            } else {
                fp.resize((size_t)n + 1);
                VGA_CUDA(cudaMemcpy(fp.data(), g->f_nodeptr.p, sizeof(uint64_t) * (n + 1), cudaMemcpyDeviceToHost));
                for (int64_t i = 0; i < nsrc; i++) src_nodes += (double)(fp[(size_t)sources[i] + 1] - fp[(size_t)sources[i]]);
                if (!g->runs_only) {
                    VGA_CUDA(cudaMemcpy(fp.data(), g->rowptr.p, sizeof(uint64_t) * (n + 1), cudaMemcpyDeviceToHost));
                    for (int64_t i = 0; i < nsrc; i++) src_edges += (double)(fp[(size_t)sources[i] + 1] - fp[(size_t)sources[i]]);
                }
            }
        }
        // Algorithmic bytes (SURVEY.md §8d with B = 64*W sources per batch), per batch and level: rows of the expanding
        // vertices + one frontier vector (B/8 bytes) read per expanding vertex + one visited/next read-modify-write (2
        // vectors) per newly reached vertex; level 0 expands the sources.  algo_bytes: rows in the format the kernels
        // read (4 bytes per pyramid node id); algo_bytes_csr: the same with 4-byte CSR entries (round 1's format).
        const double vec = 8.0 * W;
        const double state = vec * ((double)hw[1] + (double)nsrc) + 2.0 * vec * (double)hw[1];
        tm.algo_bytes = 4.0 * ((double)hw[3] + src_nodes) + state;
        tm.algo_bytes_csr = 4.0 * ((double)hw[0] + src_edges) + state;
    }
    if (dist && deepest > max_levels) {
        set_error("vga_global: level histogram needs " + std::to_string(deepest) + " columns");
        return VGA_ERR_CAPACITY;
    }
    return VGA_OK;
}

// Run-length form of sorted rows: one run per maximal stretch of consecutive ordinals (broken at the first ghost column).
int build_runs(vga_ctx *ctx, int64_t n /*rows*/, uint32_t brk /*filled cells*/, const uint64_t *rowptr, const uint32_t *arr, int shift,
               DevBuf<uint64_t> &runptr, DevBuf<uint2> &runs, int64_t *nruns) {
    cudaStream_t st = ctx->stream;
    DevBuf<u64> count;
    DevBuf<uint32_t> first_off;
    VGA_TRY(count.alloc_zero((size_t)n + 1, st));
    VGA_TRY(runptr.alloc((size_t)n + 1));
    if (n > 0) {
        k_count_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(rowptr, arr, shift, brk, n, count.p);
        ctx->timing.launches++;
    }
    VGA_TRY(exclusive_sum_u64(ctx, count.p, (u64 *)runptr.p, n + 1));
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, runptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    *nruns = (int64_t)total;
    VGA_TRY(first_off.alloc((size_t)total + 1));
    VGA_TRY(runs.alloc((size_t)total + 1));
    if (n > 0) {
        k_mark_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(rowptr, arr, shift, brk, n, runptr.p, first_off.p);
        k_emit_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(rowptr, arr, shift, n, runptr.p, first_off.p, runs.p);
        ctx->timing.launches += 2;
    }
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// in-rows as runs from the out-rows' runs (see k_trans_events); rows and columns share one numbering of the n vertices
int transpose_runs(vga_ctx *ctx, int64_t n, const DevBuf<uint64_t> &runptr, const DevBuf<uint2> &runs, DevBuf<uint64_t> &t_runptr,
                   DevBuf<uint2> &t_runs, int64_t *t_nruns) {
    cudaStream_t st = ctx->stream;
    DevBuf<u64> count;
    DevBuf<uint32_t> cur_s, cur_e, t_start, t_end;
    VGA_TRY(count.alloc_zero((size_t)n + 1, st));
    VGA_TRY(t_runptr.alloc((size_t)n + 1));
    if (n > 0) {
        k_trans_events<0><<<blocks_for(n * 32, 256), 256, 0, st>>>(n, runptr.p, runs.p, count.p, nullptr, nullptr, nullptr, nullptr, nullptr);
        ctx->timing.launches++;
    }
    VGA_TRY(exclusive_sum_u64(ctx, count.p, (u64 *)t_runptr.p, n + 1));
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, t_runptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    *t_nruns = (int64_t)total;
    VGA_TRY(cur_s.alloc_zero((size_t)n + 1, st));
    VGA_TRY(cur_e.alloc_zero((size_t)n + 1, st));
    VGA_TRY(t_start.alloc((size_t)total + 1));
    VGA_TRY(t_end.alloc((size_t)total + 1));
    VGA_TRY(t_runs.alloc((size_t)total + 1));
    if (n > 0) {
        k_trans_events<1><<<blocks_for(n * 32, 256), 256, 0, st>>>(n, runptr.p, runs.p, nullptr, t_runptr.p, cur_s.p, cur_e.p, t_start.p,
                                                                  t_end.p);
        k_trans_pair<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, t_runptr.p, t_start.p, t_end.p, t_runs.p);
        ctx->timing.launches += 2;
    }
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// node-id lists from runs: per-run node counts -> exclusive scan -> ids written at the run's offset
// cover: flag the nodes whose cells the previous row holds as well (k_emit_nodes)
int build_nodes(vga_ctx *ctx, int64_t n, const DevBuf<uint64_t> &runptr, const DevBuf<uint2> &runs, int64_t nruns,
                DevBuf<uint64_t> &nodeptr, DevBuf<uint32_t> &nodes, int64_t *nnodes, const uint32_t *leaf = nullptr, bool cover = false) {
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
    VGA_TRY(cost.alloc_zero((size_t)nruns + 1, st));
    VGA_TRY(off.alloc((size_t)nruns + 1));
    VGA_TRY(nodeptr.alloc((size_t)n + 1));
    if (nruns > 0) {
        k_run_costs<<<blocks_for(nruns, 256), 256, 0, st>>>(runs.p, nruns, (uint32_t)n, cost.p);
        ctx->timing.launches++;
    }
    VGA_TRY(exclusive_sum_u64(ctx, cost.p, off.p, nruns + 1));
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, off.p + nruns, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    *nnodes = (int64_t)total;
    VGA_TRY(nodes.alloc((size_t)total + 1));
    if (nruns > 0) {
        k_emit_nodes<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, runptr.p, runs.p, off.p, (uint32_t)n, d0, leaf, cover ? 1 : 0, nodes.p);
        ctx->timing.launches++;
    }
    k_row_node_offsets<<<blocks_for(n + 1, 256), 256, 0, st>>>(n, runptr.p, off.p, nodeptr.p);
    ctx->timing.launches++;
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// y-major rank of every vertex (perm_y) and its inverse (perm_x); needs the cells' coordinates
int ensure_perm(vga_ctx *ctx, vga_graph *g) {
    if (g->perm_y.p) return VGA_OK;
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    int maxx = 0, maxy = 0;
    for (int64_t v = 0; v < n; v++) {
        const uint32_t r = (uint32_t)g->h_refs[(size_t)v];
        maxx = std::max(maxx, (int)(r >> 16));
        maxy = std::max(maxy, (int)(r & 0xffff));
    }
    const int64_t cols = maxx + 1, rows = maxy + 1;
    DevBuf<int32_t> d_refs;
    DevBuf<u64> flag, scan;
    VGA_TRY(d_refs.alloc((size_t)n));
    VGA_TRY(flag.alloc_zero((size_t)(cols * rows) + 1, st));
    VGA_TRY(scan.alloc((size_t)(cols * rows) + 1));
    VGA_TRY(g->perm_y.alloc((size_t)n));
    VGA_TRY(g->perm_x.alloc((size_t)n));
    VGA_CUDA(cudaMemcpyAsync(d_refs.p, g->h_refs.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    k_mark_cells_ymajor<<<blocks_for(n, 256), 256, 0, st>>>(n, d_refs.p, (int)cols, flag.p);
    VGA_TRY(exclusive_sum_u64(ctx, flag.p, scan.p, cols * rows + 1));
    k_perm_from_scan<<<blocks_for(n, 256), 256, 0, st>>>(n, d_refs.p, (int)cols, scan.p, g->perm_y.p, g->perm_x.p);
    ctx->timing.launches += 2;
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// the y-major run lists of the out-rows, in x-major row order (k_yruns)
int build_yruns(vga_ctx *ctx, vga_graph *g, DevBuf<uint64_t> &yptr, DevBuf<uint2> &yruns, int64_t *nruns) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    DevBuf<u64> count;
    VGA_TRY(count.alloc_zero((size_t)n + 1, st));
    VGA_TRY(yptr.alloc((size_t)n + 1));
    const size_t words = (size_t)((n + 31) >> 5);
    const int threads = words >= 8192 ? 1024 : 256;
    const size_t smem = sizeof(uint32_t) * (words + 1 + 33 + 2 + 2 * (size_t)threads);
    if (smem > ctx->smem_optin) {
        set_error("y-major rows: the vertex bitmap exceeds shared memory");
        return VGA_ERR_UNSUPPORTED;
    }
    VGA_CUDA(cudaFuncSetAttribute(k_yruns<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VGA_CUDA(cudaFuncSetAttribute(k_yruns<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>((size_t)(2048 / threads), (ctx->smem_optin - 1024) / std::max<size_t>(smem, 1)));
    const unsigned blocks = (unsigned)std::min<int64_t>(std::max<int64_t>(n, 1), (int64_t)ctx->sm_count * per_sm);
    k_yruns<0><<<blocks, threads, smem, st>>>(n, g->f_runptr.p, g->f_runs.p, g->perm_y.p, count.p, nullptr, nullptr);
    ctx->timing.launches++;
    VGA_TRY(exclusive_sum_u64(ctx, count.p, (u64 *)yptr.p, n + 1));
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, yptr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    *nruns = (int64_t)total;
    VGA_TRY(yruns.alloc((size_t)total + 1));
    k_yruns<1><<<blocks, threads, smem, st>>>(n, g->f_runptr.p, g->f_runs.p, g->perm_y.p, nullptr, yptr.p, yruns.p);
    ctx->timing.launches++;
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// final lists of one direction: per row the shorter of the x-major list (index = row) and the y-major one (index =
// ymap ? ymap[row] : row)
int choose_lists(vga_ctx *ctx, int64_t n, const DevBuf<uint64_t> &xptr, const DevBuf<uint32_t> &xnodes, const DevBuf<uint64_t> &yptr,
                 const DevBuf<uint32_t> &ynodes, const uint32_t *ymap, DevBuf<uint64_t> &optr, DevBuf<uint32_t> &onodes,
                 DevBuf<uint8_t> &isy, int64_t *nnodes, uint32_t *split = nullptr) {
    cudaStream_t st = ctx->stream;
    DevBuf<u64> size;
    VGA_TRY(size.alloc_zero((size_t)n + 1, st));
    VGA_TRY(isy.alloc((size_t)n + 1));
    VGA_TRY(optr.alloc((size_t)n + 1));
    k_choose_lists<<<blocks_for(n, 256), 256, 0, st>>>(n, xptr.p, yptr.p, ymap, size.p, isy.p);
    ctx->timing.launches++;
    VGA_TRY(exclusive_sum_u64(ctx, size.p, (u64 *)optr.p, n + 1));
    uint64_t total = 0;
    VGA_CUDA(cudaMemcpyAsync(&total, optr.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    *nnodes = (int64_t)total;
    VGA_TRY(onodes.alloc((size_t)total + 1));
    k_copy_chosen<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, xptr.p, xnodes.p, yptr.p, ynodes.p, ymap, isy.p, optr.p, onodes.p, split);
    ctx->timing.launches++;
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// one kind of list only: the flagged nodes of every row moved behind the others (k_copy_chosen), split[v] = new nodes
int partition_lists(vga_ctx *ctx, int64_t n, const DevBuf<uint64_t> &ptr, const DevBuf<uint32_t> &nodes, int64_t nnodes,
                    DevBuf<uint32_t> &onodes, uint32_t *split) {
    cudaStream_t st = ctx->stream;
    VGA_TRY(onodes.alloc((size_t)nnodes + 1));
    k_copy_chosen<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, ptr.p, nodes.p, nullptr, nullptr, nullptr, nullptr, ptr.p, onodes.p, split);
    ctx->timing.launches++;
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

}  // namespace

int batch_source_order(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, std::vector<int32_t> &order) {
    std::vector<int64_t> pos;
    VGA_TRY(ensure_fwd_runs(ctx, g));
    return source_order(ctx, g, nullptr, src_begin, src_end - src_begin, order, pos);
}

int batch_source_order_list(vga_ctx *ctx, vga_graph *g, const int64_t *sources, int64_t nsrc, std::vector<int32_t> &order) {
    std::vector<int64_t> pos;
    VGA_TRY(ensure_fwd_runs(ctx, g));
    return source_order(ctx, g, sources, 0, nsrc, order, pos);
}

// out-rows as runs (rows are kept sorted by column; packed entries col << 6 | ...)
int ensure_fwd_runs(vga_ctx *ctx, vga_graph *g) {
    if (g->has_fwd_runs) return VGA_OK;
    if (g->src_begin != 0 || g->src_end != g->n) {
        set_error("the graph must hold the rows of all cells (gather the shards first)");
        return VGA_ERR_INVALID;
    }
    return shard_runs(ctx, g);
}

// the same for the rows [src_begin, src_end) a shard holds (what a multi-GPU run exchanges)
int shard_runs(vga_ctx *ctx, vga_graph *g) {
    if (g->has_fwd_runs || g->has_shard_runs) return VGA_OK;
    VGA_TRY(build_runs(ctx, g->src_end - g->src_begin, (uint32_t)g->n, g->rowptr.p, g->adj.p, 6, g->f_runptr, g->f_runs, &g->f_nruns));
    if (g->src_begin == 0 && g->src_end == g->n)
        g->has_fwd_runs = true;
    else
        g->has_shard_runs = true;
    return VGA_OK;
}

// entries per row of the rows this graph holds (u32), for a rank that only receives runs
int row_degrees(vga_ctx *ctx, vga_graph *g) {
    if (g->deg.p) return VGA_OK;
    const int64_t rows = g->src_end - g->src_begin;
    VGA_TRY(g->deg.alloc((size_t)rows + 1));
    if (rows > 0) {
        k_row_lengths<<<blocks_for(rows, 256), 256, 0, ctx->stream>>>(rows, g->rowptr.p, g->deg.p);
        ctx->timing.launches++;
    }
    VGA_CUDA(cudaStreamSynchronize(ctx->stream));
    return VGA_OK;
}

int ensure_bfs_lists(vga_ctx *ctx, vga_graph *g, bool transposed) {
    const int64_t n = g->n;
    if (n + 2 * (int64_t)pyr_layout(n).total >= ((int64_t)1 << 32) || pyr_layout(n).levels > PYR_LEVELS_DEV) {
        set_error("too many vertices for 32-bit pyramid node ids");
        return VGA_ERR_UNSUPPORTED;
    }
    VGA_TRY(ensure_fwd_runs(ctx, g));
    // y-major alternative lists need the cells' coordinates and a vertex bitmap that fits shared memory
    const bool want_y = ctx->opt.bfs_hybrid >= 2 ||
                        (ctx->opt.bfs_hybrid == 1 && (ctx->bfs_sources_hint < 0 || ctx->bfs_sources_hint * 4 >= n));
    const bool hybrid = want_y && n > 1 && (int64_t)g->h_refs.size() >= n &&
                        sizeof(uint32_t) * ((size_t)((n + 31) >> 5) + 1 + 35 + 2048) <= ctx->smem_optin;
    if (!g->has_f_nodes) {
        // out-rows: the nodes whose cells the previous row holds too go to the back of the list (f_split = the others)
        VGA_TRY(g->f_split.alloc((size_t)n + 1));
        if (!hybrid) {
            DevBuf<uint32_t> xnodes;
            VGA_TRY(build_nodes(ctx, n, g->f_runptr, g->f_runs, g->f_nruns, g->f_nodeptr, xnodes, &g->f_nnodes, nullptr, true));
            VGA_TRY(partition_lists(ctx, n, g->f_nodeptr, xnodes, g->f_nnodes, g->f_nodes, g->f_split.p));
        } else {
            DevBuf<uint64_t> xptr, yptr;
            DevBuf<uint32_t> xnodes, ynodes;
            int64_t nx = 0, ny = 0;
            VGA_TRY(build_nodes(ctx, n, g->f_runptr, g->f_runs, g->f_nruns, xptr, xnodes, &nx, nullptr, true));
            VGA_TRY(ensure_perm(ctx, g));
            VGA_TRY(build_yruns(ctx, g, g->fy_runptr, g->fy_runs, &g->fy_nruns));
            VGA_TRY(build_nodes(ctx, n, g->fy_runptr, g->fy_runs, g->fy_nruns, yptr, ynodes, &ny, g->perm_x.p, true));
            VGA_TRY(choose_lists(ctx, n, xptr, xnodes, yptr, ynodes, nullptr, g->f_nodeptr, g->f_nodes, g->f_isy, &g->f_nnodes,
                                 g->f_split.p));
            g->has_y = true;
        }
        g->has_f_nodes = true;
    }
    if (transposed && !g->has_t_nodes) {
        if (!g->has_runs) {
            VGA_TRY(transpose_runs(ctx, n, g->f_runptr, g->f_runs, g->t_runptr, g->t_runs, &g->t_nruns));
            g->has_runs = true;
        }
        if (!g->has_y) {
            VGA_TRY(build_nodes(ctx, n, g->t_runptr, g->t_runs, g->t_nruns, g->t_nodeptr, g->t_nodes, &g->t_nnodes));
        } else {
            cudaStream_t st = ctx->stream;
            DevBuf<uint64_t> xptr, yptr, yy_ptr, ty_ptr;
            DevBuf<uint32_t> xnodes, ynodes;
            DevBuf<uint2> yy_runs, ty_runs;
            DevBuf<u64> size_y;
            int64_t nx = 0, ny = 0, ty_nruns = 0;
            VGA_TRY(build_nodes(ctx, n, g->t_runptr, g->t_runs, g->t_nruns, xptr, xnodes, &nx));
            // the y-major out-rows in y-major ROW order, transposed in that numbering: in-rows of the y-major world
            VGA_TRY(size_y.alloc_zero((size_t)n + 1, st));
            VGA_TRY(yy_ptr.alloc((size_t)n + 1));
            VGA_TRY(yy_runs.alloc((size_t)g->fy_nruns + 1));
            k_row_sizes_permuted<<<blocks_for(n, 256), 256, 0, st>>>(n, g->fy_runptr.p, g->perm_y.p, size_y.p);
            VGA_TRY(exclusive_sum_u64(ctx, size_y.p, (u64 *)yy_ptr.p, n + 1));
            k_rows_copy_permuted<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, g->fy_runptr.p, g->fy_runs.p, g->perm_y.p, yy_ptr.p, yy_runs.p);
            ctx->timing.launches += 2;
            VGA_TRY(transpose_runs(ctx, n, yy_ptr, yy_runs, ty_ptr, ty_runs, &ty_nruns));
            VGA_TRY(build_nodes(ctx, n, ty_ptr, ty_runs, ty_nruns, yptr, ynodes, &ny, g->perm_x.p));
            // the in-row of vertex v in the y-major world is row perm_y[v]
            VGA_TRY(choose_lists(ctx, n, xptr, xnodes, yptr, ynodes, g->perm_y.p, g->t_nodeptr, g->t_nodes, g->t_isy, &g->t_nnodes));
            g->fy_runs.release();  // only needed for this derivation
            g->fy_runptr.release();
        }
        g->has_t_nodes = true;
    }
    return VGA_OK;
}

int run_global(vga_ctx *ctx, vga_graph *g, int radius, const int64_t *sources, int64_t src_begin, int64_t src_end,
               int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used) {
    const int64_t n = g->n;
    if (g->src_begin != 0 || g->src_end != n) {
        set_error("vga_global: the graph must hold the rows of all cells (gather the shards first)");
        return VGA_ERR_INVALID;
    }
    int64_t nsrc;
    if (sources) {
        nsrc = src_end;  // list length
        for (int64_t i = 0; i < nsrc; i++)
            if (sources[i] < 0 || sources[i] >= n) {
                set_error("vga_global: source ordinal out of range");
                return VGA_ERR_INVALID;
            }
        src_begin = 0;
    } else {
        if (src_end < 0 || src_end > n) src_end = n;
        if (src_begin < 0) src_begin = 0;
        nsrc = src_end - src_begin;
    }
    if (levels_used) *levels_used = 0;
    if (nsrc <= 0) return VGA_OK;
    ctx->bfs_sources_hint = nsrc;
    struct HintReset {
        vga_ctx *c;
        ~HintReset() { c->bfs_sources_hint = -1; }
    } hint_reset{ctx};
    int words = (int)ctx->opt.bfs_words;
    if (words <= 0) {
        // auto (measured on B200, profiles/README.md): four words (256 sources) per batch with the lane-cooperative kernels
        // (10^6-cell bench subset: level kernels 1792 / 1508 / 1687 ms for 2 / 4 / 8 words; C4 slice 40.7 / 46.3 ms for 4 / 8).
        // One lane per node (bfs_coop = 0): two words, four when the out-rows are long (from ~384 pyramid nodes per row).
        {
            StageTimer pt(ctx, 6, &ctx->timing.prep_ms);
            pt.start();
            VGA_TRY(ensure_bfs_lists(ctx, g, ctx->opt.bfs_mode != 0));
            pt.stop();
        }
        const bool wide = n > 0 && g->f_nnodes >= ctx->opt.bfs_wide_nodes * n;
        words = ctx->opt.bfs_coop ? 4 : (wide ? 4 : 2);
    }
    while (words > 1 && nsrc <= 64 * (words / 2)) words >>= 1;
    switch (words) {
    case 8: return run_global_w<8>(ctx, g, radius, sources, src_begin, nsrc, total_nodes, total_depth, dist, max_levels, levels_used);
    case 4: return run_global_w<4>(ctx, g, radius, sources, src_begin, nsrc, total_nodes, total_depth, dist, max_levels, levels_used);
    case 2: return run_global_w<2>(ctx, g, radius, sources, src_begin, nsrc, total_nodes, total_depth, dist, max_levels, levels_used);
    default: return run_global_w<1>(ctx, g, radius, sources, src_begin, nsrc, total_nodes, total_depth, dist, max_levels, levels_used);
    }
}

}  // namespace vga
