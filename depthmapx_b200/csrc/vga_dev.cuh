// Internal declarations shared by the CUDA translation units of libvga_b200.so.
// Public surface: include/vga_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/vga_b200.h"

// Work counters of the CPU emulation used by the tests (tests/emu/): nothing in the CUDA build.
#ifdef VGA_SIMT_EMULATION
#define VGA_COUNT(name, n)                                      \
    do {                                                        \
        static long long &vga_count_slot__ = simt::counters()[#name]; \
        vga_count_slot__ += (long long)(n);                     \
    } while (0)
#else
#define VGA_COUNT(name, n) ((void)0)
#endif

namespace vga {

void set_error(const std::string &msg);

#define VGA_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            vga::set_error(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + \
                           ":" + std::to_string(__LINE__) + ")");                                 \
            return VGA_ERR_CUDA;                                                                  \
        }                                                                                         \
    } while (0)

#define VGA_TRY(call)               \
    do {                            \
        int rc__ = (call);          \
        if (rc__ != VGA_OK) return rc__; \
    } while (0)

// Stream all allocations of the calling thread are ordered on (set by every C-ABI entry point to
// its context's stream).  Device memory comes from CUDA's stream-ordered pool with an unlimited
// release threshold, so the GB-sized BFS state / edge buffers are recycled between calls instead of
// being mapped and unmapped every time (cudaMalloc/cudaFree of GBs costs 100s of ms).
extern thread_local cudaStream_t g_alloc_stream;

// Cache of large device blocks, one per context (keyed by the context's stream).  A step of the hot path allocates and
// frees the same multi-GB buffers over and over (22 GB of entries, GBs of run / node lists at 10^6 cells); CUDA's
// stream-ordered pool re-maps physical memory whenever its free blocks were split by smaller requests in between, which
// showed up as seconds of jitter per call (bench: 1.3 s -> 4.5 s for the same build).  Blocks of BIG_BLOCK bytes or
// more are therefore parked here on release and handed out again (smallest block that fits and is not more than twice
// the request); all use is ordered on the context's single stream.
struct BlockCache {
    std::mutex m;
    std::vector<std::pair<void *, size_t>> blocks;
    cudaStream_t stream = nullptr;
    bool closed = false;
    static constexpr size_t BIG_BLOCK = (size_t)16 << 20;
    void *take(size_t bytes, size_t *got) {
        std::lock_guard<std::mutex> lock(m);
        int best = -1;
        for (int i = 0; i < (int)blocks.size(); i++)
            if (blocks[i].second >= bytes && blocks[i].second <= 2 * bytes && (best < 0 || blocks[i].second < blocks[best].second)) best = i;
        if (best < 0) return nullptr;
        void *p = blocks[best].first;
        *got = blocks[best].second;
        blocks.erase(blocks.begin() + best);
        return p;
    }
    bool give(void *p, size_t bytes) {
        std::lock_guard<std::mutex> lock(m);
        if (closed) return false;
        blocks.emplace_back(p, bytes);
        // bounded: the oldest parked blocks go back to the pool beyond 48 blocks / 96 GB
        size_t held = 0;
        for (auto &b : blocks) held += b.second;
        while (!blocks.empty() && (blocks.size() > 48 || held > ((size_t)96 << 30))) {
            held -= blocks.front().second;
            cudaFreeAsync(blocks.front().first, stream);
            blocks.erase(blocks.begin());
        }
        return true;
    }
    void drain() {
        std::lock_guard<std::mutex> lock(m);
        for (auto &b : blocks) cudaFreeAsync(b.first, stream);
        blocks.clear();
        closed = true;
    }
};
inline std::mutex g_cache_registry_mutex;
inline std::vector<std::shared_ptr<BlockCache>> g_cache_registry;
inline std::shared_ptr<BlockCache> cache_of_stream(cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_cache_registry_mutex);
    for (auto &c : g_cache_registry)
        if (c->stream == s) return c;
    return nullptr;
}
inline void register_cache(cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_cache_registry_mutex);
    auto c = std::make_shared<BlockCache>();
    c->stream = s;
    g_cache_registry.push_back(c);
}
inline void unregister_cache(cudaStream_t s) {
    std::shared_ptr<BlockCache> c;
    {
        std::lock_guard<std::mutex> lock(g_cache_registry_mutex);
        for (size_t i = 0; i < g_cache_registry.size(); i++)
            if (g_cache_registry[i]->stream == s) {
                c = g_cache_registry[i];
                g_cache_registry.erase(g_cache_registry.begin() + i);
                break;
            }
    }
    if (c) c->drain();
}

// RAII device buffer (cudaMallocAsync / cudaFreeAsync, large blocks through the context's BlockCache); zero-size
// allocations are nullptr.
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    size_t block_bytes = 0;             // size of the underlying block when it came from / goes to the cache
    std::shared_ptr<BlockCache> cache;  // where a large block returns to
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), block_bytes(o.block_bytes), cache(std::move(o.cache)) {
        o.p = nullptr;
        o.n = 0;
        o.block_bytes = 0;
    }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            block_bytes = o.block_bytes;
            cache = std::move(o.cache);
            o.p = nullptr;
            o.n = 0;
            o.block_bytes = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) {
            if (!(cache && block_bytes && cache->give(p, block_bytes))) cudaFreeAsync(p, g_alloc_stream);
        }
        p = nullptr;
        n = 0;
        block_bytes = 0;
        cache.reset();
    }
    int alloc(size_t count) {
        release();
        if (count == 0) return VGA_OK;
        const size_t bytes = count * sizeof(T);
        if (bytes >= BlockCache::BIG_BLOCK) {
            cache = cache_of_stream(g_alloc_stream);
            if (cache) {
                size_t got = 0;
                if (void *q = cache->take(bytes, &got)) {
                    p = (T *)q;
                    n = count;
                    block_bytes = got;
                    return VGA_OK;
                }
                block_bytes = bytes;
            }
        }
        VGA_CUDA(cudaMallocAsync((void **)&p, bytes, g_alloc_stream));
        n = count;
        return VGA_OK;
    }
    int alloc_zero(size_t count, cudaStream_t s) {
        VGA_TRY(alloc(count));
        if (count) VGA_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s));
        return VGA_OK;
    }
};

// Grow-only named device buffers kept by a context across calls (the GB-sized BFS state): even the
// stream-ordered pool occasionally has to map fresh memory when block sizes differ between calls,
// which showed up as 30-40 ms of jitter per vga_global call.
struct Workspace {
    std::vector<std::pair<std::string, DevBuf<unsigned char>>> bufs;
    int get(const char *name, size_t bytes, void **out) {
        for (auto &b : bufs)
            if (b.first == name) {
                if (b.second.n < bytes) VGA_TRY(b.second.alloc(bytes));
                *out = b.second.p;
                return VGA_OK;
            }
        bufs.emplace_back(std::string(name), DevBuf<unsigned char>());
        VGA_TRY(bufs.back().second.alloc(bytes));
        *out = bufs.back().second.p;
        return VGA_OK;
    }
    void clear() { bufs.clear(); }
};

// Typed view of a named workspace buffer: same alloc / alloc_zero interface as DevBuf, but the memory stays with the
// context between calls (the ~25 temporaries of one vga_graph_build were a source of millisecond-scale jitter).
template <typename T> struct WsBuf {
    Workspace *ws;
    const char *name;
    T *p = nullptr;
    size_t n = 0;
    WsBuf(Workspace &w, const char *nm) : ws(&w), name(nm) {}
    int alloc(size_t count) {
        void *q = nullptr;
        VGA_TRY(ws->get(name, (count ? count : 1) * sizeof(T), &q));
        p = (T *)q;
        n = count;
        return VGA_OK;
    }
    int alloc_zero(size_t count, cudaStream_t s) {
        VGA_TRY(alloc(count));
        if (count) VGA_CUDA(cudaMemsetAsync(p, 0, count * sizeof(T), s));
        return VGA_OK;
    }
};

struct Timing {
    double h2d_ms = 0, kernel_ms = 0, d2h_ms = 0, main_kernel_ms = 0, algo_bytes = 0, algo_bytes_csr = 0, prep_ms = 0;
    int64_t batch_words = 0;
    int64_t launches = 0, main_launches = 0;
};

// Tunables (vga_ctx_set_option / environment)
struct Options {
    int64_t bfs_mode = 2;        // 0 top-down (push) only, 1 bottom-up (pull) only, 2 direction-optimising hybrid
    int64_t bfs_words = 0;       // 64-bit words per vertex and batch (a batch = 64*words sources): 1, 2, 4, 8; 0 = auto
    int64_t local_mode = 2;      // 2: auto (4 for dense graphs that fit, 1 while the graph holds entries, 3 for a runs-only graph);
                                 // 4: tensor cores (tcgen05 int8 masked product, local_tc.cu); 1: bit-parallel batches of
                                 // 64 cells over entries; 3: one CTA per cell, bitmaps fed by run-length rows; 0: the same fed
                                 // by entries
    int64_t local_span = 0;      // run-length local kernel: bits of the vertex universe per pass (0 = what fits shared memory);
                                 // tests force several passes on small plans with it
    int64_t bfs_chunk = 0;       // 64-source words in flight; 0 = auto from free memory
    int64_t sieve_mode = 1;      // 1: thread-per-(source,octant) kernel + warp kernel for overflow, 0: warp kernel
    int64_t sieve_thread_cap = -1; // thread kernel capacity: 0 = 8 gaps / 24 blocks per task, 1 = 16 / 48, -1 = auto (1 from
                                 // 200,000 cells: on the 10^6-cell plan 30 % of the tasks exceed 8 / 24 and fall back to the
                                 // warp kernel, 4.8 % exceed 16 / 48: 1342 -> 1123 ms; small plans lose occupancy: 10.9 -> 13.4 ms)
    int64_t sieve_sort_emit = 1; // thread kernel, emit pass: tasks in order of decreasing size (counts of pass 1), so that the 32 tasks
                                 // of a warp are alike; 0 = 32 consecutive sources per warp as in the count pass
    int64_t sieve_gcap = 48;     // shared-memory gap capacity per warp
    int64_t sieve_bcap = 192;    // shared-memory block capacity per warp
    int64_t sieve_big_gcap = 4096;
    int64_t sieve_big_bcap = 32768;
    int64_t build_chunk_entries = (int64_t)1 << 30;
    int64_t build_sort = 1;      // rows ordered by 1: bitmap rank in shared memory (k_rank_sort, own kernel; measured 1.13x-1.4x
                                 // faster builds); 0: cub segmented radix sort (also the fallback for grids of > 1.2M cells)
    int64_t pull_alpha = 0;      // bottom-up step when (frontier out-nodes + 2n) * alpha > (open vertices' in-nodes + n) * beta;
                                 // 0 = 2 when the graph has x- and y-major lists, else 1
    int64_t pull_beta = 1;
    int64_t bfs_order = 2;       // 0: x-major ordinals, 1: 8x8 tiles in Morton order, 2: + wall-respecting floods
    int64_t bfs_coop = 1;        // W >= 4: W/2 lanes share a node, 16 bytes each (k_push_nodes_coop / k_pull_nodes_coop; measured
                                 // on the 10^6-cell bench subset: level kernels 1792 ms at W = 2 -> 1508 ms at W = 4); 0 = one
                                 // lane per node for every W
    int64_t bfs_hybrid = 1;      // every row is walked as the shorter of its x-major and y-major pyramid-node lists: 2 = always,
                                 // 0 = x-major only, 1 = when the call that derives the lists has at least n / 4 sources
                                 // (measured on the 10^6-cell plan: the second set of lists costs ~0.2 us per row once per
                                 // graph and saves ~1 us per source; a rank of an 8-GPU run has n / 8 sources)
    int64_t bfs_push_unroll = 4; // nodes per lane in flight in the top-down step (1, 2 or 4)
    int64_t metric_slots = 0;    // metric / angular VGA: sources (warps) in flight, 0 = resident warps bounded by memory
    int64_t bfs_pull_unroll = 4; // node loads per lane between two early-exit checks of the bottom-up step: 4 (measured on the
                                 // C5 bench subset: level kernels 1483 -> 1439 ms) or 2
    int64_t bfs_delta = 1;       // top-down step: a vertex leaves the nodes that row u-1 covers as well to vertex u-1 for the source
                                 // bits both push (k_push_delta; 93 % of the out-row nodes of an urban plan); 0 = every vertex
                                 // pushes its whole row (k_push_nodes[_coop])
    int64_t bfs_delta_unroll = 1; // k_push_delta: nodes per lane in flight (1, 2 or 4; measured on the C5 bench subset: level kernels 859 / 897 / 932 ms)
    int64_t bfs_delta_weight = 6; // direction rule with k_push_delta: one visited node counts as this many offered in-row nodes
                                 // (measured per level on a C5 slice: ~20 ps per top-down visit, 1.6-5.7 ps per node offered to
                                 // the bottom-up step; any weight from 4 to 12 picks the faster direction at every level there)
    int64_t bfs_wide_nodes = 384; // auto word width with bfs_coop = 0: 4 words from this many pyramid nodes per out-row, else 2
};

}  // namespace vga

struct vga_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    int sm_count = 148;
    size_t smem_optin = 0;
    vga::Options opt;
    vga::Timing timing;
    vga::Workspace ws;
    vga_progress_fn progress = nullptr;
    vga_cancel_fn cancel = nullptr;
    void *user = nullptr;
    int64_t bfs_sources_hint = -1;  // sources of the vga_global call in progress (-1 = unknown): decides whether y-major lists pay
};

struct vga_dgrid {
    vga_ctx *ctx = nullptr;
    int32_t cols = 0, rows = 0;
    double spacing = 0, bl_x = 0, bl_y = 0, maxdist = -1.0;
    int64_t cells = 0, n = 0, nseg = 0;
    vga::DevBuf<uint8_t> cflag;    // [cells] x-major: bit0 FILLED, bit1 has lines
    vga::DevBuf<uint8_t> cflag_t;  // [cells] y-major copy (index y*cols+x) for the y-major octants
    vga::DevBuf<uint32_t> line_off;
    vga::DevBuf<double> lines;
    vga::DevBuf<int32_t> cellord;  // [cells] ordinal, or -(1 + filled cells before) for unfilled cells
    vga::DevBuf<int32_t> cellref;  // [n] packed PixelRef per ordinal
    std::vector<int32_t> h_cellref;
    std::vector<int32_t> h_ghostref;  // packed PixelRef of every unfilled cell, x-major
};

struct vga_graph {
    vga_ctx *ctx = nullptr;
    int64_t n = 0;        // filled cells
    int64_t ghosts = 0;   // ghost vertices
    int64_t src_begin = 0, src_end = 0;
    int64_t entries = 0;  // adjacency entries of rows [src_begin, src_end)
    // rows sorted by column; packed entry = col<<6 | accepted<<5 | bin
    vga::DevBuf<uint64_t> rowptr;  // [rows+1], local (starts at 0)
    vga::DevBuf<uint32_t> adj;     // [entries]
    bool has_bins = false;
    // node stats (only for built graphs)
    bool has_stats = false;
    vga::DevBuf<int32_t> connectivity;
    vga::DevBuf<double> sum_d, sum_d2;
    vga::DevBuf<float> far_dist;    // [rows*32]
    vga::DevBuf<int32_t> bin_count; // [rows*32]
    vga::DevBuf<uint8_t> gridconn;
    std::vector<int32_t> h_refs;    // N + G packed PixelRefs (host)
    vga::DevBuf<uint8_t> noexpand;  // [n] != 0: counted but not expanded (context-filled, not even); empty = none
    // ---- BFS structures, built lazily from the sorted rows (bfs.cu) or adopted (vga_graph_from_device_runs) ----------
    // Rows of a grid visibility graph are unions of a few runs of consecutive ordinals.  f_runs: the out-row of u is the
    // union of f_runs[f_runptr[u] .. f_runptr[u+1]) = (first ordinal, length), ghost runs (first >= n) last; t_runs: the same
    // for the in-rows (transposed by bfs.cu in O(runs), no entry-sized transpose exists).  f_nodes / t_nodes: per row the
    // ids of the OR-pyramid nodes that tile its runs (id < n: the vertex's own word, id >= n: inner node id - n), row
    // offsets f_nodeptr / t_nodeptr.  The BFS reads only the node lists.
    vga::DevBuf<uint64_t> f_runptr, t_runptr;    // [n+1]
    vga::DevBuf<uint2> f_runs, t_runs;
    int64_t f_nruns = 0, t_nruns = 0, f_nnodes = 0, t_nnodes = 0;
    bool has_fwd_runs = false, has_runs = false;
    bool runs_only = false;                      // adopted from run-length rows: no entry rows, bins or statistics
    bool has_shard_runs = false;                 // f_runs hold the rows [src_begin, src_end) of a shard only
    vga::DevBuf<uint64_t> f_nodeptr, t_nodeptr;  // [n+1]
    vga::DevBuf<uint32_t> f_nodes, t_nodes;
    // out-rows: the first f_split[u] nodes of row u are "new", the others cover only cells that row u-1 holds as well (the
    // top-down step leaves those to vertex u-1 when it pushes a superset of u's source bits, bfs.cu k_push_delta)
    vga::DevBuf<uint32_t> f_split;               // [n]
    bool has_f_nodes = false, has_t_nodes = false;
    // y-major alternative (bfs.cu): perm_y[v] = rank of cell v in y-major order, perm_x = its inverse; f_isy / t_isy[v] = 1
    // when the list of row v is the y-major one (inner node ids then refer to the pyramid over the y-major order; leaf ids
    // are x-major ordinals in both kinds of list); fy_runs = y-major out-rows in x-major row order (kept until the in-rows
    // are derived)
    vga::DevBuf<uint32_t> perm_y, perm_x;
    vga::DevBuf<uint8_t> f_isy, t_isy;
    vga::DevBuf<uint64_t> fy_runptr;
    vga::DevBuf<uint2> fy_runs;
    int64_t fy_nruns = 0;
    bool has_y = false;
    vga::DevBuf<uint32_t> deg;                   // [n] entries per row when the graph holds runs only (statistics), else empty
    // bit matrices of the tensor-core local kernel (local_tc.cu), built on first use: A[v][w] and its transpose A^T[w][u]
    vga::DevBuf<uint32_t> tc_bits_a, tc_bits_t;
    // spatially coherent order of all n sources (cached: the host-side clustering costs milliseconds per call)
    std::vector<int32_t> h_order;
};

namespace vga {
// makegraph.cu
int build_graph(vga_ctx *ctx, const vga_dgrid *g, int64_t src_begin, int64_t src_end, vga_graph **out);
// bfs.cu
int ensure_fwd_runs(vga_ctx *ctx, vga_graph *g);  // run-length out-rows from the sorted entries
int shard_runs(vga_ctx *ctx, vga_graph *g);       // the same for the rows a shard holds
int row_degrees(vga_ctx *ctx, vga_graph *g);      // entries per held row (g->deg)
int ensure_bfs_lists(vga_ctx *ctx, vga_graph *g, bool transposed);  // node lists (+ transposed runs / node lists)
// cabi.cu: per-segment ascending sort of 32-bit keys (cub::DeviceSegmentedSort in slices of < 2^31 entries)
int sort_segments_u32(vga_ctx *ctx, uint32_t *keys_in, uint32_t *keys_out, int64_t entries, int64_t rows, const uint64_t *rowptr);
// spatially coherent order of the sources [src_begin, src_end) for 64-source batches
int batch_source_order(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, std::vector<int32_t> &order);
int batch_source_order_list(vga_ctx *ctx, vga_graph *g, const int64_t *sources, int64_t nsrc, std::vector<int32_t> &order);
// sources: explicit list of ordinals (nullptr = the range [src_begin, src_end)); outputs in list order
int run_global(vga_ctx *ctx, vga_graph *g, int radius, const int64_t *sources, int64_t src_begin, int64_t src_end,
               int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used);
// metric.cu: metric (angular = 0: 4 output columns) / angular (3 columns) VGA for an explicit list of sources
int run_metric_angular(vga_ctx *ctx, vga_graph *g, int angular, const uint8_t *expand, const int32_t *partner, double spacing, double radius,
                       const int64_t *sources, int64_t nsrc, float *const *out, int nout, int64_t *angle_unsafe);
// stepdepth.cu
int run_step_depth(vga_ctx *ctx, vga_graph *g, const int64_t *sources, int64_t nsrc, int32_t *depth_out);
// local_tc.cu: tcgen05 kernel for dense graphs (cluster, k, total; control comes from local.cu)
bool local_tc_applicable(const vga_graph *g);
int run_local_tc(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k, int32_t *total);
// local.cu
int run_local(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k,
              int32_t *total, float *control);

// 32 x 32 bit-matrix transpose across a warp: lane l passes row l and receives column l (bit i = bit l of lane i's
// row), five butterfly stages of one shuffle each.  popc of the result = how many lanes had bit l set.
__device__ __forceinline__ unsigned warp_transpose32(unsigned x, int lane) {
    const unsigned masks[5] = {0x0000ffffu, 0x00ff00ffu, 0x0f0f0f0fu, 0x33333333u, 0x55555555u};
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int j = 16 >> s;
        const unsigned m = masks[s];
        const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
        if (lane & j) {
            x ^= ((y >> j) ^ x) & m;
        } else {
            x ^= (((x >> j) ^ y) & m) << j;
        }
    }
    return x;
}
constexpr int DENSE_COLUMNS = 8;  // from this many non-empty bit columns on, transpose + popc beats one ballot per column

// helpers for timing: record an event pair around a region on ctx->stream
struct StageTimer {
    vga_ctx *ctx;
    cudaEvent_t a, b;
    double *acc;
    bool on = false;
    StageTimer(vga_ctx *c, int slot, double *accum) : ctx(c), a(c->ev[slot]), b(c->ev[slot + 1]), acc(accum) {}
    void start() {
        cudaEventRecord(a, ctx->stream);
        on = true;
    }
    // stop() synchronises the stream
    void stop() {
        if (!on) return;
        cudaEventRecord(b, ctx->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        *acc += ms;
        on = false;
    }
};
}  // namespace vga
