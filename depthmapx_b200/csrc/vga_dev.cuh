// Internal declarations shared by the CUDA translation units of libvga_b200.so.
// Public surface: include/vga_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

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

// RAII device buffer (cudaMallocAsync / cudaFreeAsync); zero-size allocations are nullptr.
template <typename T> struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) {
            release();
            p = o.p;
            n = o.n;
            o.p = nullptr;
            o.n = 0;
        }
        return *this;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFreeAsync(p, g_alloc_stream);
        p = nullptr;
        n = 0;
    }
    int alloc(size_t count) {
        release();
        if (count == 0) return VGA_OK;
        VGA_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), g_alloc_stream));
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

struct Timing {
    double h2d_ms = 0, kernel_ms = 0, d2h_ms = 0, main_kernel_ms = 0, algo_bytes = 0, algo_bytes_runs = 0;
    int64_t launches = 0, main_launches = 0;
};

// Tunables (vga_ctx_set_option / environment)
struct Options {
    int64_t bfs_mode = 2;        // 0 push only, 1 pull only, 2 direction-optimising hybrid
    int64_t bfs_words = 0;       // 64-bit words per vertex and batch (a batch = 64*words sources): 1, 2, 4; 0 = auto
    int64_t local_mode = 2;      // 0: one CTA per cell with bitmaps, 1: bit-parallel batches of 64 cells, 2: auto,
                                 // 3: batches with bit-sliced counters (EXPERIMENTAL, opt-in, not yet run on a GPU)
    int64_t bfs_chunk = 0;       // 64-source words in flight; 0 = auto from free memory
    int64_t sieve_mode = 1;      // 1: thread-per-(source,octant) kernel + warp kernel for overflow, 0: warp kernel
    int64_t sieve_gcap = 48;     // shared-memory gap capacity per warp
    int64_t sieve_bcap = 192;    // shared-memory block capacity per warp
    int64_t sieve_big_gcap = 4096;
    int64_t sieve_big_bcap = 32768;
    int64_t build_chunk_entries = (int64_t)1 << 30;
    int64_t build_sort = 0;      // rows sorted by 0: cub segmented radix sort; 1: bitmap rank in shared memory (k_rank_sort;
                                 // EXPERIMENTAL, opt-in, validated under SIMT emulation)
    int64_t pull_alpha = 1;      // pull when frontier edges * alpha > candidate in-edges * beta (per batch)
    int64_t pull_beta = 1;
    int64_t bfs_order = 2;       // 0: x-major ordinals, 1: 8x8 tiles in Morton order, 2: + wall-respecting floods
    int64_t bfs_group = 16;      // batches per coarse lower-bound group
    int64_t bfs_coarse = 1;      // 1: coarse pass (64 group lower bounds per bit-parallel batch) prunes the pull step
    int64_t bfs_push_unroll = 1; // adjacency entries per lane and round in the top-down step: 1 (default) or 4 (EXPERIMENTAL,
                                 // opt-in, not yet run on a GPU: more loads in flight per lane)
    int64_t bfs_push = 0;        // top-down step: 0 = adjacency entries (k_push); 1 = additionally range-OR updates over the
                                 // runs of the out-rows through a pyramid of `next` (k_push_pyr + k_pyr_down) whenever the
                                 // per-batch cost model prefers them (EXPERIMENTAL, opt-in; validated under SIMT emulation)
    int64_t bfs_pyr_cost = 100;  // percent: cost of one pyramid-push node relative to one adjacency entry in the per-batch choice
                                 // between entry push, pyramid push and pull (to be tuned on the B200)
    int64_t bfs_pyr_nodes = 0;   // with bfs_push / bfs_pull = 1: rows as lists of pyramid NODE ids instead of runs, walked by
                                 // copies of the entry kernels' inner loops (k_push_nodes / k_pull_nodes; EXPERIMENTAL, opt-in)
    int64_t bfs_pull = 0;        // bottom-up step: 0 scans in-row entries (k_pull); 1 = range-OR queries over an
                                 // OR-pyramid of the frontier with run-length in-rows (k_pull_pyr; EXPERIMENTAL, opt-in,
                                 // index logic unit-tested on CPU, kernels not yet run on a GPU)
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
    // derived analysis structures, built lazily
    vga::DevBuf<uint64_t> t_rowptr; // transpose (in-edges), [n+1]
    vga::DevBuf<uint32_t> t_col;    // [entries to filled targets]
    bool has_transpose = false;
    int64_t t_entries = 0;
    // run-length in-rows for the pyramid pull (bfs_pull = 1), built lazily: the sorted in-row of v is the union of
    // the runs t_runs[t_runptr[v] .. t_runptr[v+1]) = (first ordinal, length); t_costptr = prefix sums of the
    // pyramid loads a full scan of each in-row costs
    vga::DevBuf<uint64_t> t_runptr;  // [n+1]
    vga::DevBuf<uint2> t_runs;
    vga::DevBuf<uint64_t> t_costptr; // [n+1]
    bool has_runs = false;
    int64_t t_nruns = 0;
    // run-length out-rows for the pyramid push (bfs_push = 1), ghost columns excluded
    vga::DevBuf<uint64_t> f_runptr;
    vga::DevBuf<uint2> f_runs;
    vga::DevBuf<uint64_t> f_costptr;
    bool has_fwd_runs = false;
    int64_t f_nruns = 0;
    // node-id lists (bfs_pyr_nodes = 1): the pyramid nodes tiling each row's runs; row offsets = f_costptr / t_costptr
    vga::DevBuf<uint32_t> f_nodes, t_nodes;
    bool has_f_nodes = false, has_t_nodes = false;
};

namespace vga {
// makegraph.cu
int build_graph(vga_ctx *ctx, const vga_dgrid *g, int64_t src_begin, int64_t src_end, vga_graph **out);
// bfs.cu
int ensure_transpose(vga_ctx *ctx, vga_graph *g);
int ensure_runs(vga_ctx *ctx, vga_graph *g);
int ensure_fwd_runs(vga_ctx *ctx, vga_graph *g);
int ensure_node_lists(vga_ctx *ctx, vga_graph *g, bool fwd, bool transposed);
// cabi.cu: per-segment ascending sort of 32-bit keys (cub::DeviceSegmentedSort in slices of < 2^31 entries)
int sort_segments_u32(vga_ctx *ctx, uint32_t *keys_in, uint32_t *keys_out, int64_t entries, int64_t rows, const uint64_t *rowptr);
// spatially coherent order of the sources [src_begin, src_end) for 64-source batches
int batch_source_order(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, std::vector<int32_t> &order);
int run_global(vga_ctx *ctx, vga_graph *g, int radius, int64_t src_begin, int64_t src_end, int32_t *total_nodes,
               int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used);
// stepdepth.cu
int run_step_depth(vga_ctx *ctx, vga_graph *g, const int64_t *sources, int64_t nsrc, int32_t *depth_out);
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
