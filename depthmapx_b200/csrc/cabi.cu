// C ABI of libvga_b200.so (include/vga_b200.h): context, uploads, exports, and the host-side
// formula stages.  All compute goes to the CUDA kernels in makegraph.cu / bfs.cu / local.cu;
// there is no CPU implementation behind any compute entry point.
#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "vga_dev.cuh"

namespace vga {

static thread_local std::string g_last_error;
thread_local cudaStream_t g_alloc_stream = nullptr;

void set_error(const std::string &msg) { g_last_error = msg; }

static int64_t env_i64(const char *name, int64_t dflt) {
    const char *v = std::getenv(name);
    if (!v || !*v) return dflt;
    return std::strtoll(v, nullptr, 10);
}

static int set_option(Options &o, const std::string &key, int64_t value) {
    if (key == "bfs_mode") o.bfs_mode = value;
    else if (key == "bfs_chunk") o.bfs_chunk = value;
    else if (key == "bfs_words") o.bfs_words = value;
    else if (key == "bfs_wide_nodes") o.bfs_wide_nodes = value;
    else if (key == "bfs_coop") o.bfs_coop = value;
    else if (key == "bfs_pull_unroll") o.bfs_pull_unroll = value;
    else if (key == "bfs_hybrid") o.bfs_hybrid = value;
    else if (key == "metric_slots") o.metric_slots = value;
    else if (key == "bfs_push_unroll") o.bfs_push_unroll = value;
    else if (key == "bfs_delta") o.bfs_delta = value;
    else if (key == "bfs_delta_unroll") o.bfs_delta_unroll = value;
    else if (key == "bfs_delta_weight") o.bfs_delta_weight = value;
    else if (key == "local_mode") o.local_mode = value;
    else if (key == "local_span") o.local_span = value;
    else if (key == "sieve_mode") o.sieve_mode = value;
    else if (key == "sieve_thread_cap") o.sieve_thread_cap = value;
    else if (key == "sieve_sort_emit") o.sieve_sort_emit = value;
    else if (key == "sieve_gcap") o.sieve_gcap = value;
    else if (key == "sieve_bcap") o.sieve_bcap = value;
    else if (key == "sieve_big_gcap") o.sieve_big_gcap = value;
    else if (key == "sieve_big_bcap") o.sieve_big_bcap = value;
    else if (key == "build_chunk_entries") o.build_chunk_entries = value;
    else if (key == "build_sort") o.build_sort = value;
    else if (key == "pull_alpha") o.pull_alpha = value;
    else if (key == "pull_beta") o.pull_beta = value;
    else if (key == "bfs_order") o.bfs_order = value;
    else return VGA_ERR_INVALID;
    return VGA_OK;
}

namespace {

__global__ void k_pack_keys(const uint32_t *col, const uint8_t *bin, uint64_t n, uint32_t *keys) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    keys[e] = (col[e] << 6) | 32u | (bin ? (uint32_t)(bin[e] & 31) : 0u);
}

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

int sort_rows(vga_ctx *ctx, uint32_t *keys_in, uint32_t *keys_out, int64_t entries, int64_t rows, const uint64_t *rowptr) {
    // segmented sort in slices of < 2^31 entries
    cudaStream_t st = ctx->stream;
    std::vector<uint64_t> h((size_t)rows + 1);
    VGA_CUDA(cudaMemcpyAsync(h.data(), rowptr, sizeof(uint64_t) * (rows + 1), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    DevBuf<unsigned char> tmp;
    DevBuf<uint64_t> seg;
    VGA_TRY(seg.alloc((size_t)rows + 1));
    const uint64_t cap = (uint64_t)1 << 30;
    int64_t i = 0;
    while (i < rows) {
        int64_t j = i + 1;
        while (j < rows && h[j + 1] - h[i] <= cap) j++;
        uint64_t base = h[i], cnt = h[j] - h[i];
        if (cnt >= ((uint64_t)1 << 31)) {
            set_error("row with more than 2^31 entries");
            return VGA_ERR_CAPACITY;
        }
        if (cnt > 0) {
            std::vector<uint64_t> local((size_t)(j - i) + 1);
            for (int64_t r = i; r <= j; r++) local[(size_t)(r - i)] = h[r] - base;
            VGA_CUDA(cudaMemcpyAsync(seg.p, local.data(), sizeof(uint64_t) * local.size(), cudaMemcpyHostToDevice, st));
            size_t tb = 0;
            cub::DeviceSegmentedSort::SortKeys(nullptr, tb, keys_in + base, keys_out + base, (int)cnt, (int)(j - i), seg.p,
                                               seg.p + 1, st);
            if (tb + 16 > tmp.n) VGA_TRY(tmp.alloc(tb + 16));
            VGA_CUDA(cub::DeviceSegmentedSort::SortKeys(tmp.p, tb, keys_in + base, keys_out + base, (int)cnt, (int)(j - i),
                                                        seg.p, seg.p + 1, st));
            VGA_CUDA(cudaStreamSynchronize(st));
            ctx->timing.launches += 3;
        }
        i = j;
    }
    return VGA_OK;
}

// ---- formula stage, FP64 with glibc libm, cast to float exactly where setValue does ------------
// project-local log2 (genlib/pafmath.h:44,61): ln(a) * M_1_LN2
const double kInvLn2 = 1.4426950408889634073599246810019;
inline double log2_paf(double a) { return std::log(a) * kInvLn2; }
inline double d_value(double k) { return 2.0 * (k * (log2_paf((k + 2.0) / 3.0) - 1.0) + 1.0) / ((k - 1.0) * (k - 2.0)); }
inline double p_value(double k) { return 2.0 * (k - log2_paf(k) - 1.0) / ((k - 1.0) * (k - 2.0)); }
inline double tekl_integ(double nodes, double depth) { return std::log(0.5 * (nodes - 2.0)) / std::log(depth - nodes + 1.0); }

}  // namespace

int sort_segments_u32(vga_ctx *ctx, uint32_t *keys_in, uint32_t *keys_out, int64_t entries, int64_t rows, const uint64_t *rowptr) {
    return sort_rows(ctx, keys_in, keys_out, entries, rows, rowptr);
}
}  // namespace vga

using namespace vga;

extern "C" {

const char *vga_last_error(void) { return g_last_error.c_str(); }
const char *vga_version(void) { return "vga_b200 0.1 (sm_100a)"; }

int vga_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int vga_ctx_create(int device, vga_ctx **out) {
    if (!out) return VGA_ERR_INVALID;
    *out = nullptr;
    int n = vga_device_count();
    if (n <= 0) {
        set_error("no CUDA device available: libvga_b200 has no CPU path");
        return VGA_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device index out of range");
        return VGA_ERR_INVALID;
    }
    VGA_CUDA(cudaSetDevice(device));
    std::unique_ptr<vga_ctx> c(new vga_ctx());
    c->device = device;
    VGA_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto &e : c->ev) VGA_CUDA(cudaEventCreate(&e));
    register_cache(c->stream);
    cudaDeviceProp p;
    VGA_CUDA(cudaGetDeviceProperties(&p, device));
    c->sm_count = p.multiProcessorCount;
    c->smem_optin = p.sharedMemPerBlockOptin;
    {
        // keep freed blocks in the stream-ordered pool (see g_alloc_stream)
        cudaMemPool_t pool;
        VGA_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t keep = UINT64_MAX;
        VGA_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    Options &o = c->opt;
    o.bfs_mode = env_i64("VGA_BFS_MODE", o.bfs_mode);
    o.bfs_chunk = env_i64("VGA_BFS_CHUNK", o.bfs_chunk);
    o.sieve_gcap = env_i64("VGA_SIEVE_GCAP", o.sieve_gcap);
    o.sieve_bcap = env_i64("VGA_SIEVE_BCAP", o.sieve_bcap);
    o.pull_alpha = env_i64("VGA_PULL_ALPHA", o.pull_alpha);
    o.bfs_words = env_i64("VGA_BFS_WORDS", o.bfs_words);
    *out = c.release();
    return VGA_OK;
}

void vga_ctx_destroy(vga_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    g_alloc_stream = ctx->stream;
    ctx->ws.clear();
    unregister_cache(ctx->stream);  // parked blocks go back to the pool; buffers still alive free themselves directly
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    }
    if (g_alloc_stream == ctx->stream) g_alloc_stream = nullptr;
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int vga_ctx_set_callbacks(vga_ctx *ctx, vga_progress_fn progress, vga_cancel_fn cancel, void *user) {
    if (!ctx) return VGA_ERR_INVALID;
    ctx->progress = progress;
    ctx->cancel = cancel;
    ctx->user = user;
    return VGA_OK;
}

int vga_ctx_set_option(vga_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return VGA_ERR_INVALID;
    int rc = set_option(ctx->opt, key, value);
    if (rc != VGA_OK) set_error(std::string("unknown option ") + key);
    return rc;
}

int vga_ctx_timing(const vga_ctx *ctx, vga_timing *out) {
    if (!ctx || !out) return VGA_ERR_INVALID;
    out->h2d_ms = ctx->timing.h2d_ms;
    out->kernel_ms = ctx->timing.kernel_ms;
    out->d2h_ms = ctx->timing.d2h_ms;
    out->main_kernel_ms = ctx->timing.main_kernel_ms;
    out->launches = ctx->timing.launches;
    out->main_launches = ctx->timing.main_launches;
    out->algo_bytes = ctx->timing.algo_bytes;
    out->algo_bytes_csr = ctx->timing.algo_bytes_csr;
    out->prep_ms = ctx->timing.prep_ms;
    out->batch_words = ctx->timing.batch_words;
    return VGA_OK;
}

int vga_ctx_sync(vga_ctx *ctx) {
    if (!ctx) return VGA_ERR_INVALID;
    VGA_CUDA(cudaStreamSynchronize(ctx->stream));
    VGA_CUDA(cudaGetLastError());
    return VGA_OK;
}

// ------------------------------------------------------------------------------------ grid upload

int vga_grid_upload(vga_ctx *ctx, const vga_grid *grid, vga_dgrid **out) {
    if (!ctx || !grid || !out) return VGA_ERR_INVALID;
    *out = nullptr;
    if (grid->cols <= 0 || grid->rows <= 0 || grid->cols > 32767 || grid->rows > 32767 || !grid->state || !grid->line_off) {
        set_error("vga_grid: bad dimensions or null arrays");
        return VGA_ERR_INVALID;
    }
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    const int64_t cells = (int64_t)grid->cols * grid->rows;
    std::unique_ptr<vga_dgrid> d(new vga_dgrid());
    d->ctx = ctx;
    d->cols = grid->cols;
    d->rows = grid->rows;
    d->spacing = grid->spacing;
    d->bl_x = grid->bl_x;
    d->bl_y = grid->bl_y;
    d->maxdist = grid->maxdist;
    d->cells = cells;
    d->nseg = grid->line_off[cells];
    if (d->nseg > 0 && !grid->lines) {
        set_error("vga_grid: lines is null");
        return VGA_ERR_INVALID;
    }
    // host-side flattening: flags (+ transposed copy), ordinals, refs
    std::vector<uint8_t> cf((size_t)cells), cft((size_t)cells);
    std::vector<int32_t> ord((size_t)cells);
    int64_t n = 0;
    for (int64_t x = 0; x < grid->cols; x++)
        for (int64_t y = 0; y < grid->rows; y++) {
            int64_t c = x * grid->rows + y;
            uint16_t s = grid->state[c];
            bool filled = (s & 0x0002) != 0;
            // CONTEXTFILLED (0x0008) and MERGED (0x0040) do not matter to construction: sparkGraph2 looks at FILLED only;
            // the analyses get them through vga_graph_set_noexpand / the contracted adjacency of the host layer
            uint8_t f = (filled ? 1 : 0) | ((grid->line_off[c + 1] > grid->line_off[c]) ? 2 : 0);
            cf[(size_t)c] = f;
            cft[(size_t)(y * grid->cols + x)] = f;
            ord[(size_t)c] = filled ? (int32_t)n : -(int32_t)(1 + n);
            if (filled) {
                d->h_cellref.push_back((int32_t)(((uint32_t)x << 16) | (uint32_t)y));
                n++;
            } else {
                d->h_ghostref.push_back((int32_t)(((uint32_t)x << 16) | (uint32_t)y));
            }
        }
    d->n = n;
    cudaStream_t st = ctx->stream;
    StageTimer ht(ctx, 6, &ctx->timing.h2d_ms);
    VGA_TRY(d->cflag.alloc((size_t)cells));
    VGA_TRY(d->cflag_t.alloc((size_t)cells));
    VGA_TRY(d->line_off.alloc((size_t)cells + 1));
    VGA_TRY(d->lines.alloc((size_t)std::max<int64_t>(d->nseg, 1) * 5));
    VGA_TRY(d->cellord.alloc((size_t)cells));
    VGA_TRY(d->cellref.alloc((size_t)std::max<int64_t>(n, 1)));
    ht.start();
    VGA_CUDA(cudaMemcpyAsync(d->cflag.p, cf.data(), (size_t)cells, cudaMemcpyHostToDevice, st));
    VGA_CUDA(cudaMemcpyAsync(d->cflag_t.p, cft.data(), (size_t)cells, cudaMemcpyHostToDevice, st));
    VGA_CUDA(cudaMemcpyAsync(d->line_off.p, grid->line_off, sizeof(uint32_t) * (cells + 1), cudaMemcpyHostToDevice, st));
    if (d->nseg > 0)
        VGA_CUDA(cudaMemcpyAsync(d->lines.p, grid->lines, sizeof(double) * 5 * d->nseg, cudaMemcpyHostToDevice, st));
    VGA_CUDA(cudaMemcpyAsync(d->cellord.p, ord.data(), sizeof(int32_t) * cells, cudaMemcpyHostToDevice, st));
    if (n > 0)
        VGA_CUDA(cudaMemcpyAsync(d->cellref.p, d->h_cellref.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    ht.stop();
    *out = d.release();
    return VGA_OK;
}

void vga_dgrid_free(vga_dgrid *g) {
    if (g && g->ctx) {
        cudaSetDevice(g->ctx->device);
        g_alloc_stream = g->ctx->stream;
    }
    delete g;
}

int vga_graph_build_resident(vga_ctx *ctx, const vga_dgrid *grid, int64_t src_begin, int64_t src_end, vga_graph **out) {
    if (!ctx || !grid || !out) return VGA_ERR_INVALID;
    *out = nullptr;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    double h2d = ctx->timing.h2d_ms;
    ctx->timing = Timing();
    ctx->timing.h2d_ms = h2d;
    return build_graph(ctx, grid, src_begin, src_end, out);
}

int vga_graph_build(vga_ctx *ctx, const vga_grid *grid, int64_t src_begin, int64_t src_end, vga_graph **out) {
    if (!out) return VGA_ERR_INVALID;
    *out = nullptr;
    vga_dgrid *d = nullptr;
    VGA_TRY(vga_grid_upload(ctx, grid, &d));
    int rc = vga_graph_build_resident(ctx, d, src_begin, src_end, out);
    vga_dgrid_free(d);
    return rc;
}

int vga_graph_from_csr(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, const uint64_t *rowptr, const uint32_t *col,
                       const uint8_t *bin, vga_graph **out) {
    if (!ctx || !out || !rowptr || n_cells < 0 || n_ghosts < 0) return VGA_ERR_INVALID;
    *out = nullptr;
    if (n_cells + n_ghosts >= ((int64_t)1 << 26)) {
        set_error("more than 2^26 vertices is not supported");
        return VGA_ERR_UNSUPPORTED;
    }
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    cudaStream_t st = ctx->stream;
    const int64_t entries = (int64_t)rowptr[n_cells];
    if (entries > 0 && !col) return VGA_ERR_INVALID;
    if (rowptr[0] != 0) {
        set_error("vga_graph_from_csr: rowptr must start at 0");
        return VGA_ERR_INVALID;
    }
    for (int64_t v = 0; v < n_cells; v++)
        if (rowptr[v] > rowptr[v + 1]) {
            set_error("vga_graph_from_csr: rowptr is not non-decreasing");
            return VGA_ERR_INVALID;
        }
    for (int64_t e = 0; e < entries; e++)
        if ((int64_t)col[e] >= n_cells + n_ghosts) {
            set_error("vga_graph_from_csr: column index out of range");
            return VGA_ERR_INVALID;
        }
    std::unique_ptr<vga_graph> g(new vga_graph());
    g->ctx = ctx;
    g->n = n_cells;
    g->ghosts = n_ghosts;
    g->src_begin = 0;
    g->src_end = n_cells;
    g->entries = entries;
    g->has_bins = bin != nullptr;
    DevBuf<uint32_t> d_col, keys;
    DevBuf<uint8_t> d_bin;
    VGA_TRY(g->rowptr.alloc((size_t)n_cells + 1));
    VGA_TRY(g->adj.alloc((size_t)entries + 1));
    VGA_TRY(d_col.alloc((size_t)entries + 1));
    VGA_TRY(keys.alloc((size_t)entries + 1));
    if (bin) VGA_TRY(d_bin.alloc((size_t)entries + 1));
    StageTimer ht(ctx, 6, &ctx->timing.h2d_ms);
    ht.start();
    VGA_CUDA(cudaMemcpyAsync(g->rowptr.p, rowptr, sizeof(uint64_t) * (n_cells + 1), cudaMemcpyHostToDevice, st));
    if (entries > 0) {
        VGA_CUDA(cudaMemcpyAsync(d_col.p, col, sizeof(uint32_t) * entries, cudaMemcpyHostToDevice, st));
        if (bin) VGA_CUDA(cudaMemcpyAsync(d_bin.p, bin, (size_t)entries, cudaMemcpyHostToDevice, st));
    }
    ht.stop();
    StageTimer kt(ctx, 0, &ctx->timing.kernel_ms);
    kt.start();
    if (entries > 0) {
        k_pack_keys<<<blocks_for(entries, 256), 256, 0, st>>>(d_col.p, bin ? d_bin.p : nullptr, (uint64_t)entries, keys.p);
        ctx->timing.launches++;
        VGA_TRY(sort_rows(ctx, keys.p, g->adj.p, entries, n_cells, g->rowptr.p));
    }
    kt.stop();
    VGA_CUDA(cudaGetLastError());
    *out = g.release();
    return VGA_OK;
}

int vga_graph_from_device_rows(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, const uint64_t *d_rowptr,
                               const uint32_t *d_adj, int64_t n_entries, vga_graph **out) {
    if (!ctx || !out || !d_rowptr || n_cells < 0 || n_ghosts < 0 || n_entries < 0 || (n_entries > 0 && !d_adj)) return VGA_ERR_INVALID;
    *out = nullptr;
    if (n_cells + n_ghosts >= ((int64_t)1 << 26)) {
        set_error("more than 2^26 vertices is not supported");
        return VGA_ERR_UNSUPPORTED;
    }
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    cudaStream_t st = ctx->stream;
    std::unique_ptr<vga_graph> g(new vga_graph());
    g->ctx = ctx;
    g->n = n_cells;
    g->ghosts = n_ghosts;
    g->src_begin = 0;
    g->src_end = n_cells;
    g->entries = n_entries;
    g->has_bins = true;
    VGA_TRY(g->rowptr.alloc((size_t)n_cells + 1));
    VGA_TRY(g->adj.alloc((size_t)n_entries + 1));
    VGA_CUDA(cudaMemcpyAsync(g->rowptr.p, d_rowptr, sizeof(uint64_t) * (n_cells + 1), cudaMemcpyDeviceToDevice, st));
    if (n_entries > 0)
        VGA_CUDA(cudaMemcpyAsync(g->adj.p, d_adj, sizeof(uint32_t) * n_entries, cudaMemcpyDeviceToDevice, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    *out = g.release();
    return VGA_OK;
}

void vga_graph_free(vga_graph *g) {
    if (!g) return;
    if (g->ctx) {
        cudaSetDevice(g->ctx->device);
        g_alloc_stream = g->ctx->stream;
    }
    delete g;
}

int64_t vga_graph_num_cells(const vga_graph *g) { return g ? g->n : -1; }
int64_t vga_graph_num_ghosts(const vga_graph *g) { return g ? g->ghosts : -1; }
int64_t vga_graph_num_edges(const vga_graph *g) { return g ? g->entries : -1; }
int64_t vga_graph_src_begin(const vga_graph *g) { return g ? g->src_begin : -1; }
int64_t vga_graph_src_end(const vga_graph *g) { return g ? g->src_end : -1; }

int vga_graph_csr(const vga_graph *g, uint64_t *rowptr, uint32_t *col, uint8_t *bin, uint8_t *accepted) {
    if (!g) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(g->ctx->device));
    g_alloc_stream = g->ctx->stream;
    const int64_t rows = g->src_end - g->src_begin;
    if (rowptr) VGA_CUDA(cudaMemcpy(rowptr, g->rowptr.p, sizeof(uint64_t) * (rows + 1), cudaMemcpyDeviceToHost));
    if ((col || bin || accepted) && g->entries > 0) {
        // unpacked in slices through a bounded staging buffer (a C5 graph has 5.5e9 entries = 22 GB)
        const int64_t slice = (int64_t)1 << 26;
        std::vector<uint32_t> h((size_t)std::min<int64_t>(slice, g->entries));
        for (int64_t e0 = 0; e0 < g->entries; e0 += slice) {
            const int64_t cnt = std::min<int64_t>(slice, g->entries - e0);
            VGA_CUDA(cudaMemcpy(h.data(), g->adj.p + e0, sizeof(uint32_t) * cnt, cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < cnt; i++) {
                const uint32_t v = h[(size_t)i];
                if (col) col[e0 + i] = v >> 6;
                if (bin) bin[e0 + i] = (uint8_t)(v & 31);
                if (accepted) accepted[e0 + i] = (uint8_t)((v >> 5) & 1);
            }
        }
    }
    return VGA_OK;
}

int vga_graph_cell_refs(const vga_graph *g, int32_t *ref) {
    if (!g || !ref) return VGA_ERR_INVALID;
    if ((int64_t)g->h_refs.size() != g->n + g->ghosts) {
        set_error("vga_graph_cell_refs: this graph was adopted from a CSR and has no cell coordinates");
        return VGA_ERR_INVALID;
    }
    std::memcpy(ref, g->h_refs.data(), sizeof(int32_t) * g->h_refs.size());
    return VGA_OK;
}

int vga_graph_set_cell_refs(vga_graph *g, const int32_t *ref, int64_t count) {
    if (!g || !ref || count != g->n + g->ghosts) {
        set_error("vga_graph_set_cell_refs: count must be cells + ghosts");
        return VGA_ERR_INVALID;
    }
    g->h_refs.assign(ref, ref + count);
    return VGA_OK;
}

int vga_graph_set_noexpand(vga_graph *g, const uint8_t *flags) {
    if (!g) return VGA_ERR_INVALID;
    vga_ctx *ctx = g->ctx;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    if (!flags || g->n == 0) {
        g->noexpand.release();
        return VGA_OK;
    }
    VGA_TRY(g->noexpand.alloc((size_t)g->n));
    VGA_CUDA(cudaMemcpyAsync(g->noexpand.p, flags, (size_t)g->n, cudaMemcpyHostToDevice, ctx->stream));
    VGA_CUDA(cudaStreamSynchronize(ctx->stream));
    return VGA_OK;
}

int vga_graph_node_stats(const vga_graph *g, int32_t *connectivity, double *sum_d, double *sum_d2, float *far_bin_dists,
                         int32_t *bin_count, uint8_t *grid_connections) {
    if (!g) return VGA_ERR_INVALID;
    if (!g->has_stats) {
        set_error("vga_graph_node_stats: only graphs made by vga_graph_build carry node statistics");
        return VGA_ERR_INVALID;
    }
    VGA_CUDA(cudaSetDevice(g->ctx->device));
    g_alloc_stream = g->ctx->stream;
    const int64_t rows = g->src_end - g->src_begin;
    if (rows == 0) return VGA_OK;
    if (connectivity) VGA_CUDA(cudaMemcpy(connectivity, g->connectivity.p, sizeof(int32_t) * rows, cudaMemcpyDeviceToHost));
    if (sum_d) VGA_CUDA(cudaMemcpy(sum_d, g->sum_d.p, sizeof(double) * rows, cudaMemcpyDeviceToHost));
    if (sum_d2) VGA_CUDA(cudaMemcpy(sum_d2, g->sum_d2.p, sizeof(double) * rows, cudaMemcpyDeviceToHost));
    if (far_bin_dists) VGA_CUDA(cudaMemcpy(far_bin_dists, g->far_dist.p, sizeof(float) * rows * 32, cudaMemcpyDeviceToHost));
    if (bin_count) VGA_CUDA(cudaMemcpy(bin_count, g->bin_count.p, sizeof(int32_t) * rows * 32, cudaMemcpyDeviceToHost));
    if (grid_connections) VGA_CUDA(cudaMemcpy(grid_connections, g->gridconn.p, (size_t)rows, cudaMemcpyDeviceToHost));
    return VGA_OK;
}

int vga_graph_device_rows(const vga_graph *g, const uint64_t **d_rowptr, const uint32_t **d_adj, int64_t *n_entries) {
    if (!g) return VGA_ERR_INVALID;
    if (d_rowptr) *d_rowptr = g->rowptr.p;
    if (d_adj) *d_adj = g->adj.p;
    if (n_entries) *n_entries = g->entries;
    return VGA_OK;
}

// ------------------------------------------------------------------------------------ analysis

int vga_global(vga_ctx *ctx, const vga_graph *g, int radius, int64_t src_begin, int64_t src_end, int32_t *total_nodes,
               int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used) {
    if (!ctx || !g) return VGA_ERR_INVALID;
    if (radius < -1) {
        set_error("vga_global: radius must be -1 (n) or >= 0");
        return VGA_ERR_INVALID;
    }
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    return run_global(ctx, const_cast<vga_graph *>(g), radius, nullptr, src_begin, src_end, total_nodes, total_depth, dist,
                      max_levels, levels_used);
}

int vga_global_sources(vga_ctx *ctx, const vga_graph *g, int radius, const int64_t *sources, int64_t n_sources,
                       int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t max_levels, int32_t *levels_used) {
    if (!ctx || !g || n_sources < 0 || (n_sources > 0 && !sources)) return VGA_ERR_INVALID;
    if (radius < -1) {
        set_error("vga_global_sources: radius must be -1 (n) or >= 0");
        return VGA_ERR_INVALID;
    }
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    if (levels_used) *levels_used = 0;
    if (n_sources == 0) return VGA_OK;
    return run_global(ctx, const_cast<vga_graph *>(g), radius, sources, 0, n_sources, total_nodes, total_depth, dist, max_levels,
                      levels_used);
}

int vga_graph_list_sizes(vga_ctx *ctx, const vga_graph *g, int64_t *out_runs, int64_t *out_nodes, int64_t *in_runs, int64_t *in_nodes) {
    if (!ctx || !g) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    VGA_TRY(ensure_bfs_lists(ctx, const_cast<vga_graph *>(g), true));
    if (out_runs) *out_runs = g->f_nruns;
    if (out_nodes) *out_nodes = g->f_nnodes;
    if (in_runs) *in_runs = g->t_nruns;
    if (in_nodes) *in_nodes = g->t_nnodes;
    return VGA_OK;
}

int vga_graph_batch_order(vga_ctx *ctx, const vga_graph *g, int32_t *order) {
    if (!ctx || !g || !order) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    std::vector<int32_t> o;
    VGA_TRY(batch_source_order(ctx, const_cast<vga_graph *>(g), 0, g->n, o));
    std::memcpy(order, o.data(), sizeof(int32_t) * o.size());
    return VGA_OK;
}

int vga_graph_device_runs(vga_ctx *ctx, const vga_graph *g, const uint64_t **d_runptr, const void **d_runs, int64_t *n_runs) {
    if (!ctx || !g) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    vga_graph *gg = const_cast<vga_graph *>(g);
    VGA_TRY(shard_runs(ctx, gg));  // a shard holds the rows [src_begin, src_end): runs of those rows only
    if (d_runptr) *d_runptr = gg->f_runptr.p;
    if (d_runs) *d_runs = gg->f_runs.p;
    if (n_runs) *n_runs = gg->f_nruns;
    return VGA_OK;
}

int vga_graph_runs_alloc(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, int64_t n_runs, vga_graph **out, uint64_t **d_runptr,
                         void **d_runs, uint32_t **d_degree) {
    if (!ctx || !out || n_cells < 0 || n_ghosts < 0 || n_runs < 0) return VGA_ERR_INVALID;
    *out = nullptr;
    if (n_cells + n_ghosts >= ((int64_t)1 << 26)) {
        set_error("more than 2^26 vertices is not supported");
        return VGA_ERR_UNSUPPORTED;
    }
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    std::unique_ptr<vga_graph> g(new vga_graph());
    g->ctx = ctx;
    g->n = n_cells;
    g->ghosts = n_ghosts;
    g->src_begin = 0;
    g->src_end = n_cells;
    g->entries = 0;
    g->runs_only = true;
    VGA_TRY(g->f_runptr.alloc((size_t)n_cells + 1));
    VGA_TRY(g->f_runs.alloc((size_t)n_runs + 1));
    VGA_TRY(g->deg.alloc((size_t)n_cells + 1));
    VGA_CUDA(cudaMemsetAsync(g->deg.p, 0, sizeof(uint32_t) * ((size_t)n_cells + 1), ctx->stream));
    VGA_CUDA(cudaStreamSynchronize(ctx->stream));  // the caller fills the buffers from other streams
    g->f_nruns = n_runs;
    if (d_runptr) *d_runptr = g->f_runptr.p;
    if (d_runs) *d_runs = g->f_runs.p;
    if (d_degree) *d_degree = g->deg.p;
    *out = g.release();
    return VGA_OK;
}

int vga_graph_runs_commit(vga_graph *g) {
    if (!g || !g->runs_only || !g->f_runptr.p) return VGA_ERR_INVALID;
    g->has_fwd_runs = true;
    return VGA_OK;
}

int vga_graph_device_degrees(vga_ctx *ctx, const vga_graph *g, const uint32_t **d_degree) {
    if (!ctx || !g || !d_degree) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    VGA_TRY(row_degrees(ctx, const_cast<vga_graph *>(g)));
    *d_degree = g->deg.p;
    return VGA_OK;
}

int vga_graph_from_device_runs(vga_ctx *ctx, int64_t n_cells, int64_t n_ghosts, const uint64_t *d_runptr, const void *d_runs,
                               int64_t n_runs, const uint32_t *d_degree, vga_graph **out) {
    if (!ctx || !out || !d_runptr || n_cells < 0 || n_ghosts < 0 || n_runs < 0 || (n_runs > 0 && !d_runs)) return VGA_ERR_INVALID;
    *out = nullptr;
    if (n_cells + n_ghosts >= ((int64_t)1 << 26)) {
        set_error("more than 2^26 vertices is not supported");
        return VGA_ERR_UNSUPPORTED;
    }
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    cudaStream_t st = ctx->stream;
    std::unique_ptr<vga_graph> g(new vga_graph());
    g->ctx = ctx;
    g->n = n_cells;
    g->ghosts = n_ghosts;
    g->src_begin = 0;
    g->src_end = n_cells;
    g->entries = 0;  // no entry rows: this graph serves vga_global / vga_global_sources / vga_local only
    g->runs_only = true;
    VGA_TRY(g->f_runptr.alloc((size_t)n_cells + 1));
    VGA_TRY(g->f_runs.alloc((size_t)n_runs + 1));
    VGA_CUDA(cudaMemcpyAsync(g->f_runptr.p, d_runptr, sizeof(uint64_t) * (n_cells + 1), cudaMemcpyDeviceToDevice, st));
    if (n_runs > 0) VGA_CUDA(cudaMemcpyAsync(g->f_runs.p, d_runs, sizeof(uint2) * n_runs, cudaMemcpyDeviceToDevice, st));
    if (d_degree && n_cells > 0) {
        VGA_TRY(g->deg.alloc((size_t)n_cells));
        VGA_CUDA(cudaMemcpyAsync(g->deg.p, d_degree, sizeof(uint32_t) * n_cells, cudaMemcpyDeviceToDevice, st));
    }
    VGA_CUDA(cudaStreamSynchronize(st));
    g->f_nruns = n_runs;
    g->has_fwd_runs = true;
    *out = g.release();
    return VGA_OK;
}

int vga_global_attributes(int64_t n, const int32_t *total_nodes, const int64_t *total_depth, const int32_t *dist,
                          int32_t max_levels, float *node_count, float *mean_depth, float *integ_hh, float *integ_pv,
                          float *integ_tk, float *entropy, float *rel_entropy) {
    if (n < 0 || !total_nodes || !total_depth || (max_levels > 0 && !dist)) return VGA_ERR_INVALID;
    for (int64_t i = 0; i < n; i++) {
        const int nodes = total_nodes[i];
        const int depth = (int)total_depth[i];  // the reference accumulates total_depth in an int
        float md_f = -1.0f, hh = -1.0f, pv = -1.0f, tk = -1.0f, en = -1.0f, re = -1.0f;
        if (nodes > 1) {
            const double md = double(depth) / double(nodes - 1);
            md_f = float(md);
            if (nodes > 2 && md > 1.0) {
                const double ra = 2.0 * (md - 1.0) / double(nodes - 2);
                hh = float(1.0 / (ra / d_value(nodes)));
                pv = float(1.0 / (ra / p_value(nodes)));
                const double t = tekl_integ(nodes, depth);
                tk = (depth - nodes + 1 > 1) ? float(t) : -1.0f;
            }
            double e_sum = 0.0, r_sum = 0.0, fact = 1.0;
            for (int k = 1; k < max_levels; k++) {
                const int c = dist[i * max_levels + k];
                if (c <= 0) continue;
                const double prob = double(c) / double(nodes - 1);
                e_sum -= prob * log2_paf(prob);
                fact *= double(k + 1);
                const double q = (std::pow(md, double(k)) / fact) * std::exp(-md);
                r_sum += (float)prob * log2_paf(prob / q);
            }
            en = float(e_sum);
            re = float(r_sum);
        }
        if (node_count) node_count[i] = float(nodes);
        if (mean_depth) mean_depth[i] = md_f;
        if (integ_hh) integ_hh[i] = hh;
        if (integ_pv) integ_pv[i] = pv;
        if (integ_tk) integ_tk[i] = tk;
        if (entropy) entropy[i] = en;
        if (rel_entropy) rel_entropy[i] = re;
    }
    return VGA_OK;
}

int vga_local(vga_ctx *ctx, const vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k,
              int32_t *total, float *control) {
    if (!ctx || !g) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    return run_local(ctx, const_cast<vga_graph *>(g), src_begin, src_end, cluster, k, total, control);
}

int vga_step_depth(vga_ctx *ctx, const vga_graph *g, const int64_t *sources, int64_t n_sources, int32_t *depth) {
    if (!ctx || !g || !depth || n_sources < 0 || (n_sources > 0 && !sources)) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    return run_step_depth(ctx, const_cast<vga_graph *>(g), sources, n_sources, depth);
}

static int metric_angular_entry(vga_ctx *ctx, const vga_graph *g, int angular, const uint8_t *blocked_adjacent,
                                const int32_t *merge_partner, double spacing,
                                double radius, const int64_t *sources, int64_t n_sources, float *const *out, int nout,
                                int64_t *angle_unsafe) {
    if (!ctx || !g || n_sources < 0) return VGA_ERR_INVALID;
    VGA_CUDA(cudaSetDevice(ctx->device));
    g_alloc_stream = ctx->stream;
    ctx->timing = Timing();
    std::vector<int64_t> all;
    if (!sources) {  // every cell, x-major: the order of the reference's outer loops
        n_sources = g->n;
        all.resize((size_t)g->n);
        for (int64_t i = 0; i < g->n; i++) all[(size_t)i] = i;
        sources = all.data();
    }
    return run_metric_angular(ctx, const_cast<vga_graph *>(g), angular, blocked_adjacent, merge_partner, spacing, radius, sources, n_sources, out,
                              nout, angle_unsafe);
}

int vga_metric(vga_ctx *ctx, const vga_graph *g, const uint8_t *blocked_adjacent, const int32_t *merge_partner, double spacing, double radius,
               const int64_t *sources, int64_t n_sources, float *mean_angle, float *mean_path_dist, float *mean_line_dist,
               float *node_count, int64_t *angle_unsafe) {
    float *out[4] = {mean_angle, mean_path_dist, mean_line_dist, node_count};
    return metric_angular_entry(ctx, g, 0, blocked_adjacent, merge_partner, spacing, radius, sources, n_sources, out, 4, angle_unsafe);
}

int vga_angular(vga_ctx *ctx, const vga_graph *g, const uint8_t *blocked_adjacent, const int32_t *merge_partner, double radius, const int64_t *sources,
                int64_t n_sources, float *mean_depth, float *total_depth, float *node_count, int64_t *angle_unsafe) {
    float *out[3] = {mean_depth, total_depth, node_count};
    return metric_angular_entry(ctx, g, 1, blocked_adjacent, merge_partner, 1.0, radius, sources, n_sources, out, 3, angle_unsafe);
}

int vga_local_attributes(int64_t n, const int64_t *cluster, const int32_t *k, const int32_t *total, const float *control,
                         float *clustering, float *control_out, float *controllability) {
    if (n < 0 || !cluster || !k || !total || !control) return VGA_ERR_INVALID;
    for (int64_t i = 0; i < n; i++) {
        float a = -1.0f, b = -1.0f, c = -1.0f;
        if (k[i] > 1) {
            const double kd = double(k[i]);
            a = float(double((int)cluster[i]) / (kd * (kd - 1.0)));  // `cluster` is an int in the reference
            b = control[i];
            c = float(kd / double(total[i]));
        }
        if (clustering) clustering[i] = a;
        if (control_out) control_out[i] = b;
        if (controllability) controllability[i] = c;
    }
    return VGA_OK;
}

}  // extern "C"
