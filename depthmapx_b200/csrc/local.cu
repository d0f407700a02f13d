// VGA local measures (VGAVisualLocal::run, salalib/vgamodules/vgavisuallocal.cpp:41-81):
//   k       = |N(v)|                      (de-duplicated contents() of v's node, ghosts included)
//   cluster = sum_{u in N(v), u filled} |iter(N(u)) n N(v)|
//   total   = |U_{u in N(v), u filled} iter(N(u))|
//   control = float32 running sum over u in PixelRef order of 1.0f / float(|iter(N(u))|)
//
// The reference does this with O(k) std::find per visited pixel.  Here a CTA owns a cell v and
// keeps two bitmaps over the vertex universe (N cells + ghosts) in shared memory (global scratch
// when the universe exceeds it): membership of N(v), and the running union.  Every neighbour row
// is streamed once with coalesced 32-bit loads; an intersection is a shared-memory bit test and
// a ballot/popc, the union a test-then-atomicOr.  Cells are handed out through an atomic work
// counter because the cost per cell (sum of neighbour degrees) varies by orders of magnitude.
// The float32 control sum is order dependent and is done by one thread per cell in sorted-row
// order with explicit round-to-nearest ops (k_control).
#include <algorithm>

#include "vga_dev.cuh"

namespace vga {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int LTPB = 256;

typedef unsigned long long u64;

struct LocalDev {
    int64_t n, universe;
    const uint64_t *rowptr;
    const uint32_t *adj;
    int64_t src_begin, src_end;
    u64 *work;  // [1] next cell
    uint32_t *gscratch;  // nullptr = shared memory bitmaps
    int words;  // bitmap words (uint32)
    long long *cluster;
    int32_t *k, *total;
};

__global__ void __launch_bounds__(LTPB) k_local(LocalDev d) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ long long s_v;
    __shared__ long long s_cluster;
    __shared__ int s_total;
    uint32_t *bmN = d.gscratch ? d.gscratch + (size_t)blockIdx.x * 2 * d.words : sm;
    uint32_t *bmT = bmN + d.words;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = LTPB / 32;
    while (true) {
        if (threadIdx.x == 0) {
            s_v = (long long)atomicAdd(d.work, 1ULL) + d.src_begin;
            s_cluster = 0;
            s_total = 0;
        }
        __syncthreads();
        const int64_t v = s_v;
        if (v >= d.src_end) break;
        for (int i = threadIdx.x; i < 2 * d.words; i += LTPB) bmN[i] = 0u;
        __syncthreads();
        const uint64_t r0 = d.rowptr[v], r1 = d.rowptr[v + 1];
        for (uint64_t e = r0 + threadIdx.x; e < r1; e += LTPB) {
            uint32_t c = d.adj[e] >> 6;
            atomicOr(&bmN[c >> 5], 1u << (c & 31));
        }
        __syncthreads();
        long long wcl = 0;
        for (uint64_t e = r0 + warp; e < r1; e += nwarps) {
            uint32_t u = d.adj[e] >> 6;
            if (u >= (uint32_t)d.n) continue;  // ghost: not filled, contributes nothing
            const uint64_t q0 = d.rowptr[u], q1 = d.rowptr[u + 1];
            int cnt = 0;
            for (uint64_t f = q0 + lane; f < q1; f += 32) {
                uint32_t c = d.adj[f] >> 6;
                uint32_t bit = 1u << (c & 31);
                cnt += (bmN[c >> 5] & bit) ? 1 : 0;
                if (!(bmT[c >> 5] & bit)) atomicOr(&bmT[c >> 5], bit);
            }
            cnt = __reduce_add_sync(FULL, cnt);
            wcl += cnt;
        }
        if (lane == 0 && wcl) atomicAdd((u64 *)&s_cluster, (u64)wcl);
        __syncthreads();
        int pc = 0;
        for (int i = threadIdx.x; i < d.words; i += LTPB) pc += __popc(bmT[i]);
        pc = __reduce_add_sync(FULL, pc);
        if (lane == 0 && pc) atomicAdd(&s_total, pc);
        __syncthreads();
        if (threadIdx.x == 0) {
            int64_t o = v - d.src_begin;
            d.cluster[o] = s_cluster;
            d.k[o] = (int32_t)(r1 - r0);
            d.total[o] = s_total;
        }
        __syncthreads();
    }
}

__global__ void k_control(int64_t n, const uint64_t *rowptr, const uint32_t *adj, int64_t src_begin, int64_t src_end,
                          float *control) {
    int64_t v = src_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= src_end) return;
    float c = 0.0f;
    for (uint64_t e = rowptr[v]; e < rowptr[v + 1]; e++) {
        uint32_t u = adj[e] >> 6;
        if (u >= (uint32_t)n) continue;
        int retro = (int)(rowptr[u + 1] - rowptr[u]);
        c = __fadd_rn(c, __fdiv_rn(1.0f, (float)retro));
    }
    control[v - src_begin] = c;
}

// ---- run-length variant (local_mode = 3; the kernel for graphs that hold runs only) ---------------------------------------------------------------------------------
// The same per-cell bitmaps, fed by RUNS instead of entries: a neighbour row u is a short list of (first ordinal, length)
// pairs (vga_graph::f_runs, ~35 cells per run on urban plans), |iter(N(u)) n N(v)| is a handful of masked popcounts on the
// bitmap of N(v) and the union a handful of word ORs -- 10-100x fewer operations than one bit test per entry, and the
// only form a rank of a multi-GPU run holds (vga_graph_from_device_runs).  The vertex universe is covered in `passes`
// column ranges of `span` bits when two bitmaps of the whole universe exceed shared memory (two passes at 10^6 cells).
struct LocalRunsDev {
    int64_t n, universe;
    const uint64_t *runptr;
    const uint2 *runs;
    int64_t src_begin, src_end;
    u64 *work;  // [1] next cell
    int passes;
    uint32_t span;  // bits per pass (multiple of 32)
    long long *cluster;
    int32_t *k, *total;
};

// bits [a, b) of word w of a bitmap whose bit 0 is ordinal c0 (a, b already relative to c0)
__device__ __forceinline__ uint32_t word_mask(uint32_t w, uint32_t a, uint32_t b) {
    const uint32_t lo = w << 5, hi = lo + 32u;
    uint32_t m = 0xffffffffu;
    if (a > lo) m &= 0xffffffffu << (a - lo);
    if (b < hi) m &= 0xffffffffu >> (hi - b);
    return m;
}

__global__ void k_local_runs(LocalRunsDev d) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ long long s_v;
    __shared__ unsigned long long s_cluster;
    __shared__ int s_total, s_k;
    const uint32_t words = d.span >> 5;
    uint32_t *bmN = sm, *bmT = sm + words;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    while (true) {
        if (threadIdx.x == 0) {
            s_v = (long long)atomicAdd(d.work, 1ULL) + d.src_begin;
            s_cluster = 0;
            s_total = 0;
            s_k = 0;
        }
        __syncthreads();
        const int64_t v = s_v;
        if (v >= d.src_end) break;
        const uint64_t r0 = d.runptr[v], r1 = d.runptr[v + 1];
        for (int pass = 0; pass < d.passes; pass++) {
            const uint32_t c0 = (uint32_t)pass * d.span;
            const uint32_t c1 = (uint32_t)min((int64_t)c0 + d.span, d.universe);
            for (uint32_t i = threadIdx.x; i < 2 * words; i += blockDim.x) sm[i] = 0u;
            __syncthreads();
            // membership bitmap of N(v) for this column range; k = |N(v)| on the first pass
            int myk = 0;
            for (uint64_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
                const uint2 run = d.runs[r];
                if (pass == 0) myk += (int)run.y;
                const uint32_t a = max(run.x, c0), b = min(run.x + run.y, c1);
                if (a < b)
                    for (uint32_t w = (a - c0) >> 5; w <= (b - 1 - c0) >> 5; w++) atomicOr(&bmN[w], word_mask(w, a - c0, b - c0));
            }
            if (pass == 0) {
                myk = __reduce_add_sync(FULL, myk);
                if (lane == 0 && myk) atomicAdd(&s_k, myk);
            }
            __syncthreads();
            // every filled member u of N(v): one warp per member (member index m strided over the warps; the warp walks v's
            // runs once to turn m into an ordinal), one lane per run of u
            unsigned long long wcl = 0;
            {
                uint64_t r = r0;
                uint32_t before = 0;  // members in the runs before r
                uint2 vr = r < r1 ? d.runs[r] : make_uint2(0xffffffffu, 0u);
                for (uint32_t m = warp;; m += nwarps) {
                    while (r < r1 && m >= before + vr.y) {
                        before += vr.y;
                        r++;
                        vr = r < r1 ? d.runs[r] : make_uint2(0xffffffffu, 0u);
                    }
                    if (r >= r1 || vr.x >= (uint32_t)d.n) break;  // past the last member / ghost members: not filled, no row
                    const uint32_t u = vr.x + (m - before);
                    const uint64_t q0 = d.runptr[u], q1 = d.runptr[u + 1];
                    int cnt = 0;
                    for (uint64_t q = q0 + lane; q < q1; q += 32) {
                        const uint2 run = d.runs[q];
                        const uint32_t a = max(run.x, c0), b = min(run.x + run.y, c1);
                        if (a < b)
                            for (uint32_t w = (a - c0) >> 5; w <= (b - 1 - c0) >> 5; w++) {
                                const uint32_t mk = word_mask(w, a - c0, b - c0);
                                cnt += __popc(bmN[w] & mk);
                                if ((bmT[w] & mk) != mk) atomicOr(&bmT[w], mk);
                            }
                    }
                    cnt = __reduce_add_sync(FULL, cnt);
                    wcl += (unsigned long long)cnt;
                }
            }
            if (lane == 0 && wcl) atomicAdd(&s_cluster, wcl);
            __syncthreads();
            int pc = 0;
            for (uint32_t i = threadIdx.x; i < words; i += blockDim.x) pc += __popc(bmT[i]);
            pc = __reduce_add_sync(FULL, pc);
            if (lane == 0 && pc) atomicAdd(&s_total, pc);
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            const int64_t o = v - d.src_begin;
            d.cluster[o] = (long long)s_cluster;
            d.k[o] = s_k;
            d.total[o] = s_total;
        }
        __syncthreads();
    }
}

// control over runs: entries per row from the CSR offsets or, for a graph that holds runs only, from deg
__global__ void k_control_runs(int64_t n, const uint64_t *runptr, const uint2 *runs, const uint64_t *rowptr, const uint32_t *deg,
                               int64_t src_begin, int64_t src_end, float *control) {
    int64_t v = src_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= src_end) return;
    float c = 0.0f;
    for (uint64_t r = runptr[v]; r < runptr[v + 1]; r++) {
        const uint2 run = runs[r];
        if (run.x >= (uint32_t)n) break;
        for (uint32_t u = run.x; u < run.x + run.y; u++) {
            const int retro = rowptr ? (int)(rowptr[u + 1] - rowptr[u]) : (int)deg[u];
            c = __fadd_rn(c, __fdiv_rn(1.0f, (float)retro));
        }
    }
    control[v - src_begin] = c;
}

// ---- bit-parallel batched variant --------------------------------------------------------------
// 64 spatially coherent cells form a batch (same clusters as the BFS).  With F1[w] = mask of the
// batch's cells s that have w in their row (w ranges over the whole universe, ghosts included):
//   cluster[s] = sum over edges (u -> w), u filled, of  [s in F1[u]] * [s in F1[w]]
//              = per-bit count of  F1[u] & F1[w]  over the rows of all u with F1[u] != 0
//   total[s]   = number of w with bit s in R2[w],  R2[w] = OR over in-edges (u -> w) of F1[u]
// so every adjacency row that matters to the batch is streamed ONCE for up to 64 cells, instead of
// once per (cell, neighbour) pair: work drops from sum_v sum_{u in N(v)} deg(u) to about
// (N/64) * sum_{u in the batch's joint neighbourhood} deg(u).  Counting per source bit is ballot+popc.
struct LocalBatchDev {
    int64_t n, universe;
    const uint64_t *rowptr;
    const uint32_t *adj;
    u64 *f1, *r2;            // [batches][universe]
    u64 *cluster;            // [batches][64]
    int32_t *total;          // [batches][64]
};

__global__ void k_lb_seed(LocalBatchDev d, const int32_t *src, int64_t nsrc) {
    // one warp per source: scatter its bit into F1 of every cell of its row
    int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (i >= nsrc) return;
    int64_t b = i >> 6;
    u64 bit = 1ULL << (i & 63);
    int64_t s = src[i];
    u64 *f1 = d.f1 + b * d.universe;
    for (uint64_t e = d.rowptr[s] + lane; e < d.rowptr[s + 1]; e += 32) atomicOr(&f1[d.adj[e] >> 6], bit);
}

__global__ void __launch_bounds__(LTPB) k_lb_expand(LocalBatchDev d) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const u64 *f1 = d.f1 + (int64_t)b * d.universe;
    u64 *r2 = d.r2 + (int64_t)b * d.universe;
    __shared__ u64 s_cl[64];
    if (threadIdx.x < 64) s_cl[threadIdx.x] = 0ULL;
    __syncthreads();
    u64 c0 = 0, c1 = 0;  // lane l counts source bits l and l+32
    for (int64_t base = (int64_t)blockIdx.x * LTPB; base < d.n; base += (int64_t)gridDim.x * LTPB) {
        int64_t u = base + threadIdx.x;
        u64 f = (u < d.n) ? f1[u] : 0ULL;  // only filled cells (u < n) have rows and contribute
        uint64_t my0 = 0, my1 = 0;
        if (f != 0ULL) {
            my0 = d.rowptr[u];
            my1 = d.rowptr[u + 1];
        }
        unsigned m = __ballot_sync(FULL, f != 0ULL);
        while (m) {
            int src_lane = __ffs(m) - 1;
            m &= m - 1;
            u64 fw = __shfl_sync(FULL, f, src_lane);
            uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            for (uint64_t e = e0; e < e1; e += 32) {
                uint64_t ee = e + lane;
                u64 x = 0ULL;
                if (ee < e1) {
                    uint32_t c = d.adj[ee] >> 6;
                    x = fw & f1[c];
                    u64 add = fw & ~r2[c];
                    if (add) atomicOr(&r2[c], add);
                }
                // per-source counts = column sums of the warp's 32 x 64 bit matrix: one ballot per non-empty column, or,
                // from DENSE_COLUMNS columns on, a 32 x 32 transpose (five shuffles) and one popc per half
                const unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
                unsigned lo_any = __reduce_or_sync(FULL, lo);
                unsigned hi_any = __reduce_or_sync(FULL, hi);
                if (__popc(lo_any) >= DENSE_COLUMNS) {
                    c0 += (u64)__popc(warp_transpose32(lo, lane));
                } else {
                    while (lo_any) {
                        int bit = __ffs(lo_any) - 1;
                        lo_any &= lo_any - 1;
                        int c = __popc(__ballot_sync(FULL, (lo >> bit) & 1u));
                        if (lane == bit) c0 += c;
                    }
                }
                if (__popc(hi_any) >= DENSE_COLUMNS) {
                    c1 += (u64)__popc(warp_transpose32(hi, lane));
                } else {
                    while (hi_any) {
                        int bit = __ffs(hi_any) - 1;
                        hi_any &= hi_any - 1;
                        int c = __popc(__ballot_sync(FULL, (hi >> bit) & 1u));
                        if (lane == bit) c1 += c;
                    }
                }
            }
        }
    }
    if (c0) atomicAdd(&s_cl[lane], c0);
    if (c1) atomicAdd(&s_cl[lane + 32], c1);
    __syncthreads();
    if (threadIdx.x < 64 && s_cl[threadIdx.x]) atomicAdd(&d.cluster[(int64_t)b * 64 + threadIdx.x], s_cl[threadIdx.x]);
}

__global__ void __launch_bounds__(LTPB) k_lb_total(LocalBatchDev d) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const u64 *r2 = d.r2 + (int64_t)b * d.universe;
    __shared__ int s_cnt[64];
    if (threadIdx.x < 64) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    int c0 = 0, c1 = 0;
    for (int64_t base = (int64_t)blockIdx.x * LTPB; base < d.universe; base += (int64_t)gridDim.x * LTPB) {
        int64_t v = base + threadIdx.x;
        u64 x = (v < d.universe) ? r2[v] : 0ULL;
        const unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
        unsigned lo_any = __reduce_or_sync(FULL, lo);
        unsigned hi_any = __reduce_or_sync(FULL, hi);
        if (__popc(lo_any) >= DENSE_COLUMNS) {  // column sums by transposition, as in k_lb_expand
            c0 += __popc(warp_transpose32(lo, lane));
        } else {
            while (lo_any) {
                int bit = __ffs(lo_any) - 1;
                lo_any &= lo_any - 1;
                int c = __popc(__ballot_sync(FULL, (lo >> bit) & 1u));
                if (lane == bit) c0 += c;
            }
        }
        if (__popc(hi_any) >= DENSE_COLUMNS) {
            c1 += __popc(warp_transpose32(hi, lane));
        } else {
            while (hi_any) {
                int bit = __ffs(hi_any) - 1;
                hi_any &= hi_any - 1;
                int c = __popc(__ballot_sync(FULL, (hi >> bit) & 1u));
                if (lane == bit) c1 += c;
            }
        }
    }
    if (c0) atomicAdd(&s_cnt[lane], c0);
    if (c1) atomicAdd(&s_cnt[lane + 32], c1);
    __syncthreads();
    if (threadIdx.x < 64 && s_cnt[threadIdx.x]) atomicAdd(&d.total[(int64_t)b * 64 + threadIdx.x], s_cnt[threadIdx.x]);
}

int run_local_batched(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k,
                      int32_t *total, float *control) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n, ns = src_end - src_begin, U = n + g->ghosts;
    Timing &tm = ctx->timing;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);
    StageTimer dt(ctx, 4, &tm.d2h_ms);
    std::vector<int32_t> order;
    VGA_TRY(batch_source_order(ctx, g, src_begin, src_end, order));
    DevBuf<int32_t> d_order;
    VGA_TRY(d_order.alloc((size_t)ns));
    VGA_CUDA(cudaMemcpyAsync(d_order.p, order.data(), sizeof(int32_t) * ns, cudaMemcpyHostToDevice, st));
    const int64_t nbatch = (ns + 63) / 64;
    size_t free_b = 0, total_b = 0;
    VGA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    int64_t budget = (int64_t)std::min<size_t>((size_t)(free_b * 0.4), (size_t)32 << 30);
    int64_t chunk = std::max<int64_t>(1, budget / (16 * std::max<int64_t>(U, 1)));
    chunk = std::min<int64_t>(std::min<int64_t>(chunk, nbatch), 65535);
    DevBuf<u64> f1, r2, d_cl;
    DevBuf<int32_t> d_tot;
    DevBuf<float> d_control;
    VGA_TRY(f1.alloc((size_t)chunk * U));
    VGA_TRY(r2.alloc((size_t)chunk * U));
    VGA_TRY(d_cl.alloc((size_t)chunk * 64));
    VGA_TRY(d_tot.alloc((size_t)chunk * 64));
    VGA_TRY(d_control.alloc((size_t)ns));
    LocalBatchDev d;
    d.n = n;
    d.universe = U;
    d.rowptr = g->rowptr.p;
    d.adj = g->adj.p;
    d.f1 = f1.p;
    d.r2 = r2.p;
    d.cluster = d_cl.p;
    d.total = d_tot.p;
    std::vector<u64> h_cl;
    std::vector<int32_t> h_tot;
    std::vector<uint64_t> h_rp((size_t)n + 1);
    VGA_CUDA(cudaMemcpyAsync(h_rp.data(), g->rowptr.p, sizeof(uint64_t) * (n + 1), cudaMemcpyDeviceToHost, st));
    const unsigned xb_n = (unsigned)std::min<int64_t>((n + LTPB - 1) / LTPB, 2048);
    const unsigned xb_u = (unsigned)std::min<int64_t>((U + LTPB - 1) / LTPB, 2048);
    for (int64_t b0 = 0; b0 < nbatch; b0 += chunk) {
        const int64_t cb = std::min<int64_t>(chunk, nbatch - b0);
        const int64_t first = b0 * 64;
        const int64_t cs = std::min<int64_t>(cb * 64, ns - first);
        if (ctx->cancel && ctx->cancel(ctx->user)) {
            set_error("cancelled");
            return VGA_ERR_CANCELLED;
        }
        kt.start();
        mt.start();
        VGA_CUDA(cudaMemsetAsync(f1.p, 0, sizeof(u64) * (size_t)cb * U, st));
        VGA_CUDA(cudaMemsetAsync(r2.p, 0, sizeof(u64) * (size_t)cb * U, st));
        VGA_CUDA(cudaMemsetAsync(d_cl.p, 0, sizeof(u64) * (size_t)cb * 64, st));
        VGA_CUDA(cudaMemsetAsync(d_tot.p, 0, sizeof(int32_t) * (size_t)cb * 64, st));
        k_lb_seed<<<(unsigned)((cs * 32 + 255) / 256), 256, 0, st>>>(d, d_order.p + first, cs);
        k_lb_expand<<<dim3(xb_n, (unsigned)cb), LTPB, 0, st>>>(d);
        k_lb_total<<<dim3(xb_u, (unsigned)cb), LTPB, 0, st>>>(d);
        tm.launches += 3;
        tm.main_launches += 3;
        VGA_CUDA(cudaGetLastError());
        mt.stop();
        kt.stop();
        dt.start();
        h_cl.resize((size_t)cb * 64);
        h_tot.resize((size_t)cb * 64);
        VGA_CUDA(cudaMemcpyAsync(h_cl.data(), d_cl.p, sizeof(u64) * cb * 64, cudaMemcpyDeviceToHost, st));
        VGA_CUDA(cudaMemcpyAsync(h_tot.data(), d_tot.p, sizeof(int32_t) * cb * 64, cudaMemcpyDeviceToHost, st));
        dt.stop();
        for (int64_t i = 0; i < cs; i++) {
            const int64_t s = order[(size_t)(first + i)];
            const int64_t o = s - src_begin;
            if (cluster) cluster[o] = (int64_t)h_cl[(size_t)i];
            if (total) total[o] = h_tot[(size_t)i];
            if (k) k[o] = (int32_t)(h_rp[(size_t)s + 1] - h_rp[(size_t)s]);
        }
        if (ctx->progress) ctx->progress(ctx->user, std::min<int64_t>(ns, first + cs), ns);
    }
    kt.start();
    k_control<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(n, g->rowptr.p, g->adj.p, src_begin, src_end, d_control.p);
    tm.launches++;
    VGA_CUDA(cudaGetLastError());
    kt.stop();
    dt.start();
    if (control) VGA_CUDA(cudaMemcpyAsync(control, d_control.p, sizeof(float) * ns, cudaMemcpyDeviceToHost, st));
    dt.stop();
    VGA_CUDA(cudaStreamSynchronize(st));
    return VGA_OK;
}

int run_local_runs(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k, int32_t *total,
                   float *control) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n, ns = src_end - src_begin;
    Timing &tm = ctx->timing;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);
    StageTimer dt(ctx, 4, &tm.d2h_ms);
    StageTimer pt(ctx, 6, &tm.prep_ms);
    pt.start();
    VGA_TRY(ensure_fwd_runs(ctx, g));
    if (g->runs_only && !g->deg.p) {
        set_error("vga_local: a graph adopted from runs needs the per-row entry counts (d_degree)");
        return VGA_ERR_INVALID;
    }
    pt.stop();
    DevBuf<long long> d_cluster;
    DevBuf<int32_t> d_k, d_total;
    DevBuf<float> d_control;
    DevBuf<u64> work;
    VGA_TRY(d_cluster.alloc((size_t)ns));
    VGA_TRY(d_k.alloc((size_t)ns));
    VGA_TRY(d_total.alloc((size_t)ns));
    VGA_TRY(d_control.alloc((size_t)ns));
    VGA_TRY(work.alloc_zero(1, st));
    LocalRunsDev d;
    d.n = n;
    d.universe = n + g->ghosts;
    d.runptr = g->f_runptr.p;
    d.runs = g->f_runs.p;
    d.src_begin = src_begin;
    d.src_end = src_end;
    d.work = work.p;
    d.cluster = d_cluster.p;
    d.k = d_k.p;
    d.total = d_total.p;
    // two bitmaps of `span` bits per CTA within ~200 KB of shared memory
    const size_t smem_cap = std::min<size_t>(ctx->smem_optin, (size_t)200 << 10);
    int64_t max_span = (int64_t)(smem_cap / 8) * 32;  // bits: 2 bitmaps * span / 8 bytes
    if (ctx->opt.local_span > 0) max_span = std::min<int64_t>(max_span, std::max<int64_t>(32, ctx->opt.local_span / 32 * 32));
    d.passes = (int)std::max<int64_t>(1, (d.universe + max_span - 1) / max_span);
    d.span = (uint32_t)(((d.universe + d.passes - 1) / d.passes + 31) / 32 * 32);
    const size_t smem = (size_t)(d.span / 32) * 2 * sizeof(uint32_t);
    // small universes: many CTAs of 256 threads per SM; large ones: one CTA of 1024 threads per SM
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, smem_cap / std::max<size_t>(smem, 1)));
    const int threads = per_sm >= 4 ? 256 : (per_sm >= 2 ? 512 : 1024);
    int blocks = (int)std::min<int64_t>((int64_t)ctx->sm_count * per_sm, ns);
    VGA_CUDA(cudaFuncSetAttribute(k_local_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kt.start();
    mt.start();
    k_local_runs<<<blocks, threads, smem, st>>>(d);
    tm.launches++;
    tm.main_launches++;
    VGA_CUDA(cudaGetLastError());
    mt.stop();
    k_control_runs<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(n, g->f_runptr.p, g->f_runs.p, !g->runs_only ? g->rowptr.p : nullptr,
                                                               !g->runs_only ? nullptr : g->deg.p, src_begin, src_end, d_control.p);
    tm.launches++;
    VGA_CUDA(cudaGetLastError());
    kt.stop();
    dt.start();
    if (cluster) VGA_CUDA(cudaMemcpyAsync(cluster, d_cluster.p, sizeof(int64_t) * ns, cudaMemcpyDeviceToHost, st));
    if (k) VGA_CUDA(cudaMemcpyAsync(k, d_k.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
    if (total) VGA_CUDA(cudaMemcpyAsync(total, d_total.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
    if (control) VGA_CUDA(cudaMemcpyAsync(control, d_control.p, sizeof(float) * ns, cudaMemcpyDeviceToHost, st));
    dt.stop();
    VGA_CUDA(cudaStreamSynchronize(st));
    // algorithmic bytes (local: rowbytes(v) + sum of rowbytes(u) over N(v), no-reuse model) are not
    // accumulated here: bench.py reports cells/s for this stage
    return VGA_OK;
}

}  // namespace

int run_local(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k,
              int32_t *total, float *control) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    if (g->src_begin != 0 || g->src_end != n) {
        set_error("vga_local: the graph must hold the rows of all cells (gather the shards first)");
        return VGA_ERR_INVALID;
    }
    if (src_end < 0 || src_end > n) src_end = n;
    if (src_begin < 0) src_begin = 0;
    const int64_t ns = std::max<int64_t>(0, src_end - src_begin);
    if (ns == 0) return VGA_OK;
    {
        // auto (measured on B200, profiles/README.md): the bit-parallel batches over entries while the graph holds entries
        // (C1 11 / 58 / 120 ms, C5 slice 39 / 54 ms for batches / run-length / entry bitmaps); a graph that holds runs only
        // (the replicated graph of a multi-GPU run) is served by the run-length kernel
        int64_t lm = ctx->opt.local_mode;
        if (lm == 4 || (lm == 2 && local_tc_applicable(g))) {
            // dense graph: cluster / total as a masked int8 product on the tensor cores (C1: 16x faster than the batches
            // with the library GEMM, profiles/r2_tc_probe_C1.json); control in sorted-row order as everywhere
            if (g->runs_only) {
                set_error("vga_local: the tensor-core kernel needs a graph with entry rows");
                return VGA_ERR_INVALID;
            }
            VGA_TRY(run_local_tc(ctx, g, src_begin, src_end, cluster, k, total));
            DevBuf<float> d_control;
            VGA_TRY(d_control.alloc((size_t)ns));
            k_control<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(n, g->rowptr.p, g->adj.p, src_begin, src_end, d_control.p);
            ctx->timing.launches++;
            VGA_CUDA(cudaGetLastError());
            if (control) VGA_CUDA(cudaMemcpyAsync(control, d_control.p, sizeof(float) * ns, cudaMemcpyDeviceToHost, st));
            VGA_CUDA(cudaStreamSynchronize(st));
            return VGA_OK;
        }
        if (lm == 3 || g->runs_only) return run_local_runs(ctx, g, src_begin, src_end, cluster, k, total, control);
        if (lm == 1 || lm == 2) return run_local_batched(ctx, g, src_begin, src_end, cluster, k, total, control);
    }
    Timing &tm = ctx->timing;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);
    StageTimer dt(ctx, 4, &tm.d2h_ms);

    DevBuf<long long> d_cluster;
    DevBuf<int32_t> d_k, d_total;
    DevBuf<float> d_control;
    DevBuf<u64> work;
    DevBuf<uint32_t> scratch;
    VGA_TRY(d_cluster.alloc((size_t)ns));
    VGA_TRY(d_k.alloc((size_t)ns));
    VGA_TRY(d_total.alloc((size_t)ns));
    VGA_TRY(d_control.alloc((size_t)ns));
    VGA_TRY(work.alloc_zero(1, st));

    LocalDev d;
    d.n = n;
    d.universe = n + g->ghosts;
    d.rowptr = g->rowptr.p;
    d.adj = g->adj.p;
    d.src_begin = src_begin;
    d.src_end = src_end;
    d.work = work.p;
    d.words = (int)((d.universe + 31) / 32);
    d.cluster = d_cluster.p;
    d.k = d_k.p;
    d.total = d_total.p;
    size_t smem = (size_t)d.words * 2 * sizeof(uint32_t);
    int blocks;
    const size_t smem_cap = std::min<size_t>(ctx->smem_optin, (size_t)200 << 10);
    if (smem + 64 <= smem_cap) {
        d.gscratch = nullptr;
        int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, smem_cap / std::max<size_t>(smem, 1)));
        blocks = ctx->sm_count * per_sm;
        VGA_CUDA(cudaFuncSetAttribute(k_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    } else {
        blocks = ctx->sm_count * 8;
        VGA_TRY(scratch.alloc((size_t)blocks * 2 * d.words));
        d.gscratch = scratch.p;
        smem = 0;
    }
    blocks = (int)std::min<int64_t>(blocks, ns);

    kt.start();
    mt.start();
    k_local<<<blocks, LTPB, smem, st>>>(d);
    tm.launches++;
    tm.main_launches++;
    VGA_CUDA(cudaGetLastError());
    mt.stop();
    k_control<<<(unsigned)((ns + 127) / 128), 128, 0, st>>>(n, g->rowptr.p, g->adj.p, src_begin, src_end, d_control.p);
    tm.launches++;
    VGA_CUDA(cudaGetLastError());
    kt.stop();
    dt.start();
    if (cluster) VGA_CUDA(cudaMemcpyAsync(cluster, d_cluster.p, sizeof(int64_t) * ns, cudaMemcpyDeviceToHost, st));
    if (k) VGA_CUDA(cudaMemcpyAsync(k, d_k.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
    if (total) VGA_CUDA(cudaMemcpyAsync(total, d_total.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
    if (control) VGA_CUDA(cudaMemcpyAsync(control, d_control.p, sizeof(float) * ns, cudaMemcpyDeviceToHost, st));
    dt.stop();
    VGA_CUDA(cudaStreamSynchronize(st));
    return VGA_OK;
}

}  // namespace vga
