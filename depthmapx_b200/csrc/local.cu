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
