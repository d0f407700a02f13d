// Visual step depth: VGAVisualGlobalDepth::run,
// salalib/vgamodules/vgavisualglobaldepth.cpp:23-75 -- level-synchronous BFS from a SET of cells over
// the visibility graph; every filled cell gets the level at which it is first reached, -1 if never.
// One int32 per vertex; a level is one launch: every vertex of the current level streams its row
// with the whole warp and stamps unreached targets with level+1 (all writers of a word write the same
// value, so no atomics are needed).  Ghost columns (unfilled cells) are skipped as the reference skips
// them (p.filled() test, :45).
// Parity: tests/test_stepdepth_gpu.py (reference golden column, oracle, vga_global histogram).
#include <algorithm>

#include "vga_dev.cuh"

namespace vga {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int STPB = 256;

__global__ void k_sd_seed(int32_t *depth, const int64_t *src, int64_t nsrc) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nsrc) depth[src[i]] = 0;
}

__global__ void __launch_bounds__(STPB) k_sd_expand(int64_t n, const uint64_t *rowptr, const uint32_t *adj, int32_t *depth,
                                                    int level, int *any, const uint8_t *noexpand) {
    const int lane = threadIdx.x & 31;
    bool wrote = false;
    for (int64_t base = (int64_t)blockIdx.x * STPB; base < n; base += (int64_t)gridDim.x * STPB) {
        int64_t u = base + threadIdx.x;
        // context-filled cells that are not even are marked but not expanded beyond level 0 (vgavisualglobaldepth.cpp:53)
        bool mine = (u < n) && depth[u] == level && !(level > 0 && noexpand && noexpand[u]);
        uint64_t my0 = 0, my1 = 0;
        if (mine) {
            my0 = rowptr[u];
            my1 = rowptr[u + 1];
        }
        unsigned m = __ballot_sync(FULL, mine);
        while (m) {
            int src_lane = __ffs(m) - 1;
            m &= m - 1;
            uint64_t e0 = __shfl_sync(FULL, my0, src_lane), e1 = __shfl_sync(FULL, my1, src_lane);
            for (uint64_t e = e0 + lane; e < e1; e += 32) {
                uint32_t c = adj[e] >> 6;
                if (c < (uint32_t)n && depth[c] == -1) {
                    depth[c] = level + 1;
                    wrote = true;
                }
            }
        }
    }
    if (__any_sync(FULL, wrote) && lane == 0) *any = 1;
}

}  // namespace

int run_step_depth(vga_ctx *ctx, vga_graph *g, const int64_t *sources, int64_t nsrc, int32_t *depth_out) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    if (g->src_begin != 0 || g->src_end != n) {
        set_error("vga_step_depth: the graph must hold the rows of all cells");
        return VGA_ERR_INVALID;
    }
    for (int64_t i = 0; i < nsrc; i++)
        if (sources[i] < 0 || sources[i] >= n) {
            set_error("vga_step_depth: source ordinal out of range");
            return VGA_ERR_INVALID;
        }
    if (n == 0) return VGA_OK;
    Timing &tm = ctx->timing;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    DevBuf<int32_t> depth;
    DevBuf<int64_t> d_src;
    DevBuf<int> any;
    VGA_TRY(depth.alloc((size_t)n));
    VGA_TRY(d_src.alloc((size_t)std::max<int64_t>(nsrc, 1)));
    VGA_TRY(any.alloc(1));
    VGA_CUDA(cudaMemsetAsync(depth.p, 0xff, sizeof(int32_t) * n, st));  // -1 everywhere
    kt.start();
    if (nsrc > 0) {
        VGA_CUDA(cudaMemcpyAsync(d_src.p, sources, sizeof(int64_t) * nsrc, cudaMemcpyHostToDevice, st));
        k_sd_seed<<<(unsigned)((nsrc + 255) / 256), 256, 0, st>>>(depth.p, d_src.p, nsrc);
        tm.launches++;
        const unsigned blocks = (unsigned)std::min<int64_t>((n + STPB - 1) / STPB, 148 * 16);
        for (int level = 0;; level++) {
            VGA_CUDA(cudaMemsetAsync(any.p, 0, sizeof(int), st));
            k_sd_expand<<<blocks, STPB, 0, st>>>(n, g->rowptr.p, g->adj.p, depth.p, level, any.p, g->noexpand.p);
            tm.launches++;
            tm.main_launches++;
            int h_any = 0;
            VGA_CUDA(cudaMemcpyAsync(&h_any, any.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            VGA_CUDA(cudaStreamSynchronize(st));
            if (!h_any) break;
            if (ctx->cancel && ctx->cancel(ctx->user)) {
                set_error("cancelled");
                return VGA_ERR_CANCELLED;
            }
        }
    }
    kt.stop();
    VGA_CUDA(cudaGetLastError());
    VGA_CUDA(cudaMemcpyAsync(depth_out, depth.p, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    return VGA_OK;
}

}  // namespace vga
