// VGA local measures on the tensor cores (tcgen05, sm_100a) for DENSE visibility graphs.
//
//   cluster(v) = sum_{u in N(v), u filled} |iter(N(u)) n N(v)|   = sum_w A[v,w] * P[v,w]
//   total(v)   = |U_{u in N(v), u filled} iter(N(u))|            = #{w : P[v,w] > 0}
//   with  P = Al * A,  A[u,w] = 1 iff w is in the row of u (w over the universe: cells + ghosts),  Al = A restricted to
//   filled middle vertices u < n                                 (VGAVisualLocal::run, salalib/vgamodules/vgavisuallocal.cpp:41-81)
//
// i.e. a masked row-sum and a non-zero count of the Boolean-weighted square of the adjacency matrix: integer GEMM work.
// Measured with the library GEMM first (tools/local_tc_probe.py, profiles/r2_tc_probe_*.json): on the C1 room plan
// (deg/N = 0.32) the int8 product takes 0.69 ms against 11.2 ms for the bit-parallel batches; on sparse plans (C2,
// deg/N = 0.007) a dense product loses 15x.  This kernel serves the dense regime (run_local picks it from deg/U >= 1/20
// while the bit matrices fit), the batches / run-length kernels serve the rest.
//
// One CTA computes a 128 x 256 tile of P in TMEM (256 columns of int32) with tcgen05.mma.cta_group::1.kind::i8:
//   * operands never exist as bytes in HBM: the adjacency is kept as two bit matrices (bitA[v][w], bitAT[w][u]), and
//     every K chunk of 64 middle vertices is expanded by the CTA's threads straight into shared memory in the canonical
//     K-major no-swizzle UMMA layout (8 x 16-byte core matrices; SBO = 512 B between 8-row groups, LBO = 128 B between
//     the K halves) -- two stages, so the expansion of chunk i+1 overlaps the MMAs of chunk i;
//   * chunks whose A tile or B tile is all zero are skipped (a __syncthreads_or of the 64-bit words just read);
//   * one thread issues the MMAs and tcgen05.commit -> mbarrier; the epilogue reads the accumulator back with
//     tcgen05.ld.32x32b (thread = row), applies the mask word of A and adds cluster / total to global counters.
// Every mbarrier wait is bounded (a wrong descriptor must end in an error, not in a hung GPU).
// Bring-up: tools/tc/umma_i8_test.cu (PASS on B200: profiles/r2_tcgen05_i8_bringup.log).
#include <algorithm>

#include "vga_dev.cuh"

namespace vga {

#ifndef VGA_SIMT_EMULATION

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int TM = 128, TN = 256, KC = 64, STAGES = 2;
constexpr int STAGE_BYTES = (TM + TN) * KC;

typedef unsigned long long u64;

struct LocalTcDev {
    int64_t n, universe;
    const uint32_t *bitA;   // [n][wa] rows v, bits over the universe (zero padded)
    int64_t wa;             // words per row of bitA (multiple of 8 = one column tile)
    const uint32_t *bitAT;  // [upad][wt] rows w, bits over the filled middle vertices u < n (zero padded)
    int64_t wt;             // words per row of bitAT (multiple of 2 = one K chunk)
    int64_t src_begin, src_end;
    u64 *cluster;           // [src_end - src_begin]
    int *total;
    int *error;             // set when a barrier wait gave up
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// SmemDescriptor (sm_100 UMMA): start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48), SWIZZLE_NONE [61,64) = 0
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// InstrDescriptor: D = S32 (2) [4,6), A / B = signed 8 bit (1) [7,10) [10,13), both K-major, N >> 3 [17,23), M >> 4 [24,29)
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 24); spin++) {
        uint32_t ok;
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// 16 bits -> 16 bytes of 0 / 1 (one 16-byte K core row)
__device__ __forceinline__ uint4 expand16(uint32_t bits) {
    uint4 o;
    o.x = ((bits & 0xfu) * 0x00204081u) & 0x01010101u;
    o.y = (((bits >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
    o.z = (((bits >> 8) & 0xfu) * 0x00204081u) & 0x01010101u;
    o.w = (((bits >> 12) & 0xfu) * 0x00204081u) & 0x01010101u;
    return o;
}
// row r of a tile with KC = 64 bytes of K per row, canonical layout: (r / 8) * 512 + core * 128 + (r % 8) * 16
__device__ __forceinline__ void store_row64(uint8_t *tile, uint32_t r, uint32_t lo, uint32_t hi) {
    uint8_t *p = tile + (r >> 3) * (KC / 16 * 128) + (r & 7) * 16;
    *reinterpret_cast<uint4 *>(p) = expand16(lo & 0xffffu);
    *reinterpret_cast<uint4 *>(p + 128) = expand16(lo >> 16);
    *reinterpret_cast<uint4 *>(p + 256) = expand16(hi & 0xffffu);
    *reinterpret_cast<uint4 *>(p + 384) = expand16(hi >> 16);
}

__global__ void __launch_bounds__(TM) k_local_tc(LocalTcDev d) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t mbar[STAGES];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = d.src_begin + (int64_t)blockIdx.x * TM;
    const int64_t col0 = (int64_t)blockIdx.y * TN;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base_s;

    const int64_t v = row0 + tid;
    const bool vrow = v < d.src_end;
    const uint32_t *arow = d.bitA + (vrow ? v : 0) * d.wa;
    const uint32_t *brow0 = d.bitAT + (col0 + tid) * d.wt;
    const uint32_t *brow1 = d.bitAT + (col0 + TM + tid) * d.wt;
    uint32_t phase[STAGES] = {0, 0};
    bool pending[STAGES] = {false, false};  // an MMA group of this stage is in flight (uniform over the CTA)
    bool started = false, ok = true;
    int stage = 0;
    const int64_t nchunks = (d.n + KC - 1) / KC;
    for (int64_t kc = 0; kc < nchunks; kc++) {
        // the 64 middle vertices u of this chunk: bits of row v of A (filled u only: the tail beyond n is masked) ...
        uint32_t alo = 0, ahi = 0;
        if (vrow) {
            alo = arow[kc * 2];
            ahi = arow[kc * 2 + 1];
            const int64_t left = d.n - kc * KC;
            if (left < 64) {
                if (left <= 32) {
                    ahi = 0;
                    alo &= left >= 32 ? 0xffffffffu : ((1u << left) - 1u);
                } else {
                    ahi &= (1u << (left - 32)) - 1u;
                }
            }
        }
        // ... and of the rows w of A^T (bitAT is zero beyond n and beyond the universe)
        const uint32_t b0lo = brow0[kc * 2], b0hi = brow0[kc * 2 + 1], b1lo = brow1[kc * 2], b1hi = brow1[kc * 2 + 1];
        const int anyA = __syncthreads_or((alo | ahi) != 0u);
        const int anyB = __syncthreads_or((b0lo | b0hi | b1lo | b1hi) != 0u);
        if (!anyA || !anyB) continue;  // an empty tile contributes nothing
        if (pending[stage]) {           // the MMAs that read this stage's shared memory must have completed
            ok = mbar_wait(smem_u32(&mbar[stage]), phase[stage]) && ok;
            phase[stage] ^= 1;
            pending[stage] = false;
        }
        uint8_t *sA = smem + stage * STAGE_BYTES, *sB = sA + TM * KC;
        store_row64(sA, (uint32_t)tid, alo, ahi);
        store_row64(sB, (uint32_t)tid, b0lo, b0hi);
        store_row64(sB, (uint32_t)(TM + tid), b1lo, b1hi);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;");
#pragma unroll
            for (int kk = 0; kk < KC / 32; kk++) {
                const uint64_t da = umma_desc(smem_u32(sA) + kk * 256, 128, KC / 16 * 128);
                const uint64_t db = umma_desc(smem_u32(sB) + kk * 256, 128, KC / 16 * 128);
                const uint32_t acc = (started || kk > 0) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                             ::"r"(tmem), "l"(da), "l"(db), "r"(IDESC), "r"(acc), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar[stage])) : "memory");
        }
        started = true;
        pending[stage] = true;
        stage ^= 1;
    }
    for (int s = 0; s < STAGES; s++)
        if (pending[s]) {
            ok = mbar_wait(smem_u32(&mbar[s]), phase[s]) && ok;
            phase[s] ^= 1;
        }
    ok = __syncthreads_and(ok);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (!ok) {
        if (tid == 0) atomicExch(d.error, 1);
    } else if (started) {
        u64 cl = 0;
        int tot = 0;
        const int cmax = (int)min((int64_t)TN, d.universe - col0);  // columns of this tile inside the universe
        for (int c = 0; c < TN; c += 32) {
            if (c >= cmax) break;  // uniform over the CTA
            uint32_t p[32];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
                         "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                         : "=r"(p[0]), "=r"(p[1]), "=r"(p[2]), "=r"(p[3]), "=r"(p[4]), "=r"(p[5]), "=r"(p[6]), "=r"(p[7]), "=r"(p[8]), "=r"(p[9]),
                           "=r"(p[10]), "=r"(p[11]), "=r"(p[12]), "=r"(p[13]), "=r"(p[14]), "=r"(p[15]), "=r"(p[16]), "=r"(p[17]), "=r"(p[18]),
                           "=r"(p[19]), "=r"(p[20]), "=r"(p[21]), "=r"(p[22]), "=r"(p[23]), "=r"(p[24]), "=r"(p[25]), "=r"(p[26]), "=r"(p[27]),
                           "=r"(p[28]), "=r"(p[29]), "=r"(p[30]), "=r"(p[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (vrow) {
                const uint32_t m = arow[(col0 + c) >> 5];  // A[v, col0 + c ..]: the members of N(v) among these columns
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    tot += p[j] != 0u;
                    cl += ((m >> j) & 1u) ? (u64)p[j] : 0ULL;
                }
            }
        }
        if (vrow) {
            if (cl) atomicAdd(&d.cluster[v - d.src_begin], cl);
            if (tot) atomicAdd(&d.total[v - d.src_begin], tot);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TN));
}

// bitA[row][.]: the runs of a row as set bits (one warp per row, one lane per run)
__global__ void k_bits_from_runs(int64_t rows, const uint64_t *runptr, const uint2 *runs, int64_t wpr, uint32_t *bits) {
    const int64_t v = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (v >= rows) return;
    uint32_t *row = bits + v * wpr;
    for (uint64_t r = runptr[v] + lane; r < runptr[v + 1]; r += 32) {
        const uint2 run = runs[r];
        const uint32_t a = run.x, b = run.x + run.y;
        for (uint32_t w = a >> 5; w <= (b - 1) >> 5; w++) {
            uint32_t m = 0xffffffffu;
            if (a > (w << 5)) m &= 0xffffffffu << (a - (w << 5));
            if (b < ((w + 1) << 5)) m &= 0xffffffffu >> (((w + 1) << 5) - b);
            atomicOr(&row[w], m);
        }
    }
}
// bitAT[w][u] = 1 iff w is in the row of the filled cell u (one warp per row u, one lane per run, one atomic per member)
__global__ void k_bits_transposed(int64_t n, const uint64_t *runptr, const uint2 *runs, int64_t wt, uint32_t *bits) {
    const int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    const uint32_t bit = 1u << (u & 31);
    for (uint64_t r = runptr[u] + lane; r < runptr[u + 1]; r += 32) {
        const uint2 run = runs[r];
        for (uint32_t w = run.x; w < run.x + run.y; w++) atomicOr(&bits[(int64_t)w * wt + (u >> 5)], bit);
    }
}
// k = |N(v)| = members of the runs of the row
__global__ void k_row_members(int64_t src_begin, int64_t src_end, const uint64_t *runptr, const uint2 *runs, int32_t *k) {
    const int64_t v = src_begin + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= src_end) return;
    int c = 0;
    for (uint64_t r = runptr[v]; r < runptr[v + 1]; r++) c += (int)runs[r].y;
    k[v - src_begin] = c;
}

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

}  // namespace

bool local_tc_applicable(const vga_graph *g) {
    if (g->runs_only || g->n <= 0) return false;
    const double U = (double)(g->n + g->ghosts);
    return (double)g->n * U <= 3.5e10 /* two bit matrices of <= 4.4 GB */ && (double)g->entries * 20.0 >= (double)g->n * U;
}

int run_local_tc(vga_ctx *ctx, vga_graph *g, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k, int32_t *total) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n, U = n + g->ghosts, ns = src_end - src_begin;
    Timing &tm = ctx->timing;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);
    StageTimer dt(ctx, 4, &tm.d2h_ms);
    StageTimer pt(ctx, 6, &tm.prep_ms);
    pt.start();
    VGA_TRY(ensure_fwd_runs(ctx, g));
    const int64_t wa = (std::max<int64_t>(U, (n + KC - 1) / KC * KC) + 255) / 256 * 8;  // words: whole column tiles and K chunks
    const int64_t wt = (n + KC - 1) / KC * 2;
    const int64_t upad = (U + TN - 1) / TN * TN;
    if (!g->tc_bits_a.p) {
        VGA_TRY(g->tc_bits_a.alloc_zero((size_t)(n * wa), st));
        VGA_TRY(g->tc_bits_t.alloc_zero((size_t)(upad * wt), st));
        k_bits_from_runs<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, g->f_runptr.p, g->f_runs.p, wa, g->tc_bits_a.p);
        k_bits_transposed<<<blocks_for(n * 32, 256), 256, 0, st>>>(n, g->f_runptr.p, g->f_runs.p, wt, g->tc_bits_t.p);
        tm.launches += 2;
        VGA_CUDA(cudaGetLastError());
    }
    pt.stop();
    DevBuf<u64> d_cluster;
    DevBuf<int32_t> d_total, d_k;
    DevBuf<int> d_err;
    VGA_TRY(d_cluster.alloc_zero((size_t)ns, st));
    VGA_TRY(d_total.alloc_zero((size_t)ns, st));
    VGA_TRY(d_k.alloc((size_t)ns));
    VGA_TRY(d_err.alloc_zero(1, st));
    LocalTcDev d;
    d.n = n;
    d.universe = U;
    d.bitA = g->tc_bits_a.p;
    d.wa = wa;
    d.bitAT = g->tc_bits_t.p;
    d.wt = wt;
    d.src_begin = src_begin;
    d.src_end = src_end;
    d.cluster = d_cluster.p;
    d.total = d_total.p;
    d.error = d_err.p;
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 1024;
    VGA_CUDA(cudaFuncSetAttribute(k_local_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kt.start();
    mt.start();
    k_local_tc<<<dim3(blocks_for(ns, TM), (unsigned)(upad / TN)), TM, smem, st>>>(d);
    tm.launches++;
    tm.main_launches++;
    VGA_CUDA(cudaGetLastError());
    mt.stop();
    k_row_members<<<blocks_for(ns, 128), 128, 0, st>>>(src_begin, src_end, g->f_runptr.p, g->f_runs.p, d_k.p);
    tm.launches++;
    kt.stop();
    int h_err = 0;
    dt.start();
    VGA_CUDA(cudaMemcpyAsync(&h_err, d_err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (cluster) VGA_CUDA(cudaMemcpyAsync(cluster, d_cluster.p, sizeof(int64_t) * ns, cudaMemcpyDeviceToHost, st));
    if (k) VGA_CUDA(cudaMemcpyAsync(k, d_k.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
    if (total) VGA_CUDA(cudaMemcpyAsync(total, d_total.p, sizeof(int32_t) * ns, cudaMemcpyDeviceToHost, st));
    dt.stop();
    VGA_CUDA(cudaStreamSynchronize(st));
    if (h_err) {
        set_error("vga_local (tensor-core kernel): a tcgen05 completion barrier timed out");
        return VGA_ERR_CUDA;
    }
    return VGA_OK;
}

#else  // the SIMT emulation of the tests has no tensor cores: the other local kernels serve every graph there

bool local_tc_applicable(const vga_graph *) { return false; }
int run_local_tc(vga_ctx *, vga_graph *, int64_t, int64_t, int64_t *, int32_t *, int32_t *) {
    set_error("the tensor-core local kernel needs a B200");
    return VGA_ERR_UNSUPPORTED;
}

#endif

}  // namespace vga
