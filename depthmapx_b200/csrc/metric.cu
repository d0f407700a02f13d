// Metric and angular VGA: VGAMetric::run salalib/vgamodules/vgametric.cpp:25-136, VGAAngular::run
// salalib/vgamodules/vgaangular.cpp:22-133, Node/Bin::extractMetric / extractAngular salalib/ngraph.cpp:67-87, 329-365,
// MetricTriple / AngularTriple ordering salalib/pointdata.h:377-419, dist / angle salalib/pixelref.h:116-131.
//
// The reference runs, per source, a Dijkstra-like search over the iterated adjacency with a std::set of (key, pixel,
// last pixel) ordered by (key, pixel): key = float32 path length in cells (metric) or float32 cumulated turn in units of
// 90 degrees (angular).  A popped pixel is expanded only if its key is 0 or it is blocked / has a blocked cell among its
// eight neighbours (corners are where shortest paths turn), and every popped filled cell adds to float32 running sums
// IN POP ORDER -- so the order of the pops, the float rounding of every relaxation and the predecessor that the popped
// element carries are all part of the result.
//
// Here: ONE WARP PER SOURCE, thousands of sources in flight.  The warp keeps the reference's per-pixel state
// (m_dist, m_cumangle) plus, per vertex, the smallest key ever queued and the predecessor queued with it: the set holds
// several elements per pixel, but only the smallest is ever acted on (later ones find the pixel finalised), and of two
// elements with an equal (key, pixel) the set keeps the first -- which is what "replace only when strictly smaller" gives.
// The queue is an indexed 32-ary min-heap of (key bits << 32 | ordinal) with decrease-key: PixelRef order is ordinal order,
// keys are non-negative floats, so the 64-bit integer order is the set's order; the 32 children of a node are read by the 32
// lanes and reduced with redux.sync, so a pop costs three levels at 10^5 cells.  A row is relaxed by the whole warp over
// the RUN form of the out-rows (consecutive ordinals -> the 16-byte vertex states are read coalesced); the cells of 32
// runs are dealt to the lanes through a shuffle binary search so no lane idles on short runs.  Unfilled cells in the gaps
// of diagonal bins (ghost columns) are queued by the reference but never expanded or counted: they are skipped.
//
// Arithmetic is the reference's operation for operation (float + double comparisons, float sums, IEEE sqrt / div /
// mul with explicit rounding, no FMA).  The one non-IEEE function is acos in angle(): CUDA's and glibc's double results may
// differ in the last bits, which changes the float32 turn only when the double lies within a few ulps of a float rounding
// boundary (probability ~1e-8 per evaluation); such evaluations are COUNTED and returned (`angle_unsafe`), so a caller knows
// when bit-equality with the reference is guaranteed (0) -- distances and node counts never depend on it in the metric
// analysis.  Merge links (Point::m_merge) are followed as the reference does: the partner of a popped cell is expanded from
// the same key and finalised without being counted.
// Parity: tests/test_gpu_metric.py (oracle = oracle/vga_oracle.c vgao_metric / vgao_angular, pinned against the reference).
#include <algorithm>

#include "vga_dev.cuh"

namespace vga {

namespace {

using u64 = unsigned long long;
constexpr unsigned FULL = 0xffffffffu;
constexpr uint32_t NONE = 0xffffffffu;
constexpr uint32_t KEY_INF = 0x7f800000u;     // never queued
constexpr uint32_t KEY_POPPED = 0xffffffffu;  // finalised (Point::m_misc == ~0)
constexpr int MWARPS = 4;                     // warps (= sources in flight) per CTA
constexpr int MU = 4;                         // chunks of 32 row cells whose states are in flight together

struct MaArgs {
    int64_t n;
    const uint64_t *runptr;
    const uint2 *runs;
    const int32_t *refs;     // packed PixelRef per ordinal
    const uint8_t *expand;   // [n] blocked or blocked-adjacent
    const int32_t *partner;  // [n] merge partner (Point::m_merge), -1 = none; nullptr = no merge links
    const int64_t *sources;  // [nsrc] ordinals
    int64_t nsrc;
    double spacing, radius;
    int angular;
    uint4 *state;   // [slots][n]: x = m_dist bits, y = m_cumangle bits, z = smallest queued key bits, w = its predecessor
    uint32_t *pos;  // [slots][n] heap position of a queued vertex
    u64 *heap;      // [slots][n]
    unsigned long long *next_source;
    float *out;     // [4][nsrc]
    unsigned long long *unsafe;
};

__device__ __forceinline__ int ref_x(int32_t r) { return (int)(int16_t)(r >> 16); }
__device__ __forceinline__ int ref_y(int32_t r) { return (int)(int16_t)(r & 0xffff); }

// pixelref.h:116-119 (sqr of ints is an int)
__device__ __forceinline__ double cell_dist(int dx, int dy) { return __dsqrt_rn((double)(dx * dx + dy * dy)); }

// (float)(angle(a, b, c) / (M_PI * 0.5)), pixelref.h:121-131, ab = a - b, bc = b - c; counts evaluations whose float
// rounding could differ under a few ulps of error of acos
__device__ __forceinline__ float turn_angle(int abx, int aby, int bcx, int bcy, double nbc, unsigned long long *unsafe) {
    const double q = __ddiv_rn((double)(abx * bcx + aby * bcy), __dadd_rn(__dmul_rn(cell_dist(abx, aby), nbc), 1e-12));
    const double t = __ddiv_rn(acos(q), 1.5707963267948966 /* M_PI * 0.5 */);
    const float f = __double2float_rn(t);
    if (f > 0.0f) {
        const double up = 0.5 * ((double)f + (double)__int_as_float(__float_as_int(f) + 1));
        const double dn = 0.5 * ((double)f + (double)__int_as_float(__float_as_int(f) - 1));
        const double tol = t * 1.8e-15;  // 8 ulps of double
        if (fabs(t - up) <= tol || fabs(t - dn) <= tol) atomicAdd(unsafe, 1ULL);
    }
    return f;
}

// 32-ary indexed min-heap.  Lane 0 is the only writer; the other lanes read children in heap_pop, and every batch of
// mutations ends with __syncwarp() (lanes of a warp are not guaranteed to run in lockstep).
__device__ __forceinline__ void heap_up(u64 *heap, uint32_t *pos, uint32_t i, u64 x) {
    while (i > 0) {
        const uint32_t p = (i - 1) >> 5;
        const u64 hp = heap[p];
        if (x < hp) {
            heap[i] = hp;
            pos[(uint32_t)hp] = i;
            i = p;
        } else {
            break;
        }
    }
    heap[i] = x;
    pos[(uint32_t)x] = i;
}

__device__ __forceinline__ u64 heap_pop(u64 *heap, uint32_t *pos, uint32_t &size, int lane) {
    const u64 top = heap[0];
    size--;
    const u64 last = heap[size];
    __syncwarp();  // every lane has read the root before lane 0 may overwrite it
    if (size > 0) {
        uint32_t i = 0;
        for (;;) {
            const uint32_t c0 = 32 * i + 1;
            if (c0 >= size) break;
            const uint32_t idx = c0 + lane;
            const u64 mine = idx < size ? heap[idx] : ~0ULL;
            const uint32_t hi = (uint32_t)(mine >> 32);
            const uint32_t mh = __reduce_min_sync(FULL, hi);
            const uint32_t lo = hi == mh ? (uint32_t)mine : 0xffffffffu;
            const uint32_t ml = __reduce_min_sync(FULL, lo);
            const u64 cmin = ((u64)mh << 32) | ml;
            if (cmin < last) {
                const int which = __ffs(__ballot_sync(FULL, mine == cmin)) - 1;
                if (lane == 0) {
                    heap[i] = cmin;
                    pos[ml] = i;
                }
                i = c0 + (uint32_t)which;
            } else {
                break;
            }
        }
        if (lane == 0) {
            heap[i] = last;
            pos[(uint32_t)last] = i;
        }
    }
    __syncwarp();
    return top;
}

// Node::extractMetric / extractAngular of the pixel uu popped with key k: every cell of its row that is not finalised is
// relaxed (Bin::extractMetric / extractAngular, ngraph.cpp:329-365); cum_u = m_cumangle of uu, pred = the last pixel of the
// popped element (NONE = NoPixel: no turn).
__device__ __forceinline__ void expand_row(const MaArgs &a, uint4 *state, uint32_t *pos, u64 *heap, uint32_t &size, int lane,
                                           bool angular, uint32_t uu, float k, float cum_u, uint32_t pred) {
    const int64_t n = a.n;
    const int32_t ru = a.refs[uu];
    const int ux = ref_x(ru), uy = ref_y(ru);
    int bcx = 0, bcy = 0;
    double nbc = 0.0;
    if (pred != NONE) {
        const int32_t rp = a.refs[pred];
        bcx = ux - ref_x(rp);
        bcy = uy - ref_y(rp);
        nbc = cell_dist(bcx, bcy);
    }
    const uint64_t r0 = a.runptr[uu], r1 = a.runptr[uu + 1];
    for (uint64_t rb = r0; rb < r1; rb += 32) {
        // 32 runs, their cells dealt to the lanes
        uint2 run = make_uint2(0u, 0u);
        if (rb + lane < r1) run = a.runs[rb + lane];
        if (run.x >= (uint32_t)n) run.y = 0u;  // ghost columns: queued by the reference, never expanded or counted
        uint32_t incl = run.y;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += t;
        }
        const uint32_t excl = incl - run.y;
        const uint32_t cells = __shfl_sync(FULL, incl, 31);
        // MU chunks of 32 cells per round: the vertex states and coordinates of all of them are requested before the first is
        // used (the search is latency bound: one dependent pop -> row -> state chain per warp)
        for (uint32_t cb = 0; cb < cells; cb += 32 * MU) {
            uint32_t vv[MU];
            uint4 svv[MU];
            int32_t rvv[MU];
            bool actv[MU];
#pragma unroll
            for (int j = 0; j < MU; j++) {
                const uint32_t c = cb + 32 * j + lane;
                int lo = 0, hi = 32;  // last lane whose first cell index is <= c
#pragma unroll
                for (int it = 0; it < 5; it++) {
                    const int mid = (lo + hi) >> 1;
                    const uint32_t ev = __shfl_sync(FULL, excl, mid);
                    if (ev <= c) lo = mid; else hi = mid;
                }
                const uint32_t first = __shfl_sync(FULL, run.x, lo);
                const uint32_t eoff = __shfl_sync(FULL, excl, lo);
                actv[j] = c < cells;
                vv[j] = first + (c - eoff);
                svv[j] = make_uint4(0u, 0u, KEY_POPPED, 0u);
                rvv[j] = 0;
                if (actv[j]) {
                    svv[j] = state[vv[j]];
                    rvv[j] = a.refs[vv[j]];
                }
            }
#pragma unroll
            for (int j = 0; j < MU; j++) {
                if (cb + 32 * j >= cells) break;  // uniform
                const uint32_t v = vv[j];
                bool queue = false;
                uint32_t newkey = 0, oldkey = 0;
                uint4 sv = svv[j];
                if (actv[j] && sv.z != KEY_POPPED) {
                    const int32_t rv = rvv[j];
                    const int abx = ref_x(rv) - ux, aby = ref_y(rv) - uy;
                    bool relax;
                    float nk = 0.0f, ncum = 0.0f;
                    if (!angular) {
                        const double w = cell_dist(abx, aby);
                        const float m = __uint_as_float(sv.x);
                        // pt.m_dist == -1.0 || curs.dist + dist(pix, curs.pixel) < pt.m_dist   (float + double)
                        relax = m == -1.0f || __dadd_rn((double)k, w) < (double)m;
                        if (relax) {
                            nk = __fadd_rn(k, __double2float_rn(w));
                            const float ang = pred == NONE ? 0.0f : turn_angle(abx, aby, bcx, bcy, nbc, a.unsafe);
                            ncum = __fadd_rn(cum_u, ang);
                            sv.x = __float_as_uint(nk);
                        }
                    } else {
                        const float ang = pred == NONE ? 0.0f : turn_angle(abx, aby, bcx, bcy, nbc, a.unsafe);
                        const float m = __uint_as_float(sv.y);
                        // pt.m_cumangle == -1.0 || curs.angle + ang < pt.m_cumangle   (float + float)
                        relax = m == -1.0f || __fadd_rn(k, ang) < m;
                        if (relax) {
                            ncum = __fadd_rn(cum_u, ang);
                            nk = ncum;
                        }
                    }
                    if (relax) {
                        sv.y = __float_as_uint(ncum);
                        oldkey = sv.z;
                        newkey = __float_as_uint(nk);
                        // the set keeps the first of two equal (key, pixel) elements and pops the smallest first
                        if (oldkey == KEY_INF || nk < __uint_as_float(oldkey)) {
                            sv.z = newkey;
                            sv.w = uu;
                            queue = true;
                        }
                        state[v] = sv;
                    }
                }
                unsigned qm = __ballot_sync(FULL, queue);
                while (qm) {
                    const int l = __ffs(qm) - 1;
                    qm &= qm - 1;
                    const uint32_t qv = __shfl_sync(FULL, v, l);
                    const uint32_t qk = __shfl_sync(FULL, newkey, l);
                    const uint32_t qo = __shfl_sync(FULL, oldkey, l);
                    const uint32_t fresh_at = size;
                    if (qo == KEY_INF) size++;
                    if (lane == 0) heap_up(heap, pos, qo == KEY_INF ? fresh_at : pos[qv], ((u64)qk << 32) | qv);
                }
            }
        }
    }
    __syncwarp();  // state and heap writes of this row before anything reads them
}

__global__ void __launch_bounds__(MWARPS * 32, 6) k_metric_angular(MaArgs a) {
    const int lane = threadIdx.x & 31;
    const int64_t slot = (int64_t)blockIdx.x * MWARPS + (threadIdx.x >> 5);
    const int64_t n = a.n;
    uint4 *state = a.state + slot * n;
    uint32_t *pos = a.pos + slot * n;
    u64 *heap = a.heap + slot * n;
    const bool angular = a.angular != 0;
    for (;;) {
        unsigned long long si = 0;
        if (lane == 0) si = atomicAdd(a.next_source, 1ULL);
        si = __shfl_sync(FULL, si, 0);
        if ((int64_t)si >= a.nsrc) break;
        const uint32_t src = (uint32_t)a.sources[si];
        const int32_t rsrc = a.refs[src];
        // Point::m_misc = 0, m_dist = -1 (metric) / 0, m_cumangle = 0 (metric) / -1 (angular)
        const uint4 fresh = make_uint4(__float_as_uint(-1.0f), __float_as_uint(angular ? -1.0f : 0.0f), KEY_INF, NONE);
        for (int64_t v = lane; v < n; v += 32) state[v] = fresh;
        __syncwarp();
        uint32_t size = 1;
        if (lane == 0) {
            heap[0] = (u64)src;  // key 0.0f
            pos[src] = 0;
            uint4 s0 = fresh;
            s0.z = 0u;
            if (angular) s0.y = 0u;  // map.getPoint(curs).m_cumangle = 0.0f
            state[src] = s0;
        }
        __syncwarp();
        float total_depth = 0.0f, total_angle = 0.0f, euclid_depth = 0.0f;
        int total_nodes = 0;
        while (size > 0) {
            const u64 top = heap_pop(heap, pos, size, lane);
            const uint32_t u = (uint32_t)top;
            const float k = __uint_as_float((uint32_t)(top >> 32));
            if (a.radius != -1.0 && (angular ? (double)k : __dmul_rn((double)k, a.spacing)) > a.radius) break;
            const uint4 su = state[u];
            if (su.z == KEY_POPPED) continue;  // finalised as the merge partner of an earlier pop (m_misc == ~0)
            const float cum_u = __uint_as_float(su.y);
            if (k == 0.0f || a.expand[u]) expand_row(a, state, pos, heap, size, lane, angular, u, k, cum_u, su.w);
            __syncwarp();  // every lane has read state[u]
            if (lane == 0) state[u].z = KEY_POPPED;
            __syncwarp();
            // merge link (vgametric.cpp:96-104, vgaangular.cpp:91-99): the partner takes over the cumulated angle and is
            // expanded from the same key with no last pixel, then finalised without being counted
            const int32_t u2 = a.partner ? a.partner[u] : -1;
            if (u2 >= 0 && state[u2].z != KEY_POPPED) {
                __syncwarp();  // every lane has read the partner's state
                if (k == 0.0f || a.expand[u2]) expand_row(a, state, pos, heap, size, lane, angular, (uint32_t)u2, k, cum_u, NONE);
                if (lane == 0) {
                    state[u2].y = __float_as_uint(cum_u);
                    state[u2].z = KEY_POPPED;
                }
                __syncwarp();
            }
            const int32_t ru = a.refs[u];
            if (!angular) {
                total_depth = __fadd_rn(total_depth, __double2float_rn(__dmul_rn((double)k, a.spacing)));
                euclid_depth = __fadd_rn(euclid_depth, __double2float_rn(__dmul_rn(a.spacing, cell_dist(ref_x(ru) - ref_x(rsrc), ref_y(ru) - ref_y(rsrc)))));
            }
            total_angle = __fadd_rn(total_angle, cum_u);
            total_nodes += 1;
        }
        if (lane == 0) {
            const double tn = (double)total_nodes;
            if (!angular) {
                a.out[0 * a.nsrc + si] = __double2float_rn(__ddiv_rn((double)total_angle, tn));
                a.out[1 * a.nsrc + si] = __double2float_rn(__ddiv_rn((double)total_depth, tn));
                a.out[2 * a.nsrc + si] = __double2float_rn(__ddiv_rn((double)euclid_depth, tn));
                a.out[3 * a.nsrc + si] = (float)total_nodes;
            } else {
                a.out[0 * a.nsrc + si] = total_nodes > 0 ? __double2float_rn(__ddiv_rn((double)total_angle, tn)) : -1.0f;
                a.out[1 * a.nsrc + si] = total_angle;
                a.out[2 * a.nsrc + si] = (float)total_nodes;
            }
        }
    }
}

}  // namespace

int run_metric_angular(vga_ctx *ctx, vga_graph *g, int angular, const uint8_t *expand, const int32_t *partner, double spacing, double radius,
                       const int64_t *sources, int64_t nsrc, float *const *out, int nout, int64_t *angle_unsafe) {
    cudaStream_t st = ctx->stream;
    const int64_t n = g->n;
    const char *who = angular ? "vga_angular" : "vga_metric";
    if (g->src_begin != 0 || g->src_end != n) {
        set_error(std::string(who) + ": the graph must hold the rows of all cells");
        return VGA_ERR_INVALID;
    }
    if ((int64_t)g->h_refs.size() < n) {
        set_error(std::string(who) + ": the graph has no cell coordinates (vga_graph_set_cell_refs)");
        return VGA_ERR_INVALID;
    }
    if (!expand && n > 0) {
        set_error(std::string(who) + ": the blocked / blocked-adjacent flags are required");
        return VGA_ERR_INVALID;
    }
    for (int64_t i = 0; i < nsrc; i++)
        if (sources[i] < 0 || sources[i] >= n) {
            set_error(std::string(who) + ": source ordinal out of range");
            return VGA_ERR_INVALID;
        }
    if (angle_unsafe) *angle_unsafe = 0;
    if (nsrc <= 0 || n == 0) return VGA_OK;
    VGA_TRY(ensure_fwd_runs(ctx, g));
    Timing &tm = ctx->timing;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);
    // sources in flight: 28 bytes per (slot, vertex); bounded by the resident warps (6 CTAs of 4 warps per SM) and by 40 % of
    // the free memory
    size_t free_b = 0, total_b = 0;
    VGA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const int64_t per_slot = 28 * n;
    int64_t slots = std::min<int64_t>((int64_t)ctx->sm_count * 6 * MWARPS, (int64_t)((double)free_b * 0.4) / std::max<int64_t>(per_slot, 1));
    if (ctx->opt.metric_slots > 0) slots = ctx->opt.metric_slots;
    slots = std::max<int64_t>(MWARPS, std::min<int64_t>(slots, (nsrc + MWARPS - 1) / MWARPS * MWARPS));
    slots = slots / MWARPS * MWARPS;
    uint4 *state = nullptr;
    uint32_t *pos = nullptr;
    u64 *heap = nullptr;
    VGA_TRY(ctx->ws.get("ma_state", sizeof(uint4) * (size_t)slots * (size_t)n, (void **)&state));
    VGA_TRY(ctx->ws.get("ma_pos", sizeof(uint32_t) * (size_t)slots * (size_t)n, (void **)&pos));
    VGA_TRY(ctx->ws.get("ma_heap", sizeof(u64) * (size_t)slots * (size_t)n, (void **)&heap));
    DevBuf<int32_t> d_refs;
    DevBuf<uint8_t> d_expand;
    DevBuf<int32_t> d_partner;
    DevBuf<int64_t> d_src;
    DevBuf<float> d_out;
    DevBuf<unsigned long long> d_ctr;  // [0] next source, [1] unsafe angle evaluations
    VGA_TRY(d_refs.alloc((size_t)n));
    VGA_TRY(d_expand.alloc((size_t)n));
    VGA_TRY(d_src.alloc((size_t)nsrc));
    VGA_TRY(d_out.alloc((size_t)nsrc * 4));
    VGA_TRY(d_ctr.alloc_zero(2, st));
    VGA_CUDA(cudaMemcpyAsync(d_refs.p, g->h_refs.data(), sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    VGA_CUDA(cudaMemcpyAsync(d_expand.p, expand, (size_t)n, cudaMemcpyHostToDevice, st));
    if (partner) {
        for (int64_t v = 0; v < n; v++)
            if (partner[v] >= n || (partner[v] >= 0 && (partner[v] == v || partner[partner[v]] != v))) {
                set_error(std::string(who) + ": merge links must pair distinct cells symmetrically");
                return VGA_ERR_INVALID;
            }
        VGA_TRY(d_partner.alloc((size_t)n));
        VGA_CUDA(cudaMemcpyAsync(d_partner.p, partner, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    }
    MaArgs a{};
    a.n = n;
    a.runptr = g->f_runptr.p;
    a.runs = g->f_runs.p;
    a.refs = d_refs.p;
    a.expand = d_expand.p;
    a.partner = partner ? d_partner.p : nullptr;
    a.spacing = spacing;
    a.radius = radius;
    a.angular = angular;
    a.state = state;
    a.pos = pos;
    a.heap = heap;
    a.out = d_out.p;
    a.unsafe = d_ctr.p + 1;
    a.next_source = d_ctr.p;
    kt.start();
    mt.start();
    // several launches so that progress can be reported and a cancel request honoured between them
    const int64_t per_launch = std::max<int64_t>(slots * 8, 1);
    std::vector<float> h_out((size_t)nsrc * 4);
    for (int64_t b = 0; b < nsrc; b += per_launch) {
        const int64_t cnt = std::min(per_launch, nsrc - b);
        VGA_CUDA(cudaMemcpyAsync(d_src.p + b, sources + b, sizeof(int64_t) * cnt, cudaMemcpyHostToDevice, st));
        VGA_CUDA(cudaMemsetAsync(d_ctr.p, 0, sizeof(unsigned long long), st));
        a.sources = d_src.p + b;
        a.nsrc = cnt;
        a.out = d_out.p + 4 * b;
        const unsigned blocks = (unsigned)(std::min<int64_t>(slots, (cnt + MWARPS - 1) / MWARPS * MWARPS) / MWARPS);
        k_metric_angular<<<blocks, MWARPS * 32, 0, st>>>(a);
        tm.launches++;
        tm.main_launches++;
        VGA_CUDA(cudaMemcpyAsync(h_out.data() + 4 * b, d_out.p + 4 * b, sizeof(float) * 4 * cnt, cudaMemcpyDeviceToHost, st));
        VGA_CUDA(cudaStreamSynchronize(st));
        VGA_CUDA(cudaGetLastError());
        for (int c = 0; c < nout; c++)
            if (out[c]) std::copy(h_out.begin() + 4 * b + (int64_t)c * cnt, h_out.begin() + 4 * b + (int64_t)(c + 1) * cnt, out[c] + b);
        if (ctx->progress) ctx->progress(ctx->user, b + cnt, nsrc);
        if (ctx->cancel && ctx->cancel(ctx->user)) {
            set_error("cancelled");
            return VGA_ERR_CANCELLED;
        }
    }
    mt.stop();
    kt.stop();
    unsigned long long h_unsafe = 0;
    VGA_CUDA(cudaMemcpyAsync(&h_unsafe, d_ctr.p + 1, sizeof(h_unsafe), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    if (angle_unsafe) *angle_unsafe = (int64_t)h_unsafe;
    return VGA_OK;
}

}  // namespace vga
