// TEST INFRASTRUCTURE ONLY: a small SIMT emulator that lets the CUDA translation units of libvga_b200 run on a CPU, so
// that kernel logic (index arithmetic, warp masks, shuffles, ballots, reductions, shared memory, grid-stride loops) and
// the host-side launch sequences can be exercised by `pytest -m "not gpu"` in a container without a GPU.
//
// Never shipped, never loaded by the product: tests/emu/build_emu.py preprocesses csrc/*.cu (kernel launches become
// simt::launch calls), compiles them with g++ against this header (which stands in for <cuda_runtime.h> and
// <cub/cub.cuh>) into a scratch directory, and the tests point their ctypes loader at that directory.
//
// Model: blocks run one after the other; the threads of a block are fibers (ucontext) that run until they finish or
// reach a warp/block collective (__shfl_sync, __ballot_sync, __reduce_*_sync, __any_sync, __syncwarp, __syncthreads),
// where they wait until every lane named in the mask (resp. every live thread of the block) has arrived with the same
// operation.  Lanes that have exited count as arrived.  A collective that can never complete (mismatched masks,
// divergent __syncthreads) is reported as a deadlock and aborts.  Everything is single-threaded, so data races are
// NOT detected; memory is the host heap ("device" pointers are host pointers).
#pragma once

#if defined(__x86_64__) && !defined(SIMT_USE_UCONTEXT)
#define SIMT_ASM_SWITCH 1
#else
#include <ucontext.h>
#endif

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define VGA_SIMT_EMULATION 1

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint2 {
    unsigned x, y;
};
inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
struct alignas(16) uint4 {
    unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
struct alignas(16) ulonglong2 {
    unsigned long long x, y;
};
inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { return ulonglong2{a, b}; }

// ---------------------------------------------------------------------------------------------- runtime API
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
typedef int cudaStream_t_tag;
typedef cudaStream_t_tag *cudaStream_t;
struct cudaEvent_rec {
    std::chrono::steady_clock::time_point t;
};
typedef cudaEvent_rec *cudaEvent_t;
typedef int cudaMemPool_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1 };
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize };
struct cudaDeviceProp {
    int multiProcessorCount;
    size_t sharedMemPerBlockOptin;
};

inline const char *cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) {
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    p->multiProcessorCount = 4;
    p->sharedMemPerBlockOptin = 227 * 1024;
    return cudaSuccess;
}
inline cudaError_t cudaMalloc(void **p, size_t n) {
    *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256);
    return *p ? cudaSuccess : cudaErrorUnknown;
}
inline cudaError_t cudaFree(void *p) {
    free(p);
    return cudaSuccess;
}
inline cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
inline cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
    if (n) memmove(d, s, n);
    return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t) {
    if (n) memset(d, v, n);
    return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) {
    *s = new cudaStream_t_tag(0);
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
    delete s;
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) {
    *e = new cudaEvent_rec();
    return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = std::chrono::steady_clock::now();
    return cudaSuccess;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
    return cudaSuccess;
}
inline cudaError_t cudaMemGetInfo(size_t *f, size_t *t) {
    *f = (size_t)2 << 30;
    *t = (size_t)4 << 30;
    return cudaSuccess;
}
inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t *p, int) {
    *p = 0;
    return cudaSuccess;
}
inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void *) { return cudaSuccess; }
inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
template <typename F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }

// ---------------------------------------------------------------------------------------------- SIMT core
namespace simt {

enum Op { OP_NONE, OP_BALLOT, OP_ANY, OP_SHFL, OP_SHFL_DOWN, OP_SHFL_UP, OP_SHFL_XOR, OP_RED_OR, OP_RED_ADD, OP_RED_MIN, OP_RED_MAX, OP_SYNCWARP, OP_SYNCTHREADS };

// Fiber switch.  x86-64: a hand-written switch of the callee-saved registers and the stack pointer (glibc's swapcontext
// makes a signal-mask system call per switch, which dominated the run time); elsewhere: ucontext.
#if defined(SIMT_ASM_SWITCH)
typedef void *Context;  // saved stack pointer
extern "C" void simt_switch(Context *save, Context load);
asm(R"(
.text
.weak simt_switch
.type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size simt_switch,.-simt_switch
)");
#else
typedef ucontext_t Context;
#endif

struct Lane {
    Context ctx;
    bool done = false, waiting = false;
    Op op = OP_NONE;
    unsigned mask = 0;
    uint64_t val = 0, result = 0;
    int arg = 0;
    uint3 tid{0, 0, 0};
};

struct Runtime {
    Context sched;
    std::vector<Lane> lanes;
    std::vector<char> stacks;
    Lane *cur = nullptr;
    uint3 bidx{0, 0, 0};
    dim3 bdim, gdim;
    const std::function<void()> *body = nullptr;
    std::vector<char> dyn;
    long launches = 0;
};
inline Runtime R;
constexpr size_t STACK = 256 * 1024;

inline void *dyn_smem() { return R.dyn.data(); }

inline void to_scheduler(Lane *me) {
#if defined(SIMT_ASM_SWITCH)
    simt_switch(&me->ctx, R.sched);
#else
    swapcontext(&me->ctx, &R.sched);
#endif
}
inline void to_lane(Lane *l) {
    R.cur = l;
#if defined(SIMT_ASM_SWITCH)
    simt_switch(&R.sched, l->ctx);
#else
    swapcontext(&R.sched, &l->ctx);
#endif
}

inline void trampoline() {
    (*R.body)();
    R.cur->done = true;
    R.cur->waiting = false;
#if defined(SIMT_ASM_SWITCH)
    to_scheduler(R.cur);  // a finished fiber is never resumed
    abort();
#endif
}

inline void make_lane_context(Lane &l, char *stack, size_t size) {
#if defined(SIMT_ASM_SWITCH)
    // initial frame: six callee-saved registers, the entry point for `ret`, and a null return address for the entry
    uintptr_t top = ((uintptr_t)stack + size) & ~(uintptr_t)15;
    void **sp = (void **)top;
    *--sp = nullptr;
    *--sp = (void *)trampoline;
    for (int i = 0; i < 6; i++) *--sp = nullptr;
    l.ctx = (Context)sp;
#else
    getcontext(&l.ctx);
    l.ctx.uc_stack.ss_sp = stack;
    l.ctx.uc_stack.ss_size = size;
    l.ctx.uc_link = &R.sched;
    makecontext(&l.ctx, (void (*)())trampoline, 0);
#endif
}

inline void resolve_warp(Lane *w, int count, bool &progress) {
    bool handled[32] = {false};
    for (int i = 0; i < count; i++) {
        Lane &a = w[i];
        if (a.done || !a.waiting || handled[i] || a.op == OP_SYNCTHREADS) continue;
        // participants: the lanes named in the mask that exist and have not exited
        bool ready = true;
        for (int l = 0; l < 32 && ready; l++) {
            if (!((a.mask >> l) & 1u) || l >= count || w[l].done) continue;
            if (!(w[l].waiting && w[l].op == a.op && w[l].mask == a.mask)) ready = false;
        }
        if (!((a.mask >> i) & 1u)) {
            fprintf(stderr, "simt: lane %d calls a collective with mask %08x that does not name it\n", i, a.mask);
            abort();
        }
        if (!ready) continue;
        auto in = [&](int l) { return l >= 0 && l < count && ((a.mask >> l) & 1u) && !w[l].done; };
        uint64_t red = a.op == OP_RED_MIN ? ~0ull : 0;
        for (int l = 0; l < 32; l++)
            if (in(l)) {
                if (a.op == OP_RED_MIN && w[l].val < red) red = w[l].val;
                if (a.op == OP_RED_MAX && w[l].val > red) red = w[l].val;
                if (a.op == OP_BALLOT || a.op == OP_ANY) red |= (w[l].val ? 1ull : 0ull) << l;
                if (a.op == OP_RED_OR) red |= w[l].val;
                if (a.op == OP_RED_ADD) red += w[l].val;
            }
        uint64_t res[32];
        for (int l = 0; l < 32; l++) {
            if (!in(l)) continue;
            switch (a.op) {
            case OP_BALLOT: res[l] = red; break;
            case OP_ANY: res[l] = red != 0; break;
            case OP_RED_OR: case OP_RED_ADD: case OP_RED_MIN: case OP_RED_MAX: res[l] = red; break;
            case OP_SHFL: { int s = w[l].arg & 31; res[l] = in(s) ? w[s].val : w[l].val; break; }
            case OP_SHFL_DOWN: { int s = l + w[l].arg; res[l] = (s < 32 && in(s)) ? w[s].val : w[l].val; break; }
            case OP_SHFL_UP: { int s = l - w[l].arg; res[l] = (s >= 0 && in(s)) ? w[s].val : w[l].val; break; }
            case OP_SHFL_XOR: { int s = (l ^ w[l].arg) & 31; res[l] = in(s) ? w[s].val : w[l].val; break; }
            default: res[l] = 0;
            }
        }
        for (int l = 0; l < 32; l++)
            if (in(l)) {
                w[l].result = res[l];
                w[l].waiting = false;
                handled[l] = true;
            }
        progress = true;
    }
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads <= 0 || nthreads > 1024) {
        fprintf(stderr, "simt: bad block size %d\n", nthreads);
        abort();
    }
    R.launches++;
    R.bdim = block;
    R.gdim = grid;
    R.body = &body;
    R.dyn.assign(smem + 16, 0);
    if (R.stacks.size() < (size_t)nthreads * STACK) R.stacks.resize((size_t)nthreads * STACK);
    R.lanes.resize((size_t)nthreads);
    for (unsigned bz = 0; bz < grid.z; bz++)
        for (unsigned by = 0; by < grid.y; by++)
            for (unsigned bx = 0; bx < grid.x; bx++) {
                R.bidx = uint3{bx, by, bz};
                for (int t = 0; t < nthreads; t++) {
                    Lane &l = R.lanes[(size_t)t];
                    l.done = false;
                    l.waiting = false;
                    l.op = OP_NONE;
                    l.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
                    make_lane_context(l, R.stacks.data() + (size_t)t * STACK, STACK);
                }
                for (;;) {
                    bool progress = false, alldone = true;
                    for (int t = 0; t < nthreads; t++) {
                        Lane &l = R.lanes[(size_t)t];
                        if (l.done) continue;
                        alldone = false;
                        if (l.waiting) continue;
                        to_lane(&l);
                        progress = true;
                    }
                    if (alldone) break;
                    for (int w0 = 0; w0 < nthreads; w0 += 32) resolve_warp(&R.lanes[(size_t)w0], std::min(32, nthreads - w0), progress);
                    // __syncthreads: every live thread of the block must be waiting on it
                    bool bar = true, any_live = false;
                    for (int t = 0; t < nthreads; t++) {
                        Lane &l = R.lanes[(size_t)t];
                        if (l.done) continue;
                        any_live = true;
                        if (!(l.waiting && l.op == OP_SYNCTHREADS)) bar = false;
                    }
                    if (any_live && bar) {
                        for (int t = 0; t < nthreads; t++) R.lanes[(size_t)t].waiting = false;
                        progress = true;
                    }
                    if (!progress) {
                        fprintf(stderr, "simt: deadlock in block (%u,%u,%u): ", bx, by, bz);
                        for (int t = 0; t < nthreads && t < 64; t++) {
                            Lane &l = R.lanes[(size_t)t];
                            fprintf(stderr, "[%d:%s op%d m%08x] ", t, l.done ? "done" : l.waiting ? "wait" : "run", (int)l.op, l.mask);
                        }
                        fprintf(stderr, "\n");
                        abort();
                    }
                }
            }
    R.body = nullptr;
    R.cur = nullptr;
}

inline uint64_t collective(Op op, unsigned mask, uint64_t val, int arg) {
    Lane *me = R.cur;
    me->op = op;
    me->mask = mask;
    me->val = val;
    me->arg = arg;
    me->waiting = true;
    to_scheduler(me);
    return me->result;
}

template <typename T> inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "shuffle of a type wider than 64 bits");
    uint64_t b = 0;
    memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T> inline T from_bits(uint64_t b) {
    T v;
    memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace simt

// Work counters: VGA_COUNT(name, n) in kernel code (a no-op in the CUDA build, csrc/vga_dev.cuh) adds n to a named counter
// here, so that an emulated run reports how many adjacency entries / pyramid nodes / atomics a schedule really touches.
namespace simt {
inline std::map<std::string, long long> &counters() {
    static std::map<std::string, long long> c;
    return c;
}
}  // namespace simt
extern "C" __attribute__((used, weak)) const char *simt_counters_dump(int reset) {
    static std::string out;
    out = "{";
    for (auto &kv : simt::counters()) out += (out.size() > 1 ? ", \"" : "\"") + kv.first + "\": " + std::to_string(kv.second);
    out += "}";
    if (reset)
        for (auto &kv : simt::counters()) kv.second = 0;  // slots stay: call sites hold references to them
    return out.c_str();
}

#define threadIdx (simt::R.cur->tid)
#define blockIdx (simt::R.bidx)
#define blockDim (simt::R.bdim)
#define gridDim (simt::R.gdim)

inline void __syncthreads() { simt::collective(simt::OP_SYNCTHREADS, 0xffffffffu, 0, 0); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { simt::collective(simt::OP_SYNCWARP, mask, 0, 0); }
inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned)simt::collective(simt::OP_BALLOT, mask, pred != 0, 0) & mask; }
inline int __any_sync(unsigned mask, int pred) { return (int)simt::collective(simt::OP_ANY, mask, pred != 0, 0); }
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int = 32) {
    return simt::from_bits<T>(simt::collective(simt::OP_SHFL, mask, simt::to_bits(v), src));
}
template <typename T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int = 32) {
    return simt::from_bits<T>(simt::collective(simt::OP_SHFL_DOWN, mask, simt::to_bits(v), (int)delta));
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int = 32) {
    return simt::from_bits<T>(simt::collective(simt::OP_SHFL_UP, mask, simt::to_bits(v), (int)delta));
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int = 32) {
    return simt::from_bits<T>(simt::collective(simt::OP_SHFL_XOR, mask, simt::to_bits(v), lanemask));
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned v) { return (unsigned)simt::collective(simt::OP_RED_OR, mask, v, 0); }
inline unsigned __reduce_add_sync(unsigned mask, unsigned v) { return (unsigned)simt::collective(simt::OP_RED_ADD, mask, v, 0); }
inline unsigned __reduce_min_sync(unsigned mask, unsigned v) { return (unsigned)simt::collective(simt::OP_RED_MIN, mask, v, 0); }
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return (unsigned)simt::collective(simt::OP_RED_MAX, mask, v, 0); }
inline int __reduce_add_sync(unsigned mask, int v) { return (int)(unsigned)simt::collective(simt::OP_RED_ADD, mask, (unsigned)v, 0); }

// position of the offset-th set bit of mask counting from bit `base` upwards (offset > 0), 0xffffffff if there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset > 0) {
        for (unsigned b = base; b < 32; b++)
            if ((mask >> b) & 1u)
                if (--offset == 0) return b;
    } else if (offset < 0) {
        for (int b = (int)base; b >= 0; b--)
            if ((mask >> b) & 1u)
                if (++offset == 0) return (unsigned)b;
    } else if ((mask >> base) & 1u) {
        return base;
    }
    return 0xffffffffu;
}
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

// arithmetic intrinsics with explicit rounding: plain IEEE operations (compile with -ffp-contract=off)
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return std::sqrt(a); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __double2float_rn(double a) { return (float)a; }
inline unsigned __float_as_uint(float v) { unsigned u; std::memcpy(&u, &v, 4); return u; }
inline float __uint_as_float(unsigned u) { float v; std::memcpy(&v, &u, 4); return v; }
inline int __float_as_int(float v) { int u; std::memcpy(&u, &v, 4); return u; }
inline float __int_as_float(int u) { float v; std::memcpy(&v, &u, 4); return v; }
inline double __longlong_as_double(long long v) { return simt::from_bits<double>((uint64_t)v); }
inline long long __double_as_longlong(double v) { return (long long)simt::to_bits(v); }

// cache-hinted loads / stores: plain accesses on the CPU
template <typename T> inline T __ldcs(const T *p) { return *p; }
template <typename T> inline T __ldcg(const T *p) { return *p; }
template <typename T> inline void __stcs(T *p, T v) { *p = v; }
template <typename T, typename V> inline T atomicAdd(T *p, V v) {
    T o = *p;
    *p = (T)(o + (T)v);
    return o;
}
template <typename T, typename V> inline T atomicOr(T *p, V v) {
    T o = *p;
    *p = (T)(o | (T)v);
    return o;
}
template <typename T, typename V> inline T atomicExch(T *p, V v) {
    T o = *p;
    *p = (T)v;
    return o;
}
template <typename T, typename V> inline T atomicMax(T *p, V v) {
    T o = *p;
    if ((T)v > o) *p = (T)v;
    return o;
}
template <typename T, typename V> inline T atomicMin(T *p, V v) {
    T o = *p;
    if ((T)v < o) *p = (T)v;
    return o;
}

// CUDA's device-side min / max overloads
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
inline double min(double a, double b) { return a < b ? a : b; }
inline double max(double a, double b) { return a > b ? a : b; }
inline float min(float a, float b) { return a < b ? a : b; }
inline float max(float a, float b) { return a > b ? a : b; }

// ---------------------------------------------------------------------------------------------- cub stand-ins
namespace cub {
struct DeviceScan {
    template <typename In, typename Out>
    static cudaError_t ExclusiveSum(void *tmp, size_t &bytes, In in, Out out, int n, cudaStream_t = nullptr) {
        if (!tmp) {
            bytes = 16;
            return cudaSuccess;
        }
        auto sum = decltype(out[0] + out[0])(0);
        for (int i = 0; i < n; i++) {
            auto v = in[i];
            out[i] = sum;
            sum += v;
        }
        return cudaSuccess;
    }
};
struct DeviceSegmentedSort {
    template <typename K, typename OffA, typename OffB>
    static cudaError_t SortKeys(void *tmp, size_t &bytes, const K *in, K *out, int num_items, int num_segments, OffA begin, OffB end,
                                cudaStream_t = nullptr) {
        if (!tmp) {
            bytes = 16;
            return cudaSuccess;
        }
        if (num_items > 0 && in != out) memmove(out, in, sizeof(K) * (size_t)num_items);
        for (int s = 0; s < num_segments; s++) std::sort(out + begin[s], out + end[s]);
        return cudaSuccess;
    }
};
struct DeviceRadixSort {
    // stable ascending sort of (key, value) pairs on the key bits [begin_bit, end_bit)
    template <typename K, typename V>
    static cudaError_t SortPairs(void *tmp, size_t &bytes, const K *kin, K *kout, const V *vin, V *vout, int n, int begin_bit = 0,
                                 int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
        if (!tmp) {
            bytes = 16;
            return cudaSuccess;
        }
        std::vector<int> idx((size_t)n);
        for (int i = 0; i < n; i++) idx[(size_t)i] = i;
        const K mask = end_bit - begin_bit >= (int)sizeof(K) * 8 ? ~K(0) : (K)(((K(1) << (end_bit - begin_bit)) - 1));
        std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return ((kin[a] >> begin_bit) & mask) < ((kin[b] >> begin_bit) & mask); });
        for (int i = 0; i < n; i++) {
            kout[i] = kin[idx[(size_t)i]];
            vout[i] = vin[idx[(size_t)i]];
        }
        return cudaSuccess;
    }
};
}  // namespace cub
