// OR-pyramid over the frontier words of one BFS batch, and range-OR queries over it (bfs_pull = 1).
//
// Rows of a grid visibility graph are unions of a few runs of consecutive ordinals (a cell sees contiguous
// vertical spans of cells, and ordinals number the filled cells x-major): office plans ~9, urban plans ~35
// entries per run.  The bottom-up step asks, for a vertex w, for OR_{u in inrow(w)} frontier[u]; with the in-row
// kept as runs [a, a+len) that is a handful of range-OR queries, and a pyramid
//      P_0 = frontier,  P_k[i] = P_{k-1}[2i] | P_{k-1}[2i+1]      (n_k = ceil(n_{k-1} / 2))
// answers each with at most 2*log2(len) loads instead of len (5x fewer loads for a full scan of an office plan,
// 16x for an urban plan; measured with the oracle's adjacency).
//
// This header is plain C++ apart from the VGA_HD annotation so that the index logic (level layout, query
// decomposition, run cost) is unit-tested on the CPU as compiled code (tests/native/pyramid_check.cpp,
// tests/test_pyramid_logic.py) -- the same functions the kernels in bfs.cu call.
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define VGA_HD __host__ __device__ __forceinline__
#else
#define VGA_HD inline
#endif

namespace vga {

constexpr int PYR_MAX_LEVELS = 34;  // enough for n < 2^32

// Level layout for n vertices: level k >= 1 has cnt[k] = ceil(cnt[k-1] / 2) nodes (cnt[0] = n) stored at word offset
// off[k] of the pyramid array (level 0 is the frontier array itself and is not stored).  levels = number of levels
// incl. level 0; the top level has one node.
struct PyrLayout {
    int levels;
    int64_t cnt[PYR_MAX_LEVELS];
    int64_t off[PYR_MAX_LEVELS];
    int64_t total;  // nodes of levels >= 1
};

inline PyrLayout pyr_layout(int64_t n) {
    PyrLayout L;
    L.levels = 1;
    L.cnt[0] = n;
    L.off[0] = 0;
    L.total = 0;
    while (L.cnt[L.levels - 1] > 1 && L.levels < PYR_MAX_LEVELS) {
        const int k = L.levels;
        L.cnt[k] = (L.cnt[k - 1] + 1) / 2;
        L.off[k] = L.total;
        L.total += L.cnt[k];
        L.levels++;
    }
    for (int k = L.levels; k < PYR_MAX_LEVELS; k++) {
        L.cnt[k] = 0;
        L.off[k] = L.total;
    }
    return L;
}

// Decomposes [a, a+len) into pyramid nodes and calls visit(level, index) for each: node (k, i) covers
// [i*2^k, (i+1)*2^k) clipped to n.  At most 2 nodes per level.  Returns the number of nodes visited.
template <typename Visit> VGA_HD int pyr_decompose(uint32_t a, uint32_t len, Visit &&visit) {
    uint32_t l = a, r = a + len;
    int k = 0, nodes = 0;
    while (l < r) {
        if (l & 1u) {
            visit(k, l);
            l++;
            nodes++;
        }
        if (l < r && (r & 1u)) {
            r--;
            visit(k, r);
            nodes++;
        }
        l >>= 1;
        r >>= 1;
        k++;
    }
    return nodes;
}

// W consecutive words of a node: 16-byte accesses in device code (nodes are 8 * W bytes apart and 16-byte aligned for W >= 2),
// plain words on the host (unit tests, emulation)
template <int W> VGA_HD void pyr_ld(const unsigned long long *p, unsigned long long (&o)[W]) {
#if defined(__CUDA_ARCH__)
    if constexpr (W >= 2) {
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            const ulonglong2 t = *reinterpret_cast<const ulonglong2 *>(p + j);
            o[j] = t.x;
            o[j + 1] = t.y;
        }
        return;
    }
#endif
#pragma unroll
    for (int j = 0; j < W; j++) o[j] = p[j];
}
template <int W> VGA_HD void pyr_st(unsigned long long *p, const unsigned long long (&o)[W]) {
#if defined(__CUDA_ARCH__)
    if constexpr (W >= 2) {
#pragma unroll
        for (int j = 0; j < W; j += 2) {
            ulonglong2 t;
            t.x = o[j];
            t.y = o[j + 1];
            *reinterpret_cast<ulonglong2 *>(p + j) = t;
        }
        return;
    }
#endif
#pragma unroll
    for (int j = 0; j < W; j++) p[j] = o[j];
}
template <int W> VGA_HD bool pyr_any(const unsigned long long (&o)[W]) {
    unsigned long long a = 0ULL;
#pragma unroll
    for (int j = 0; j < W; j++) a |= o[j];
    return a != 0ULL;
}

// Builds three pyramid levels at once: work item t ORs the aligned group of 8 nodes [8t, 8t+8) of level k (src, cnt0
// nodes of W words each) into 4 nodes of level k+1, 2 of level k+2 and 1 of level k+3 (d1/d2/d3 with cnt1/cnt2/cnt3
// nodes; a count of 0 = that level does not exist).  Work items: ceil(cnt0 / 8).
// `leaf` (optional, level 0 only): the words of level-0 node idx live at src[leaf[idx]] -- a pyramid over a PERMUTED
// order of the vertices (the y-major pyramid of bfs.cu) reads / writes the shared x-major state through it.
template <int W>
VGA_HD void pyr_build_group(const unsigned long long *src, int64_t cnt0, unsigned long long *d1, int64_t cnt1,
                            unsigned long long *d2, int64_t cnt2, unsigned long long *d3, int64_t cnt3, int64_t t,
                            const uint32_t *leaf = nullptr) {
    unsigned long long a[8][W];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t idx = 8 * t + i;
        const int64_t at = (leaf && idx < cnt0) ? (int64_t)leaf[idx] : idx;
        if (idx < cnt0) {
            pyr_ld<W>(src + at * W, a[i]);
        } else {
#pragma unroll
            for (int j = 0; j < W; j++) a[i][j] = 0ULL;
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < W; j++) a[i][j] = a[2 * i][j] | a[2 * i + 1][j];
        if (4 * t + i < cnt1) pyr_st<W>(d1 + (4 * t + i) * W, a[i]);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < W; j++) a[i][j] = a[2 * i][j] | a[2 * i + 1][j];
        if (2 * t + i < cnt2) pyr_st<W>(d2 + (2 * t + i) * W, a[i]);
    }
    if (t < cnt3) {
#pragma unroll
        for (int j = 0; j < W; j++) a[0][j] |= a[1][j];
        pyr_st<W>(d3 + t * W, a[0]);
    }
}

// The reverse direction, for range-OR UPDATES (top-down step with run-length rows): words are ORed into the nodes of
// pyr_decompose(a, len), and a down pass then pushes every node's word to the leaves it covers.  Work item t takes the
// word of node t of level k+3 (already complete: the pass runs from the top), ORs in the nodes of levels k+2 and k+1
// below it and adds the result to the 8 nodes [8t, 8t+8) of level k (dst); the nodes it read are cleared for the next
// round.  Same grouping as pyr_build_group; a count of 0 = that level does not exist.  Returns the mask of level-k nodes written.
template <int W>
VGA_HD unsigned pyr_down_group(unsigned long long *dst, int64_t cnt0, unsigned long long *s1, int64_t cnt1, unsigned long long *s2,
                               int64_t cnt2, unsigned long long *s3, int64_t cnt3, int64_t t, const uint32_t *leaf = nullptr) {
    unsigned long long a3[W], a2[2][W], a1[4][W], zero[W];
#pragma unroll
    for (int j = 0; j < W; j++) a3[j] = zero[j] = 0ULL;
    if (t < cnt3) {
        pyr_ld<W>(s3 + t * W, a3);
        if (pyr_any<W>(a3)) pyr_st<W>(s3 + t * W, zero);
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < W; j++) a2[i][j] = a3[j];
        if (2 * t + i < cnt2) {
            unsigned long long v[W];
            pyr_ld<W>(s2 + (2 * t + i) * W, v);
            if (pyr_any<W>(v)) pyr_st<W>(s2 + (2 * t + i) * W, zero);
#pragma unroll
            for (int j = 0; j < W; j++) a2[i][j] |= v[j];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < W; j++) a1[i][j] = a2[i >> 1][j];
        if (4 * t + i < cnt1) {
            unsigned long long v[W];
            pyr_ld<W>(s1 + (4 * t + i) * W, v);
            if (pyr_any<W>(v)) pyr_st<W>(s1 + (4 * t + i) * W, zero);
#pragma unroll
            for (int j = 0; j < W; j++) a1[i][j] |= v[j];
        }
    }
    unsigned written = 0u;  // bit i: node 8t+i of level k received something
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t idx = 8 * t + i;
        if (idx < cnt0 && pyr_any<W>(a1[i >> 1])) {
            const int64_t at = leaf ? (int64_t)leaf[idx] : idx;
            unsigned long long cur[W];
            pyr_ld<W>(dst + at * W, cur);
#pragma unroll
            for (int j = 0; j < W; j++) cur[j] |= a1[i >> 1][j];
            pyr_st<W>(dst + at * W, cur);
            written |= 1u << i;
        }
    }
    return written;
}

// number of pyramid loads a query of [a, a+len) costs (the pull step's work estimate)
VGA_HD int pyr_cost(uint32_t a, uint32_t len) {
    return pyr_decompose(a, len, [](int, uint32_t) {});
}

}  // namespace vga
