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
#pragma unroll
        for (int j = 0; j < W; j++) a[i][j] = idx < cnt0 ? src[at * W + j] : 0ULL;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int j = 0; j < W; j++) a[i][j] = a[2 * i][j] | a[2 * i + 1][j];
        if (4 * t + i < cnt1) {
#pragma unroll
            for (int j = 0; j < W; j++) d1[(4 * t + i) * W + j] = a[i][j];
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
#pragma unroll
        for (int j = 0; j < W; j++) a[i][j] = a[2 * i][j] | a[2 * i + 1][j];
        if (2 * t + i < cnt2) {
#pragma unroll
            for (int j = 0; j < W; j++) d2[(2 * t + i) * W + j] = a[i][j];
        }
    }
    if (t < cnt3) {
#pragma unroll
        for (int j = 0; j < W; j++) d3[t * W + j] = a[0][j] | a[1][j];
    }
}

// The reverse direction, for range-OR UPDATES (top-down step with run-length rows): words are ORed into the nodes of
// pyr_decompose(a, len), and a down pass then pushes every node's word to the leaves it covers.  Work item t takes the
// word of node t of level k+3 (already complete: the pass runs from the top), ORs in the nodes of levels k+2 and k+1
// below it and adds the result to the 8 nodes [8t, 8t+8) of level k (dst); the nodes it read are cleared for the next
// round.  Same grouping as pyr_build_group; a count of 0 = that level does not exist.
template <int W>
VGA_HD void pyr_down_group(unsigned long long *dst, int64_t cnt0, unsigned long long *s1, int64_t cnt1, unsigned long long *s2,
                           int64_t cnt2, unsigned long long *s3, int64_t cnt3, int64_t t, const uint32_t *leaf = nullptr) {
    unsigned long long a3[W], a2[2][W], a1[4][W];
#pragma unroll
    for (int j = 0; j < W; j++) {
        a3[j] = 0ULL;
        if (t < cnt3) {
            a3[j] = s3[t * W + j];
            if (a3[j]) s3[t * W + j] = 0ULL;
        }
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < W; j++) {
            a2[i][j] = a3[j];
            if (2 * t + i < cnt2) {
                const unsigned long long v = s2[(2 * t + i) * W + j];
                if (v) s2[(2 * t + i) * W + j] = 0ULL;
                a2[i][j] |= v;
            }
        }
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < W; j++) {
            a1[i][j] = a2[i >> 1][j];
            if (4 * t + i < cnt1) {
                const unsigned long long v = s1[(4 * t + i) * W + j];
                if (v) s1[(4 * t + i) * W + j] = 0ULL;
                a1[i][j] |= v;
            }
        }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int64_t idx = 8 * t + i;
        if (idx < cnt0) {
            const int64_t at = leaf ? (int64_t)leaf[idx] : idx;
#pragma unroll
            for (int j = 0; j < W; j++)
                if (a1[i >> 1][j]) dst[at * W + j] |= a1[i >> 1][j];
        }
    }
}

// number of pyramid loads a query of [a, a+len) costs (the pull step's work estimate)
VGA_HD int pyr_cost(uint32_t a, uint32_t len) {
    return pyr_decompose(a, len, [](int, uint32_t) {});
}

}  // namespace vga
