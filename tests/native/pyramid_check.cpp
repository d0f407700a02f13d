// CPU build of the index logic of depthmapx_b200/csrc/pyramid.cuh (the functions the BFS kernels call), exposed to
// tests/test_pyramid_logic.py through a few C entry points.  Test infrastructure: compiled by the test into a temporary
// shared object, never shipped.
#include <cstdint>
#include <vector>

#include "../../depthmapx_b200/csrc/pyramid.cuh"

using vga::PyrLayout;
typedef unsigned long long u64;

namespace {
// pyramid of levels >= 1 built with pyr_build_group, three levels per pass, exactly as run_levels launches it
template <int W> void build(const u64 *fr, int64_t n, std::vector<u64> &pyr, PyrLayout &L) {
    L = vga::pyr_layout(n);
    pyr.assign((size_t)(L.total * W + 1), 0xdeadbeefdeadbeefULL);  // poison: every node must be written
    for (int k = 0; k + 1 < L.levels; k += 3) {
        const u64 *src = k == 0 ? fr : pyr.data() + L.off[k] * W;
        auto lvl = [&](int kk) { return kk < L.levels ? pyr.data() + L.off[kk] * W : (u64 *)nullptr; };
        auto cnt = [&](int kk) { return kk < L.levels ? L.cnt[kk] : (int64_t)0; };
        const int64_t groups = (L.cnt[k] + 7) / 8;
        for (int64_t t = 0; t < groups; t++)
            vga::pyr_build_group<W>(src, L.cnt[k], lvl(k + 1), cnt(k + 1), lvl(k + 2), cnt(k + 2), lvl(k + 3), cnt(k + 3), t);
    }
}

template <int W> int query(const u64 *fr, const std::vector<u64> &pyr, const PyrLayout &L, uint32_t a, uint32_t len, u64 *out) {
    for (int j = 0; j < W; j++) out[j] = 0;
    return vga::pyr_decompose(a, len, [&](int k, uint32_t i) {
        const u64 *p = k == 0 ? fr + (int64_t)i * W : pyr.data() + (L.off[k] + (int64_t)i) * W;
        for (int j = 0; j < W; j++) out[j] |= p[j];
    });
}
}  // namespace

namespace {
// range-OR updates through the pyramid + down pass (top chunk first), as the top-down step with run-length rows does it
template <int W> int updates(int64_t n, const uint32_t *a, const uint32_t *len, const u64 *word, int64_t nu, u64 *leaves) {
    PyrLayout L = vga::pyr_layout(n);
    std::vector<u64> pyr((size_t)(L.total * W + 1), 0ULL);
    for (int64_t q = 0; q < nu; q++)
        vga::pyr_decompose(a[q], len[q], [&](int k, uint32_t i) {
            u64 *p = k == 0 ? leaves + (int64_t)i * W : pyr.data() + (L.off[k] + (int64_t)i) * W;
            for (int j = 0; j < W; j++) p[j] |= word[q * W + j];
        });
    int kmax = 0;
    while (kmax + 3 + 1 < L.levels) kmax += 3;  // largest multiple of 3 with a level above it
    for (int k = kmax; k >= 0; k -= 3) {
        if (k + 1 >= L.levels) continue;
        u64 *dst = k == 0 ? leaves : pyr.data() + L.off[k] * W;
        auto lvl = [&](int kk) { return kk < L.levels ? pyr.data() + L.off[kk] * W : (u64 *)nullptr; };
        auto cnt = [&](int kk) { return kk < L.levels ? L.cnt[kk] : (int64_t)0; };
        const int64_t groups = (L.cnt[k] + 7) / 8;
        for (int64_t t = 0; t < groups; t++)
            vga::pyr_down_group<W>(dst, L.cnt[k], lvl(k + 1), cnt(k + 1), lvl(k + 2), cnt(k + 2), lvl(k + 3), cnt(k + 3), t);
    }
    for (int64_t i = 0; i < L.total * W; i++)
        if (pyr[(size_t)i]) return -1;  // every node must have been cleared by the down pass
    return 0;
}
}  // namespace

extern "C" {

// leaves[n*w] |= word[q] over [a[q], a[q]+len[q]) for every update q, done through pyramid nodes + the down pass
int pyrchk_updates(int w, int64_t n, const uint32_t *a, const uint32_t *len, const u64 *word, int64_t nu, u64 *leaves) {
    return w == 1 ? updates<1>(n, a, len, word, nu, leaves) : w == 2 ? updates<2>(n, a, len, word, nu, leaves)
                                                                     : updates<4>(n, a, len, word, nu, leaves);
}

int pyrchk_layout(int64_t n, int64_t *cnt, int64_t *off, int64_t *total) {
    PyrLayout L = vga::pyr_layout(n);
    for (int k = 0; k < vga::PYR_MAX_LEVELS; k++) {
        cnt[k] = L.cnt[k];
        off[k] = L.off[k];
    }
    *total = L.total;
    return L.levels;
}

// runs queries [a[i], a[i]+len[i]) over a pyramid built from fr (n nodes of w words); out = w words per query,
// loads = nodes visited per query; returns -1 if a node index left its level, a poisoned node was read, or the node
// count differs from pyr_cost
int pyrchk_queries(int w, const u64 *fr, int64_t n, const uint32_t *a, const uint32_t *len, int64_t nq, u64 *out, int32_t *loads) {
    std::vector<u64> pyr;
    PyrLayout L;
    if (w == 1) build<1>(fr, n, pyr, L); else if (w == 2) build<2>(fr, n, pyr, L); else build<4>(fr, n, pyr, L);
    for (int64_t i = 0; i < L.total * w; i++)
        if (pyr[(size_t)i] == 0xdeadbeefdeadbeefULL) return -2;  // a node was never written
    for (int64_t q = 0; q < nq; q++) {
        bool bad = false;
        vga::pyr_decompose(a[q], len[q], [&](int k, uint32_t i) {
            if (k >= L.levels || (int64_t)i >= L.cnt[k]) bad = true;
        });
        if (bad) return -1;
        int nodes = w == 1 ? query<1>(fr, pyr, L, a[q], len[q], out + q * w)
                  : w == 2 ? query<2>(fr, pyr, L, a[q], len[q], out + q * w)
                           : query<4>(fr, pyr, L, a[q], len[q], out + q * w);
        if (nodes != vga::pyr_cost(a[q], len[q])) return -1;
        loads[q] = nodes;
    }
    return 0;
}
}
