// makegraph: grid visibility-graph construction on sm_100a.
//
// Replaces PointMap::sparkGraph2 / sparkPixel2 / sieve2 (salalib/pointdata.cpp:1246-1565) and
// sparkSieve2 (salalib/sparksieve2.cpp:33-173).  The algorithm is NOT a port of the reference's
// per-pixel std::list / std::vector / std::set code; it is re-derived for a warp:
//
//   * one warp owns one (source cell, octant) task; the 8 octants of a source sit in 8
//     consecutive warps so their cell reads share L1/L2 lines;
//   * the angular gap list and the row's wall "blocks" live in shared memory (per-warp slices);
//     rare tasks that overflow the slices are re-run by a second launch with global scratch;
//   * a depth row is swept by all 32 lanes at once: the union of the gaps' widened index ranges
//     is enumerated with a warp scan, every lane tests one candidate cell (centre-gap rule, FILLED,
//     axis/diagonal de-duplication, FP64 line test against the cell's own wall segments), accepted
//     cells are emitted in ascending index order with a ballot/popc prefix -- this equals the
//     reference's (gap, ind) visiting order because gaps are disjoint and sorted;
//   * blocks are sorted (start asc, end desc) by a warp bitonic network and subtracted from the
//     gap list by the reference's sequential rule (erase-if <= start+1e-10, split, "stay on this
//     block"), which is order dependent and therefore executed by one lane;
//   * all geometry is IEEE FP64 with explicit round-to-nearest mul/add (no FMA contraction), and
//     x/0 = +-inf is kept as the reference relies on it (sparksieve2.cpp:143-173).
//
// Two passes (count, emit) size the edge buffer exactly; rows are then finalised:
//   node_stats  -- Connectivity, the order-dependent double sums of the two moments, far bin
//                  distances, per-bin counts, grid connections (pointdata.cpp:1463-1497, 1735-1768)
//   make_keys + cub segmented sort -- rows sorted by x-major ordinal, packed col<<6|accepted<<5|bin.
#include <cub/cub.cuh>

#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <memory>

#include "vga_dev.cuh"

namespace vga {

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int WARPS_PER_BLOCK = 8;

struct Zone {
    double s, e;
};

struct Ln {
    double blx, bly, trx, try_;
    int parity;
};

struct GridDev {
    int cols, rows;
    double spacing, blx, bly, maxdist;
    const uint8_t *cflag;
    const uint8_t *cflag_t;
    const uint32_t *line_off;
    const double *lines;
    const int32_t *cellord;
    const int32_t *cellref;
    int64_t n;
};

// ---- FP64 helpers: explicit rn ops so that no fused multiply-add can ever be formed ------------
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ Ln load_line(const double *p) {
    Ln l;
    l.blx = p[0];
    l.bly = p[1];
    l.trx = p[2];
    l.try_ = p[3];
    l.parity = p[4] != 0.0;
    return l;
}

// normal form of the segment a->b (genlib/p2dpoly.cpp:291-336 semantics)
__device__ __forceinline__ Ln make_line(double ax, double ay, double bx, double by) {
    Ln l;
    bool swap = ax > bx;
    double lx = swap ? bx : ax, ly = swap ? by : ay;  // left end (or a when vertical)
    double rx = swap ? ax : bx, ry = swap ? ay : by;
    l.blx = lx;
    l.trx = rx;
    if (ax == bx) {
        l.parity = 1;
        l.bly = fmin(ay, by);
        l.try_ = fmax(ay, by);
    } else if (ly <= ry) {
        l.parity = 1;
        l.bly = ly;
        l.try_ = ry;
    } else {
        l.parity = 0;
        l.bly = ry;
        l.try_ = ly;
    }
    return l;
}

__device__ __forceinline__ double l_ay(const Ln &l) { return l.parity ? l.bly : l.try_; }
__device__ __forceinline__ double l_by(const Ln &l) { return l.parity ? l.try_ : l.bly; }
__device__ __forceinline__ double l_w(const Ln &l) { return fabs(sub(l.trx, l.blx)); }
__device__ __forceinline__ double l_h(const Ln &l) { return fabs(sub(l.try_, l.bly)); }

__device__ __forceinline__ bool overlap1(double abl, double atr, double bbl, double btr, double tol) {
    return (abl > bbl) ? (btr >= sub(abl, tol)) : (atr >= sub(bbl, tol));
}

// intersect_region && intersect_line (genlib/p2dpoly.cpp:247-279, 350-363)
__device__ __forceinline__ bool line_hits(const Ln &a, const Ln &b, double tol) {
    if (!(overlap1(a.blx, a.trx, b.blx, b.trx, tol) && overlap1(a.bly, a.try_, b.bly, b.try_, tol))) return false;
    double aax = a.blx, aay = l_ay(a), abx = a.trx, aby = l_by(a);
    double bax = b.blx, bay = l_ay(b), bbx = b.trx, bby = l_by(b);
    double ady = sub(aay, aby), adx = sub(abx, aax);
    double t1 = add(mul(ady, sub(bax, aax)), mul(adx, sub(bay, aay)));
    double t2 = add(mul(ady, sub(bbx, aax)), mul(adx, sub(bby, aay)));
    if (!(mul(t1, t2) <= tol)) return false;
    double bdy = sub(bay, bby), bdx = sub(bbx, bax);
    double t3 = add(mul(bdy, sub(aax, bax)), mul(bdx, sub(aay, bay)));
    double t4 = add(mul(bdy, sub(abx, bax)), mul(bdx, sub(aby, bay)));
    return mul(t3, t4) <= tol;
}

// Line::crop semantics (genlib/p2dpoly.cpp:626-667): clip order left, right, bottom, top, each with
// the current width/height.
__device__ bool crop_line(Ln &l, double rblx, double rbly, double rtrx, double rtry) {
    double sign = l.parity ? 1.0 : -1.0;
    if (!(l.trx >= rblx)) return false;
    if (l.blx < rblx) {
        double d = mul(sign, dvd(mul(l_h(l), sub(rblx, l.blx)), l_w(l)));
        if (l.parity)
            l.bly = add(l.bly, d);
        else
            l.try_ = add(l.try_, d);
        l.blx = rblx;
    }
    if (!(l.blx <= rtrx)) return false;
    if (l.trx > rtrx) {
        double d = dvd(mul(mul(sign, l_h(l)), sub(l.trx, rtrx)), l_w(l));
        if (l.parity)
            l.try_ = sub(l.try_, d);
        else
            l.bly = sub(l.bly, d);
        l.trx = rtrx;
    }
    if (!(l.try_ >= rbly)) return false;
    if (l.bly < rbly) {
        double d = dvd(mul(l_w(l), sub(rbly, l.bly)), l_h(l));
        if (l.parity)
            l.blx = add(l.blx, d);
        else
            l.trx = sub(l.trx, d);
        l.bly = rbly;
    }
    if (!(l.bly <= rtry)) return false;
    if (l.try_ > rtry) {
        double d = dvd(mul(l_w(l), sub(l.try_, rtry)), l_h(l));
        if (l.parity)
            l.trx = sub(l.trx, d);
        else
            l.blx = add(l.blx, d);
        l.try_ = rtry;
    }
    return true;
}

__device__ __forceinline__ double tanify(double cx, double cy, double px, double py, int q) {
    switch (q) {
    case 0: return dvd(sub(py, cy), sub(cx, px));
    case 1: return dvd(sub(py, cy), sub(px, cx));
    case 2: return dvd(sub(cy, py), sub(cx, px));
    case 3: return dvd(sub(cy, py), sub(px, cx));
    case 4: return dvd(sub(cx, px), sub(cy, py));
    case 5: return dvd(sub(px, cx), sub(cy, py));
    case 6: return dvd(sub(cx, px), sub(py, cy));
    default: return dvd(sub(px, cx), sub(py, cy));
    }
}

// angular interval covered by a wall segment, padded by 1e-10 (sparksieve2.cpp:67-82)
__device__ __forceinline__ Zone make_block(const Ln &l, double cx, double cy, int q) {
    double a = tanify(cx, cy, l.blx, l_ay(l), q);
    double b = tanify(cx, cy, l.trx, l_by(l), q);
    Zone z;
    if (a < b) {
        z.s = sub(a, 1e-10);
        z.e = add(b, 1e-10);
    } else {
        z.s = sub(b, 1e-10);
        z.e = add(a, 1e-10);
    }
    return z;
}

__device__ __forceinline__ bool zone_less(const Zone &a, const Zone &b) {
    return (a.s == b.s) ? (a.e > b.e) : (a.s < b.s);
}

// direction bin of a target (salalib/pointdata.h:432-520 semantics)
__device__ __forceinline__ int which_bin(double gx, double gy) {
    double ax = fabs(gx), ay = fabs(gy);
    int bin;
    double ratio;
    if (ay > ax) {
        ratio = dvd(ax, ay);
        bin = (gy > 0.0) ? ((gx >= 0.0) ? -8 : 8) : ((gx >= 0.0) ? 24 : -24);
    } else {
        ratio = dvd(ay, ax);
        bin = (gx > 0.0) ? ((gy >= 0.0) ? 0 : -32) : ((gy >= 0.0) ? -16 : 16);
    }
    if (ratio < 1e-12) {
    } else if (ratio < 0.2679491924311227)
        bin += 1;
    else if (ratio < 0.5773502691896257)
        bin += 2;
    else if (ratio < 1.0 - 1e-12)
        bin += 3;
    else
        bin += 4;
    if (bin < 0) bin = -bin;
    return bin % 32;
}

__device__ __forceinline__ int warp_incl_sum(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += t;
    }
    return v;
}
__device__ __forceinline__ int warp_incl_max(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v = max(v, t);
    }
    return v;
}

// Per-warp scratch (shared or global memory)
struct Scratch {
    Zone *ga, *gb, *blk;
    int *rstart, *rprefix;
    int gcap, bcap;
};

enum { TASK_OK = 0, TASK_OVERFLOW = 1, TASK_NAN = 2 };

// Sort the row's blocks and subtract them from the gap list.  Returns the new gap count
// (written to gout) or -1 on capacity overflow.
__device__ int subtract_blocks(const Zone *gin, int ng, Zone *blk, int nb, Zone *gout, int gcap, int bcap, int lane) {
    if (nb > 1) {
        int P = 1;
        while (P < nb) P <<= 1;  // P <= bcap is guaranteed by the caller (bcap is a power of two)
        for (int i = nb + lane; i < P; i += 32) {
            blk[i].s = __longlong_as_double(0x7ff0000000000000LL);   // +inf
            blk[i].e = __longlong_as_double(0xfff0000000000000LL);   // -inf: sorts after every real block
        }
        __syncwarp();
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = lane; i < P; i += 32) {
                    int p = i ^ j;
                    if (p > i) {
                        Zone a = blk[i], b = blk[p];
                        bool up = ((i & k) == 0);
                        if (zone_less(b, a) == up) {
                            blk[i] = b;
                            blk[p] = a;
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    int ngo = 0;
    if (lane == 0) {
        // sequential interval subtraction (sparksieve2.cpp:89-132), duplicates of the sorted block
        // list skipped on the fly (std::unique), output streamed to gout
        int gi = 0, bi = 0;
        Zone cur = gin[0];
        bool over = false;
        while (bi < nb && gi < ng) {
            Zone B = blk[bi];
            if (bi > 0) {
                Zone Bp = blk[bi - 1];
                if (Bp.s == B.s && Bp.e == B.e) {
                    bi++;
                    continue;
                }
            }
            if (B.e < cur.s) {
                bi++;
                continue;
            }
            bool create = true;
            if (B.s <= cur.s) {
                create = false;
                if (B.e > cur.s) cur.s = B.e;
            }
            if (B.e >= cur.e) {
                create = false;
                if (B.s < cur.e) cur.e = B.s;
            }
            if (cur.e <= add(cur.s, 1e-10)) {
                gi++;
                if (gi < ng) cur = gin[gi];
                continue;
            } else if (B.e > cur.e) {
                if (ngo < gcap) gout[ngo] = cur; else over = true;
                ngo++;
                gi++;
                if (gi < ng) cur = gin[gi];
                continue;
            } else if (create) {
                Zone left;
                left.s = cur.s;
                left.e = B.s;
                if (ngo < gcap) gout[ngo] = left; else over = true;
                ngo++;
                cur.s = B.e;
            }
            bi++;
        }
        if (gi < ng) {
            if (ngo < gcap) gout[ngo] = cur; else over = true;
            ngo++;
            for (int r = gi + 1; r < ng; r++) {
                if (ngo < gcap) gout[ngo] = gin[r]; else over = true;
                ngo++;
            }
        }
        if (over) ngo = -1;
    }
    ngo = __shfl_sync(FULL, ngo, 0);
    __syncwarp();
    return ngo;
}

struct TaskOut {
    uint32_t nacc, nfill;
    int status;
};

template <bool EMIT>
__device__ TaskOut sieve_task(const GridDev &g, int cx, int cy, int q, const Scratch &sc, uint8_t *ghostflag,
                              uint64_t acc_off, uint64_t fill_off, uint32_t *e_ref, uint8_t *e_bin) {
    const int lane = threadIdx.x & 31;
    const unsigned ltmask = (1u << lane) - 1u;
    const double s = g.spacing;
    const double c0x = add(g.blx, mul(s, (double)cx));
    const double c0y = add(g.bly, mul(s, (double)cy));
    const double tol = mul(s, 1e-10);
    const bool xmajor = q < 4;
    const int sx = (q & 1) ? 1 : -1;
    const int sy = (q <= 1 || q >= 6) ? 1 : -1;
    const int diagbin = (q == 0) ? 12 : (q == 1) ? 4 : (q == 2) ? 20 : 28;

    Zone *G = sc.ga, *Gn = sc.gb;
    TaskOut out;
    out.nacc = 0;
    out.nfill = 0;
    out.status = TASK_OK;
    int ng = 1;
    if (lane == 0) {
        G[0].s = 0.0;
        G[0].e = 1.0;
    }
    __syncwarp();
    bool bad_nan = false;

    // ---- depth 0: own-cell wall lines clipped to the octant's quadrant (pointdata.cpp:1399-1449)
    {
        int64_t c = (int64_t)cx * g.rows + cy;
        uint32_t lo = g.line_off[c];
        int nl = (int)(g.line_off[c + 1] - lo);
        if (nl > 0) {
            double fx = (double)cx, fy = (double)cy;
            double vblx = add(g.blx, mul(s, sub(sub(fx, 0.5), 1e-10)));
            double vbly = add(g.bly, mul(s, sub(sub(fy, 0.5), 1e-10)));
            double vtrx = add(g.blx, mul(s, add(add(fx, 0.5), 1e-10)));
            double vtry = add(g.bly, mul(s, add(add(fy, 0.5), 1e-10)));
            const double border = tol;
            switch (q) {
            case 0: vtrx = c0x; vbly = sub(c0y, border); break;
            case 6: vtrx = add(c0x, border); vbly = c0y; break;
            case 1: vblx = c0x; vbly = sub(c0y, border); break;
            case 7: vblx = sub(c0x, border); vbly = c0y; break;
            case 2: vtrx = c0x; vtry = add(c0y, border); break;
            case 4: vtrx = add(c0x, border); vtry = c0y; break;
            case 3: vblx = c0x; vtry = add(c0y, border); break;
            default: vblx = sub(c0x, border); vtry = c0y; break;
            }
            int nb = 0;
            for (int base = 0; base < nl; base += 32) {
                int i = base + lane;
                bool ok = false;
                Zone z;
                z.s = z.e = 0.0;
                if (i < nl) {
                    Ln l = load_line(g.lines + 5 * (size_t)(lo + i));
                    ok = crop_line(l, vblx, vbly, vtrx, vtry);
                    if (ok) {
                        z = make_block(l, c0x, c0y, q);
                        if (z.s != z.s || z.e != z.e) bad_nan = true;
                    }
                }
                unsigned m = __ballot_sync(FULL, ok);
                if (ok) {
                    int pos = nb + __popc(m & ltmask);
                    if (pos < sc.bcap) sc.blk[pos] = z;
                }
                nb += __popc(m);
            }
            __syncwarp();
            if (nb > sc.bcap) {
                out.status = TASK_OVERFLOW;
                return out;
            }
            if (nb > 0) {
                ng = subtract_blocks(G, ng, sc.blk, nb, Gn, sc.gcap, sc.bcap, lane);
                if (ng < 0) {
                    out.status = TASK_OVERFLOW;
                    return out;
                }
                Zone *t = G;
                G = Gn;
                Gn = t;
            }
        }
    }

    uint32_t nacc = 0, nfill = 0;
    int diag_last = 0;

    for (int depth = 1; ng > 0; depth++) {
        // ---- enumerate the union of the gaps' widened index ranges (pointdata.cpp:1522-1534)
        const double dd = (double)depth;
        int T = 0;
        int run_max = -1;
        for (int base = 0; base < ng; base += 32) {
            int i = base + lane;
            int lo = 0, hi = -1;
            if (i < ng) {
                Zone z = G[i];
                lo = (int)ceil(sub(mul(z.s, sub(dd, 0.5)), 0.5));
                hi = (int)floor(add(mul(z.e, add(dd, 0.5)), 0.5));
                hi = min(hi, depth);
            }
            bool nonempty = (i < ng) && (lo <= hi);
            int hv = nonempty ? hi : -1;
            int imax = warp_incl_max(hv, lane);
            int emax = __shfl_up_sync(FULL, imax, 1);
            if (lane == 0) emax = -1;
            emax = max(emax, run_max);
            int st = max(lo, emax + 1);
            int len = nonempty ? max(0, hi - st + 1) : 0;
            int isum = warp_incl_sum(len, lane);
            if (i < ng) {
                sc.rstart[i] = st;
                sc.rprefix[i] = T + isum - len;
            }
            T += __shfl_sync(FULL, isum, 31);
            run_max = max(run_max, __shfl_sync(FULL, imax, 31));
        }
        if (lane == 0) sc.rprefix[ng] = T;
        __syncwarp();

        bool anyin = false;
        int nbrow = 0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            int t = t0 + lane;
            bool act = t < T;
            int ind = 0, gi = 0;
            if (act) {
                int lo = 0, hi = ng;  // last gi with rprefix[gi] <= t
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (sc.rprefix[mid] <= t) lo = mid; else hi = mid;
                }
                gi = lo;
                ind = sc.rstart[gi] + (t - sc.rprefix[gi]);
            }
            int ox = xmajor ? depth : ind;
            int oy = xmajor ? ind : depth;
            int hx = cx + sx * ox, hy = cy + sy * oy;
            bool ing = act && hx >= 0 && hx < g.cols && hy >= 0 && hy < g.rows;
            anyin |= (__ballot_sync(FULL, ing) != 0);
            uint8_t fl = 0;
            int64_t c = (int64_t)hx * g.rows + hy;
            if (ing) fl = xmajor ? g.cflag[c] : g.cflag_t[(int64_t)hy * g.cols + hx];
            const bool haslines = (fl & 2) != 0;
            uint32_t llo = 0;
            int nl = 0;
            if (haslines) {
                llo = g.line_off[c];
                nl = (int)(g.line_off[c + 1] - llo);
            }
            bool cand = false;
            if (ing && (fl & 1)) {
                // don't repeat axes / diagonals (pointdata.cpp:1551)
                if ((ind != 0 || q == 0 || q == 1 || q == 5 || q == 6) && (ind != depth || q < 4)) {
                    double di = (double)ind;
                    for (int k = gi; k < ng; k++) {
                        Zone z = G[k];
                        if (di < mul(z.s, dd)) break;
                        if (di <= mul(z.e, dd)) {
                            cand = true;
                            break;
                        }
                    }
                }
            }
            double px = 0.0, py = 0.0;
            if (cand) {
                px = add(g.blx, mul(s, (double)hx));
                py = add(g.bly, mul(s, (double)hy));
                if (nl > 0 || g.maxdist != -1.0) {
                    Ln l = make_line(c0x, c0y, px, py);
                    if (g.maxdist != -1.0) {
                        double w = sub(l.trx, l.blx), h = sub(l.try_, l.bly);
                        double len = __dsqrt_rn(add(mul(w, w), mul(h, h)));
                        if (len > g.maxdist) cand = false;
                    }
                    for (int j = 0; cand && j < nl; j++) {
                        Ln wl = load_line(g.lines + 5 * (size_t)(llo + j));
                        if (line_hits(l, wl, tol)) cand = false;
                    }
                }
            }
            unsigned am = __ballot_sync(FULL, cand);
            // diagonal bins are stored by the reference as ONE run first..last (ngraph.cpp:243-259):
            // cells between two accepted diagonal cells that were not accepted themselves still
            // belong to the iterated adjacency ("fill-ins")
            unsigned dm = __ballot_sync(FULL, cand && xmajor && ind == depth);
            if (dm) {
                if (diag_last > 0) {
                    int gap = depth - 1 - diag_last;
                    for (int d = diag_last + 1 + lane; d < depth; d += 32) {
                        int fx = cx + sx * d, fy = cy + sy * d;
                        int64_t fc = (int64_t)fx * g.rows + fy;
                        if (EMIT) {
                            uint64_t p = fill_off + nfill + (uint32_t)(d - diag_last - 1);
                            e_ref[p] = ((uint32_t)fx << 16) | ((uint32_t)fy & 0xffffu);
                            e_bin[p] = (uint8_t)(diagbin | 0x80);
                        } else if (!(g.cflag[fc] & 1)) {
                            *ghostflag = 1;  // an unfilled cell is part of the iterated adjacency
                        }
                    }
                    nfill += (uint32_t)gap;
                }
                diag_last = depth;
            }
            if (EMIT && cand) {
                uint64_t p = acc_off + nacc + (uint32_t)__popc(am & ltmask);
                e_ref[p] = ((uint32_t)hx << 16) | ((uint32_t)hy & 0xffffu);
                e_bin[p] = (uint8_t)which_bin(sub(px, c0x), sub(py, c0y));
            }
            nacc += (uint32_t)__popc(am);

            // every visited in-grid cell contributes its wall segments as blocks (pointdata.cpp:1558)
            int isum = warp_incl_sum(nl, lane);
            int tot = __shfl_sync(FULL, isum, 31);
            if (tot > 0) {
                int off = nbrow + isum - nl;
                for (int j = 0; j < nl; j++) {
                    Ln wl = load_line(g.lines + 5 * (size_t)(llo + j));
                    Zone z = make_block(wl, c0x, c0y, q);
                    if (z.s != z.s || z.e != z.e) bad_nan = true;
                    if (off + j < sc.bcap) sc.blk[off + j] = z;
                }
                nbrow += tot;
            }
        }
        __syncwarp();
        if (nbrow > sc.bcap) {
            out.status = TASK_OVERFLOW;
            return out;
        }
        if (nbrow > 0) {
            ng = subtract_blocks(G, ng, sc.blk, nbrow, Gn, sc.gcap, sc.bcap, lane);
            if (ng < 0) {
                out.status = TASK_OVERFLOW;
                return out;
            }
            Zone *t = G;
            G = Gn;
            Gn = t;
        }
        if (!anyin) break;
    }
    if (__any_sync(FULL, bad_nan)) out.status = TASK_NAN;
    out.nacc = nacc;
    out.nfill = nfill;
    return out;
}

struct SieveArgs {
    GridDev g;
    int64_t src_begin;      // first source ordinal of this launch's numbering
    int64_t ntasks;         // tasks in this launch
    const int64_t *tasklist;  // explicit global task ids (big mode) or nullptr = task range
    int64_t task_begin;     // first global task id when tasklist == nullptr
    int gcap, bcap;
    Zone *gscratch;         // nullptr = shared memory
    uint8_t *bigflag;       // [ntasks_total] 1 = task needs the big-capacity launch
    // count outputs
    uint32_t *cnt;          // [nsrc*8]
    uint32_t *fillcnt;      // [nsrc*4]
    uint8_t *ghostflag;     // [1] set when any unfilled cell is covered by a diagonal run
    int64_t *overflow_list;
    unsigned long long *n_overflow;
    int *error_flag;
    // emit inputs/outputs (offsets relative to the chunk's temp buffers)
    const uint64_t *row_off;  // [nsrc+1] global row offsets
    uint64_t chunk_base;      // subtracted from row_off
    uint32_t *e_ref;
    uint8_t *e_bin;
    const uint32_t *order;    // thread kernel, emit pass: thread t runs task slot order[t] (tasks sorted by size), or nullptr
};

template <bool EMIT> __global__ void __launch_bounds__(WARPS_PER_BLOCK * 32) k_sieve(SieveArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    Scratch sc;
    sc.gcap = a.gcap;
    sc.bcap = a.bcap;
    const size_t per_warp = (size_t)(2 * a.gcap + a.bcap) * sizeof(Zone) + (size_t)(2 * a.gcap + 2) * sizeof(int);
    const bool big = a.gscratch != nullptr;
    int64_t wglobal = (int64_t)blockIdx.x * WARPS_PER_BLOCK + wib;
    int64_t wstride = (int64_t)gridDim.x * WARPS_PER_BLOCK;
    {
        unsigned char *base = big ? ((unsigned char *)a.gscratch + (size_t)wglobal * ((per_warp + 15) / 16 * 16))
                                  : (smem + (size_t)wib * ((per_warp + 15) / 16 * 16));
        sc.ga = (Zone *)base;
        sc.gb = sc.ga + a.gcap;
        sc.blk = sc.gb + a.gcap;
        sc.rstart = (int *)(sc.blk + a.bcap);
        sc.rprefix = sc.rstart + a.gcap;
    }
    for (int64_t w = wglobal; w < a.ntasks; w += wstride) {
        int64_t task = a.tasklist ? a.tasklist[w] : (a.task_begin + w);
        if (!big && a.bigflag[task]) continue;  // handled by the big-capacity launch
        int64_t src_local = task >> 3;          // relative to src_begin
        int q = (int)(task & 7);
        int32_t ref = a.g.cellref[a.src_begin + src_local];
        int cx = ref >> 16, cy = ref & 0xffff;
        uint64_t acc_off = 0, fill_off = 0;
        if (EMIT) {
            uint64_t ro = a.row_off[src_local] - a.chunk_base;
            uint32_t before = 0, all = 0;
            for (int k = 0; k < 8; k++) {
                uint32_t cv = a.cnt[src_local * 8 + k];
                if (k < q) before += cv;
                all += cv;
            }
            acc_off = ro + before;
            uint32_t fb = 0;
            for (int k = 0; k < 4 && k < q; k++) fb += a.fillcnt[src_local * 4 + k];
            fill_off = ro + all + fb;
        }
        TaskOut o = sieve_task<EMIT>(a.g, cx, cy, q, sc, a.ghostflag, acc_off, fill_off, a.e_ref, a.e_bin);
        if (lane == 0) {
            if (o.status == TASK_OVERFLOW) {
                if (!big) {
                    a.bigflag[task] = 1;
                    unsigned long long p = atomicAdd(a.n_overflow, 1ULL);
                    a.overflow_list[p] = task;
                } else {
                    atomicExch(a.error_flag, VGA_ERR_CAPACITY);
                }
            } else if (o.status == TASK_NAN) {
                atomicExch(a.error_flag, VGA_ERR_UNSUPPORTED);
            } else if (!EMIT) {
                a.cnt[src_local * 8 + q] = o.nacc;
                if (q < 4) a.fillcnt[src_local * 4 + q] = o.nfill;
            }
        }
        __syncwarp();
    }
}

// ---- thread-per-task variant ------------------------------------------------------------------
// One THREAD per (source, octant).  ncu on the warp kernel (profiles/r1_*) showed ~230 warp
// instructions per depth row at 45 % lane use and a long tail (8 warps of a CTA wait for the slowest
// octant): on plans with short rows (rooms, corridors) a row holds only a handful of cells, so the
// warp-wide scans cost more than the cells.  Here every thread walks its own octant sequentially
// (the same per-gap, per-index order as sieve2), 32 consecutive threads take the same octant of 32
// consecutive cells (octant-major numbering) so that warps stay convergent, the gap list lives in
// shared memory (TG entries per thread, interleaved), the row's blocks in local memory.  Tasks that
// exceed TG gaps / TB blocks are flagged and re-run by the warp kernel with global scratch.
constexpr int TT = 128;  // threads per CTA
// TG = gaps per thread (shared memory), TB = blocks per row per thread (local memory): (8, 24) by default, (16, 48) with
// sieve_thread_cap = 1 (fewer tasks fall back to the warp kernel, at twice the shared memory per CTA)

template <bool EMIT, int TG, int TB> __global__ void __launch_bounds__(TT) k_sieve_thread(SieveArgs a, int64_t nsrc) {
    __shared__ Zone s_gaps[TG * TT];
    const int tid = threadIdx.x;
    int64_t t = (int64_t)blockIdx.x * TT + tid;
    if (t >= a.ntasks) return;
    if (EMIT && a.order) t = a.order[t];
    const int q = (int)(t / nsrc);
    const int64_t src_local = t % nsrc;
    const int64_t task = src_local * 8 + q;
    if (a.bigflag[task]) return;
    const GridDev &g = a.g;
    const int32_t ref = g.cellref[a.src_begin + src_local];
    const int cx = ref >> 16, cy = ref & 0xffff;
    uint64_t acc_off = 0, fill_off = 0;
    if (EMIT) {
        uint64_t ro = a.row_off[src_local] - a.chunk_base;
        uint32_t before = 0, all = 0;
        for (int k = 0; k < 8; k++) {
            uint32_t cv = a.cnt[src_local * 8 + k];
            if (k < q) before += cv;
            all += cv;
        }
        acc_off = ro + before;
        uint32_t fb = 0;
        for (int k = 0; k < 4 && k < q; k++) fb += a.fillcnt[src_local * 4 + k];
        fill_off = ro + all + fb;
    }
#define GAP(i) s_gaps[(i) * TT + tid]
    const double s = g.spacing;
    const double c0x = add(g.blx, mul(s, (double)cx));
    const double c0y = add(g.bly, mul(s, (double)cy));
    const double tol = mul(s, 1e-10);
    const bool xmajor = q < 4;
    const int sx = (q & 1) ? 1 : -1;
    const int sy = (q <= 1 || q >= 6) ? 1 : -1;
    const int diagbin = (q == 0) ? 12 : (q == 1) ? 4 : (q == 2) ? 20 : 28;
    Zone blk[TB];
    int ng = 1, nb = 0;
    {
        Zone z0;
        z0.s = 0.0;
        z0.e = 1.0;
        GAP(0) = z0;
    }
    bool overflow = false, bad_nan = false;

    // sort the row's blocks (insertion sort, start asc / end desc) and subtract them from the gap
    // list in place, exactly the sequential rule of sparksieve2.cpp:89-132
    auto subtract = [&]() {
        for (int i = 1; i < nb; i++) {
            Zone x = blk[i];
            int j = i - 1;
            while (j >= 0 && zone_less(x, blk[j])) {
                blk[j + 1] = blk[j];
                j--;
            }
            blk[j + 1] = x;
        }
        int gi = 0, bi = 0;
        while (bi < nb && gi < ng) {
            Zone B = blk[bi];
            if (bi > 0 && blk[bi - 1].s == B.s && blk[bi - 1].e == B.e) {
                bi++;
                continue;
            }
            Zone G = GAP(gi);
            if (B.e < G.s) {
                bi++;
                continue;
            }
            bool create = true;
            if (B.s <= G.s) {
                create = false;
                if (B.e > G.s) G.s = B.e;
            }
            if (B.e >= G.e) {
                create = false;
                if (B.s < G.e) G.e = B.s;
            }
            if (G.e <= add(G.s, 1e-10)) {
                for (int r = gi; r + 1 < ng; r++) GAP(r) = GAP(r + 1);
                ng--;
                continue;
            } else if (B.e > G.e) {
                GAP(gi) = G;
                gi++;
                continue;
            } else if (create) {
                if (ng == TG) {
                    overflow = true;
                    return;
                }
                for (int r = ng; r > gi + 1; r--) GAP(r) = GAP(r - 1);
                ng++;
                Zone left;
                left.s = G.s;
                left.e = B.s;
                GAP(gi) = left;
                G.s = B.e;
                GAP(gi + 1) = G;
                gi++;
            } else {
                GAP(gi) = G;
            }
            bi++;
        }
        nb = 0;
    };
    auto push_blocks = [&](uint32_t llo, int nl) {
        for (int j = 0; j < nl; j++) {
            if (nb == TB) {
                overflow = true;
                return;
            }
            Ln wl = load_line(g.lines + 5 * (size_t)(llo + j));
            Zone z = make_block(wl, c0x, c0y, q);
            if (z.s != z.s || z.e != z.e) bad_nan = true;
            blk[nb++] = z;
        }
    };

    // ---- depth 0: own-cell lines clipped to the octant's quadrant
    {
        int64_t c = (int64_t)cx * g.rows + cy;
        uint32_t lo = g.line_off[c];
        int nl = (int)(g.line_off[c + 1] - lo);
        if (nl > 0) {
            double fx = (double)cx, fy = (double)cy;
            double vblx = add(g.blx, mul(s, sub(sub(fx, 0.5), 1e-10)));
            double vbly = add(g.bly, mul(s, sub(sub(fy, 0.5), 1e-10)));
            double vtrx = add(g.blx, mul(s, add(add(fx, 0.5), 1e-10)));
            double vtry = add(g.bly, mul(s, add(add(fy, 0.5), 1e-10)));
            switch (q) {
            case 0: vtrx = c0x; vbly = sub(c0y, tol); break;
            case 6: vtrx = add(c0x, tol); vbly = c0y; break;
            case 1: vblx = c0x; vbly = sub(c0y, tol); break;
            case 7: vblx = sub(c0x, tol); vbly = c0y; break;
            case 2: vtrx = c0x; vtry = add(c0y, tol); break;
            case 4: vtrx = add(c0x, tol); vtry = c0y; break;
            case 3: vblx = c0x; vtry = add(c0y, tol); break;
            default: vblx = sub(c0x, tol); vtry = c0y; break;
            }
            for (int i = 0; i < nl && !overflow; i++) {
                Ln l = load_line(g.lines + 5 * (size_t)(lo + i));
                if (crop_line(l, vblx, vbly, vtrx, vtry)) {
                    if (nb == TB) {
                        overflow = true;
                        break;
                    }
                    Zone z = make_block(l, c0x, c0y, q);
                    if (z.s != z.s || z.e != z.e) bad_nan = true;
                    blk[nb++] = z;
                }
            }
            if (!overflow && nb > 0) subtract();
        }
    }

    uint32_t nacc = 0, nfill = 0;
    int diag_last = 0;
    for (int depth = 1; ng > 0 && !overflow; depth++) {
        const double dd = (double)depth;
        bool hasgaps = false;
        int firstind = 0;
        for (int gi = 0; gi < ng && !overflow; gi++) {
            const Zone z = GAP(gi);
            const int lo = (int)ceil(sub(mul(z.s, sub(dd, 0.5)), 0.5));
            int hi = (int)floor(add(mul(z.e, add(dd, 0.5)), 0.5));
            if (hi > depth) hi = depth;
            const double lo_c = mul(z.s, dd), hi_c = mul(z.e, dd);
            for (int ind = max(lo, firstind); ind <= hi; ind++) {
                firstind = ind;
                const int ox = xmajor ? depth : ind, oy = xmajor ? ind : depth;
                const int hx = cx + sx * ox, hy = cy + sy * oy;
                if (hx < 0 || hx >= g.cols || hy < 0 || hy >= g.rows) continue;
                hasgaps = true;
                const int64_t c = (int64_t)hx * g.rows + hy;
                const uint8_t fl = g.cflag[c];
                uint32_t llo = 0;
                int nl = 0;
                if (fl & 2) {
                    llo = g.line_off[c];
                    nl = (int)(g.line_off[c + 1] - llo);
                }
                const double di = (double)ind;
                if ((fl & 1) && di >= lo_c && di <= hi_c &&
                    (ind != 0 || q == 0 || q == 1 || q == 5 || q == 6) && (ind != depth || q < 4)) {
                    const double px = add(g.blx, mul(s, (double)hx));
                    const double py = add(g.bly, mul(s, (double)hy));
                    bool ok = true;
                    if (nl > 0 || g.maxdist != -1.0) {
                        Ln l = make_line(c0x, c0y, px, py);
                        if (g.maxdist != -1.0) {
                            double w = sub(l.trx, l.blx), h = sub(l.try_, l.bly);
                            if (__dsqrt_rn(add(mul(w, w), mul(h, h))) > g.maxdist) ok = false;
                        }
                        for (int j = 0; ok && j < nl; j++) {
                            Ln wl = load_line(g.lines + 5 * (size_t)(llo + j));
                            if (line_hits(l, wl, tol)) ok = false;
                        }
                    }
                    if (ok) {
                        if (xmajor && ind == depth) {
                            // diagonal bin: cells skipped between two accepted diagonal cells are fill-ins
                            if (diag_last > 0) {
                                for (int d = diag_last + 1; d < depth; d++) {
                                    const int fx = cx + sx * d, fy = cy + sy * d;
                                    if (EMIT) {
                                        uint64_t p = fill_off + nfill + (uint32_t)(d - diag_last - 1);
                                        a.e_ref[p] = ((uint32_t)fx << 16) | ((uint32_t)fy & 0xffffu);
                                        a.e_bin[p] = (uint8_t)(diagbin | 0x80);
                                    } else if (!(g.cflag[(int64_t)fx * g.rows + fy] & 1)) {
                                        *a.ghostflag = 1;
                                    }
                                }
                                nfill += (uint32_t)(depth - 1 - diag_last);
                            }
                            diag_last = depth;
                        }
                        if (EMIT) {
                            uint64_t p = acc_off + nacc;
                            a.e_ref[p] = ((uint32_t)hx << 16) | ((uint32_t)hy & 0xffffu);
                            a.e_bin[p] = (uint8_t)which_bin(sub(px, c0x), sub(py, c0y));
                        }
                        nacc++;
                    }
                }
                if (nl > 0) push_blocks(llo, nl);
            }
        }
        if (overflow) break;
        if (nb > 0) subtract();
        if (!hasgaps) break;
    }
#undef GAP
    if (overflow) {
        a.bigflag[task] = 1;
        unsigned long long p = atomicAdd(a.n_overflow, 1ULL);
        a.overflow_list[p] = task;
    } else if (bad_nan) {
        atomicExch(a.error_flag, VGA_ERR_UNSUPPORTED);
    } else if (!EMIT) {
        a.cnt[src_local * 8 + q] = nacc;
        if (q < 4) a.fillcnt[src_local * 4 + q] = nfill;
    }
}

// Emit pass of the thread kernel: the 32 tasks of a warp are 32 consecutive sources looking into the same octant, and their
// sizes differ (a warp is as slow as its largest task: lane efficiency 0.56 by accepted cells on an urban plan, measured with
// the oracle's rows).  The counts of pass 1 are known, so the emit pass runs the tasks in order of decreasing size: key =
// complement of the accepted cells / 16 (ties keep their spatial order: the radix sort is stable), value = thread slot.
__global__ void k_task_keys(const uint32_t *cnt, int64_t nsrc, uint32_t *key, uint32_t *val) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nsrc * 8) return;
    const int q = (int)(t / nsrc);
    const int64_t s = t % nsrc;
    key[t] = 0x0fffffffu - min(cnt[s * 8 + q] >> 4, 0x0fffffffu);
    val[t] = (uint32_t)t;
}

// per-source row sizes -> totals (accepted + fill-ins)
__global__ void k_row_totals(const uint32_t *cnt, const uint32_t *fillcnt, int64_t nsrc, uint64_t *tot) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nsrc) return;
    uint64_t t = 0;
    for (int k = 0; k < 8; k++) t += cnt[i * 8 + k];
    for (int k = 0; k < 4; k++) t += fillcnt[i * 4 + k];
    tot[i] = t;
}

// Connectivity, moments, far distances, bin counts, grid connections of a source from its unsorted row (pointdata.cpp:
// 1463-1497, 1735-1768).  One WARP per source: 32 entries are loaded at a time (coalesced) and their distances computed in
// parallel; the per-bin maxima and counts and the grid-connection bits do not depend on the order (the float maximum of
// the reference's running comparison equals the maximum of the rounded values: rounding is monotone), but the two moments
// are double running sums whose rounding depends on the order of the additions, so lane 0 / lane 1 add the 32 distances /
// squares of a round one after the other, in row order, exactly like the reference's loop.
constexpr int NS_TPB = 256;
__global__ void __launch_bounds__(NS_TPB) k_node_stats(GridDev g, int64_t src_begin, int64_t chunk_first, int64_t nsrc_chunk,
                                                       const uint64_t *row_off, uint64_t chunk_base, const uint32_t *cnt,
                                                       const uint32_t *e_ref, const uint8_t *e_bin, int32_t *connectivity,
                                                       double *sum_d, double *sum_d2, float *far_dist, int32_t *bin_count,
                                                       uint8_t *gridconn) {
    __shared__ unsigned s_far[NS_TPB / 32][32];
    __shared__ int s_bc[NS_TPB / 32][32];
    __shared__ double s_d[NS_TPB / 32][32], s_d2[NS_TPB / 32][32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t i = ((int64_t)blockIdx.x * NS_TPB + threadIdx.x) >> 5;
    if (i >= nsrc_chunk) return;  // whole warps leave
    const int64_t sl = chunk_first + i;  // source index relative to src_begin
    const int32_t ref = g.cellref[src_begin + sl];
    const int cx = ref >> 16, cy = ref & 0xffff;
    s_far[wid][lane] = 0u;  // bits of 0.0f
    s_bc[wid][lane] = 0;
    const uint64_t beg = row_off[sl] - chunk_base, end = row_off[sl + 1] - chunk_base;
    uint32_t nacc = 0;
    for (int k = 0; k < 8; k++) nacc += cnt[sl * 8 + k];
    double acc = 0.0;  // lane 0: sum of distances, lane 1: sum of squares
    unsigned gc = 0u;
    __syncwarp();
    for (uint64_t e0 = beg; e0 < end; e0 += 32) {
        const uint64_t e = e0 + lane;
        double d = 0.0;
        if (e < end) {
            const uint32_t r = e_ref[e];
            const int b = e_bin[e] & 31;
            const int dx = (int)(r >> 16) - cx, dy = (int)(r & 0xffff) - cy;
            if (e - beg < nacc) {  // accepted cells come first, the fill-ins of diagonal bins after them
                const double fx = (double)dx, fy = (double)dy;
                d = mul(__dsqrt_rn(add(mul(fx, fx), mul(fy, fy))), g.spacing);
                atomicMax(&s_far[wid][b], __float_as_uint((float)d));  // non-negative floats order like their bits
                atomicAdd(&s_bc[wid][b], 1);
            }
            if (dx >= -1 && dx <= 1 && dy >= -1 && dy <= 1) {
                // neighbour i of the 8-neighbourhood (E, NE, N, NW, W, SW, S, SE) is looked up in bin 4i
                const int idx_of[9] = {5, 6, 7, 4, -1, 0, 3, 2, 1};  // [(dy+1)*3 + (dx+1)]
                const int idx = idx_of[(dy + 1) * 3 + (dx + 1)];
                if (idx >= 0 && b == idx * 4) gc |= 1u << idx;
            }
        }
        s_d[wid][lane] = d;
        s_d2[wid][lane] = mul(d, d);
        __syncwarp();
        const uint64_t done = e0 - beg;
        const int m = done >= nacc ? 0 : (int)min((uint64_t)32, (uint64_t)nacc - done);  // accepted entries of this round
        if (lane == 0)
            for (int l = 0; l < m; l++) acc = add(acc, s_d[wid][l]);
        else if (lane == 1)
            for (int l = 0; l < m; l++) acc = add(acc, s_d2[wid][l]);
        __syncwarp();
    }
    gc = __reduce_or_sync(0xffffffffu, gc);
    if (lane == 0) {
        connectivity[sl] = (int32_t)nacc;
        sum_d[sl] = acc;
        gridconn[sl] = (uint8_t)gc;
    } else if (lane == 1) {
        sum_d2[sl] = acc;
    }
    far_dist[sl * 32 + lane] = __uint_as_float(s_far[wid][lane]);
    bin_count[sl * 32 + lane] = s_bc[wid][lane];
}

// sort keys: col<<6 | accepted<<5 | bin.  col = ordinal of a filled cell; an unfilled cell c (ghost)
// gets N + (number of unfilled cells before c) = N + c - (filled cells before c): a static numbering
// that every shard / rank computes identically.  cellord[c] holds -(1 + filled cells before c) for
// unfilled cells.
__global__ void k_make_keys(GridDev g, const uint32_t *e_ref, const uint8_t *e_bin, uint64_t n_entries, uint32_t *keys) {
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_entries) return;
    uint32_t r = e_ref[e];
    uint8_t b = e_bin[e];
    int64_t c = (int64_t)(r >> 16) * g.rows + (r & 0xffff);
    int32_t ord = g.cellord[c];
    uint32_t col = (ord >= 0) ? (uint32_t)ord : (uint32_t)(g.n + c - (int64_t)(-(ord + 1)));
    keys[e] = (col << 6) | ((b & 0x80) ? 0u : 32u) | (uint32_t)(b & 31);
}

// Row ordering (default; build_sort = 0 selects the CUB segmented radix sort, which also serves grids whose bitmap exceeds
// shared memory): a row is a SET of vertex numbers below `total`, so
// its sorted order follows from a bitmap of the row -- rank(col) = number of set bits below col.  One CTA per row: set
// the bits in shared memory, popc-sum groups of 32 words, scan the group sums, then every entry looks its rank up
// (group prefix + at most 31 word popcs + one masked popc) and is written to its final place.  O(total/32 + deg) per
// row instead of four radix passes over small segments.  Needs unique columns per row (the sieve's de-duplication rule
// guarantees it).  Shared memory: total/32 words + total/1024 group prefixes.
constexpr int RS_TPB = 256;
// shared memory of k_rank_sort for a universe of `total` vertices: bitmap words, group prefixes, per-word prefixes (u16), range
inline size_t rank_sort_smem(uint32_t total) {
    const size_t nw = ((size_t)total + 31u) >> 5, ng = (nw + 31u) >> 5;
    return sizeof(uint32_t) * (nw + ng + 4) + sizeof(uint16_t) * ((nw + 1) & ~(size_t)1);
}
// One CTA per row: the row's columns set as bits of a bitmap of the whole vertex universe in shared memory; rank(col) = set
// bits below col = prefix of its group of 32 words + prefix of its word inside the group (u16) + masked popc of the word.
// The bitmap is zero between rows and only the word range a row touches is counted, scanned and cleared again: a room cell
// sees a few columns of the plan, i.e. a narrow range of ordinals.
__global__ void __launch_bounds__(RS_TPB) k_rank_sort(const uint32_t *keys, const uint64_t *seg_off, int64_t nrows, uint32_t total,
                                                      uint32_t *out) {
    extern __shared__ __align__(16) uint32_t rs_smem[];
    const uint32_t nw = (total + 31u) >> 5, ng = (nw + 31u) >> 5;
    uint32_t *bm = rs_smem, *gpre = rs_smem + nw, *range = gpre + ng;
    uint16_t *wpre = reinterpret_cast<uint16_t *>(range + 4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = RS_TPB / 32;
    for (uint32_t w = threadIdx.x; w < nw; w += RS_TPB) bm[w] = 0u;
    if (threadIdx.x == 0) {
        range[0] = 0xffffffffu;
        range[1] = 0u;
    }
    __syncthreads();
    for (int64_t row = blockIdx.x; row < nrows; row += gridDim.x) {
        const uint64_t e0 = seg_off[row], e1 = seg_off[row + 1];
        uint32_t lo = 0xffffffffu, hi = 0u;
        for (uint64_t e = e0 + threadIdx.x; e < e1; e += RS_TPB) {
            const uint32_t col = keys[e] >> 6;
            atomicOr(&bm[col >> 5], 1u << (col & 31u));
            lo = min(lo, col >> 5);
            hi = max(hi, col >> 5);
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0 && lo <= hi) {
            atomicMin(&range[0], lo);
            atomicMax(&range[1], hi);
        }
        __syncthreads();
        const uint32_t wlo = range[0], whi = range[1];
        if (wlo <= whi) {  // block-uniform
            const uint32_t glo = wlo >> 5, ghi = whi >> 5;
            for (uint32_t g = glo + warp; g <= ghi; g += nwarps) {  // per word: prefix inside its group; per group: its sum
                const uint32_t w = g * 32u + lane;
                const unsigned c = (unsigned)(w < nw ? __popc(bm[w]) : 0);
                unsigned incl = c;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                if (w < nw) wpre[w] = (uint16_t)(incl - c);
                if (lane == 31) gpre[g] = incl;
            }
            __syncthreads();
            if (warp == 0) {  // exclusive scan of the group sums: every lane a contiguous slice, then a warp scan of the slices
                const uint32_t cnt = ghi - glo + 1u, per = (cnt + 31u) >> 5, b0 = glo + min(lane * per, cnt), b1 = glo + min((lane + 1u) * per, cnt);
                unsigned mine = 0;
                for (uint32_t g = b0; g < b1; g++) mine += gpre[g];
                unsigned incl = mine;
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += up;
                }
                unsigned run = incl - mine;
                for (uint32_t g = b0; g < b1; g++) {
                    const unsigned c = gpre[g];
                    gpre[g] = run;
                    run += c;
                }
            }
            __syncthreads();
            for (uint64_t e = e0 + threadIdx.x; e < e1; e += RS_TPB) {
                const uint32_t key = keys[e], col = key >> 6, w = col >> 5;
                const unsigned rank = gpre[w >> 5] + (unsigned)wpre[w] + (unsigned)__popc(bm[w] & ((1u << (col & 31u)) - 1u));
                out[e0 + rank] = key;
            }
            __syncthreads();
            // leave the bitmap zero for the next row
            for (uint32_t w = wlo + threadIdx.x; w <= whi; w += RS_TPB) bm[w] = 0u;
            if (threadIdx.x == 0) {
                range[0] = 0xffffffffu;
                range[1] = 0u;
            }
        }
        __syncthreads();
    }
}

__global__ void k_rebase(const uint64_t *in, int64_t n, uint64_t base, uint64_t *out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i] - base;
}

inline unsigned blocks_for(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

size_t per_warp_bytes(int gcap, int bcap) {
    size_t b = (size_t)(2 * gcap + bcap) * sizeof(Zone) + (size_t)(2 * gcap + 2) * sizeof(int);
    return (b + 15) / 16 * 16;
}

int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

int build_graph(vga_ctx *ctx, const vga_dgrid *dg, int64_t src_begin, int64_t src_end, vga_graph **out) {
    cudaStream_t st = ctx->stream;
    const int64_t N = dg->n;
    if (src_end < 0 || src_end > N) src_end = N;
    if (src_begin < 0) src_begin = 0;
    if (src_begin > src_end) src_begin = src_end;
    const int64_t nsrc = src_end - src_begin;
    if (dg->cells >= ((int64_t)1 << 26)) {
        set_error("vga_graph_build: more than 2^26 cells is not supported");
        return VGA_ERR_UNSUPPORTED;
    }

    GridDev g;
    g.cols = dg->cols;
    g.rows = dg->rows;
    g.spacing = dg->spacing;
    g.blx = dg->bl_x;
    g.bly = dg->bl_y;
    g.maxdist = dg->maxdist;
    g.cflag = dg->cflag.p;
    g.cflag_t = dg->cflag_t.p;
    g.line_off = dg->line_off.p;
    g.lines = dg->lines.p;
    g.cellord = dg->cellord.p;
    g.cellref = dg->cellref.p;
    g.n = N;

    std::unique_ptr<vga_graph> gr(new vga_graph());
    gr->ctx = ctx;
    gr->n = N;
    gr->src_begin = src_begin;
    gr->src_end = src_end;
    gr->has_bins = true;
    gr->has_stats = true;
    gr->h_refs = dg->h_cellref;

    Timing &tm = ctx->timing;
    StageTimer kt(ctx, 0, &tm.kernel_ms);
    StageTimer mt(ctx, 2, &tm.main_kernel_ms);

    const int64_t ntasks = nsrc * 8;
    // temporaries live in the context's grow-only workspace (no allocation after the first call of a given size)
    WsBuf<uint32_t> cnt(ctx->ws, "mk_cnt"), fillcnt(ctx->ws, "mk_fillcnt");
    WsBuf<uint8_t> ghostflag(ctx->ws, "mk_ghostflag"), bigflag(ctx->ws, "mk_bigflag");
    WsBuf<int64_t> overflow_list(ctx->ws, "mk_overflow");
    WsBuf<unsigned long long> n_overflow(ctx->ws, "mk_noverflow");
    WsBuf<int> error_flag(ctx->ws, "mk_error");
    WsBuf<uint64_t> row_tot(ctx->ws, "mk_rowtot"), row_off(ctx->ws, "mk_rowoff");
    VGA_TRY(cnt.alloc_zero((size_t)nsrc * 8 + 8, st));
    VGA_TRY(fillcnt.alloc_zero((size_t)nsrc * 4 + 4, st));
    VGA_TRY(ghostflag.alloc_zero(4, st));
    VGA_TRY(bigflag.alloc_zero((size_t)ntasks + 1, st));
    VGA_TRY(overflow_list.alloc((size_t)ntasks + 1));
    VGA_TRY(n_overflow.alloc_zero(1, st));
    VGA_TRY(error_flag.alloc_zero(1, st));
    VGA_TRY(row_tot.alloc((size_t)nsrc + 1));
    VGA_TRY(row_off.alloc((size_t)nsrc + 1));
    VGA_TRY(gr->rowptr.alloc((size_t)nsrc + 1));
    VGA_TRY(gr->connectivity.alloc((size_t)nsrc + 1));
    VGA_TRY(gr->sum_d.alloc((size_t)nsrc + 1));
    VGA_TRY(gr->sum_d2.alloc((size_t)nsrc + 1));
    VGA_TRY(gr->far_dist.alloc((size_t)nsrc * 32 + 32));
    VGA_TRY(gr->bin_count.alloc((size_t)nsrc * 32 + 32));
    VGA_TRY(gr->gridconn.alloc((size_t)nsrc + 1));

    const int gcap = (int)ctx->opt.sieve_gcap;
    const int bcap = next_pow2((int)ctx->opt.sieve_bcap);
    const int big_gcap = (int)ctx->opt.sieve_big_gcap;
    const int big_bcap = next_pow2((int)ctx->opt.sieve_big_bcap);
    const size_t smem_bytes = per_warp_bytes(gcap, bcap) * WARPS_PER_BLOCK;
    if (smem_bytes > ctx->smem_optin) {
        set_error("vga_graph_build: sieve shared-memory slices exceed the device limit");
        return VGA_ERR_INVALID;
    }
    VGA_CUDA(cudaFuncSetAttribute(k_sieve<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    VGA_CUDA(cudaFuncSetAttribute(k_sieve<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));

    SieveArgs a;
    a.g = g;
    a.src_begin = src_begin;
    a.ntasks = ntasks;
    a.tasklist = nullptr;
    a.task_begin = 0;
    a.gcap = gcap;
    a.bcap = bcap;
    a.gscratch = nullptr;
    a.bigflag = bigflag.p;
    a.cnt = cnt.p;
    a.fillcnt = fillcnt.p;
    a.ghostflag = ghostflag.p;
    a.overflow_list = overflow_list.p;
    a.n_overflow = n_overflow.p;
    a.error_flag = error_flag.p;
    a.row_off = nullptr;
    a.chunk_base = 0;
    a.e_ref = nullptr;
    a.e_bin = nullptr;
    a.order = nullptr;

    // big-capacity scratch (allocated on demand)
    WsBuf<unsigned char> big_scratch(ctx->ws, "mk_bigscratch");
    const int big_warps = ctx->sm_count * WARPS_PER_BLOCK;
    unsigned long long h_over = 0;
    int h_err = 0;

    const int64_t thread_cap = ctx->opt.sieve_thread_cap >= 0 ? ctx->opt.sieve_thread_cap : (N >= 200000 ? 1 : 0);
    kt.start();
    mt.start();
    // ---- pass 1: count
    if (ntasks > 0) {
        if (ctx->opt.sieve_mode == 1 && thread_cap == 1)
            k_sieve_thread<false, 16, 48><<<blocks_for(ntasks, TT), TT, 0, st>>>(a, nsrc);
        else if (ctx->opt.sieve_mode == 1)
            k_sieve_thread<false, 8, 24><<<blocks_for(ntasks, TT), TT, 0, st>>>(a, nsrc);
        else
            k_sieve<false><<<blocks_for(ntasks, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, smem_bytes, st>>>(a);
        tm.launches++;
        tm.main_launches++;
        VGA_CUDA(cudaGetLastError());
    }
    VGA_CUDA(cudaMemcpyAsync(&h_over, n_overflow.p, sizeof(h_over), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    if (std::getenv("VGA_DEBUG_TIMING"))
        fprintf(stderr, "[vga_graph_build] %lld of %lld (source, octant) tasks exceed the thread kernel's capacity\n", (long long)h_over,
                (long long)ntasks);
    if (h_over > 0) {
        VGA_TRY(big_scratch.alloc(per_warp_bytes(big_gcap, big_bcap) * (size_t)big_warps));
        SieveArgs b = a;
        b.ntasks = (int64_t)h_over;
        b.tasklist = overflow_list.p;
        b.gcap = big_gcap;
        b.bcap = big_bcap;
        b.gscratch = (Zone *)big_scratch.p;
        k_sieve<false><<<ctx->sm_count, WARPS_PER_BLOCK * 32, 0, st>>>(b);
        tm.launches++;
        tm.main_launches++;
        VGA_CUDA(cudaGetLastError());
    }
    mt.stop();
    VGA_CUDA(cudaMemcpyAsync(&h_err, error_flag.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    VGA_CUDA(cudaStreamSynchronize(st));
    if (h_err != 0) {
        set_error(h_err == VGA_ERR_CAPACITY ? "vga_graph_build: gap/block capacity exceeded even in big mode"
                                            : "vga_graph_build: NaN angular block (a wall passes exactly through a cell "
                                              "centre); the reference's result is unspecified for this input");
        return h_err;
    }

    // ---- ghosts: unfilled cells covered by a diagonal first..last run get ordinals N..; ALL unfilled
    // cells are numbered (statically, in x-major order) so that every shard and every rank agrees on
    // the vertex universe without communication, whether or not a fill-in actually occurred
    gr->ghosts = dg->cells - N;
    gr->h_refs.insert(gr->h_refs.end(), dg->h_ghostref.begin(), dg->h_ghostref.end());

    // ---- row offsets
    if (nsrc > 0) {
        k_row_totals<<<blocks_for(nsrc, 256), 256, 0, st>>>(cnt.p, fillcnt.p, nsrc, row_tot.p);
        tm.launches++;
    }
    VGA_CUDA(cudaMemsetAsync(row_tot.p + nsrc, 0, sizeof(uint64_t), st));
    {
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, row_tot.p, row_off.p, (int)(nsrc + 1), st);
        WsBuf<unsigned char> tmp(ctx->ws, "mk_scantmp");
        VGA_TRY(tmp.alloc(tb + 16));
        cub::DeviceScan::ExclusiveSum(tmp.p, tb, row_tot.p, row_off.p, (int)(nsrc + 1), st);
        tm.launches++;
        VGA_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<uint64_t> h_off((size_t)nsrc + 1);
    VGA_CUDA(cudaMemcpy(h_off.data(), row_off.p, sizeof(uint64_t) * (nsrc + 1), cudaMemcpyDeviceToHost));
    const uint64_t total = h_off[nsrc];
    gr->entries = (int64_t)total;
    VGA_TRY(gr->adj.alloc((size_t)total + 1));
    VGA_CUDA(cudaMemcpyAsync(gr->rowptr.p, row_off.p, sizeof(uint64_t) * (nsrc + 1), cudaMemcpyDeviceToDevice, st));

    // ---- pass 2 in chunks of sources: emit, node stats, keys, segmented sort into the final rows
    const uint64_t chunk_cap = (uint64_t)std::max<int64_t>(ctx->opt.build_chunk_entries, 1 << 20);
    uint64_t max_chunk = 0;
    {
        int64_t i = 0;
        while (i < nsrc) {
            int64_t j = i + 1;
            while (j < nsrc && h_off[j + 1] - h_off[i] <= chunk_cap) j++;
            max_chunk = std::max<uint64_t>(max_chunk, h_off[j] - h_off[i]);
            i = j;
        }
    }
    if (max_chunk >= ((uint64_t)1 << 31)) {
        set_error("vga_graph_build: a single row chunk exceeds 2^31 entries");
        return VGA_ERR_CAPACITY;
    }
    WsBuf<uint32_t> e_ref(ctx->ws, "mk_eref"), keys(ctx->ws, "mk_keys");
    WsBuf<uint8_t> e_bin(ctx->ws, "mk_ebin");
    WsBuf<uint64_t> seg_off(ctx->ws, "mk_segoff");
    WsBuf<unsigned char> sort_tmp(ctx->ws, "mk_sorttmp");
    WsBuf<uint32_t> tk_in(ctx->ws, "mk_tkin"), tk_out(ctx->ws, "mk_tkout"), tv_in(ctx->ws, "mk_tvin"), tv_out(ctx->ws, "mk_tvout");
    WsBuf<unsigned char> tsort_tmp(ctx->ws, "mk_tsorttmp");
    const bool sort_tasks = ctx->opt.sieve_mode == 1 && ctx->opt.sieve_sort_emit != 0 && nsrc * 8 < ((int64_t)1 << 31);
    if (sort_tasks) {
        VGA_TRY(tk_in.alloc((size_t)nsrc * 8 + 8));
        VGA_TRY(tk_out.alloc((size_t)nsrc * 8 + 8));
        VGA_TRY(tv_in.alloc((size_t)nsrc * 8 + 8));
        VGA_TRY(tv_out.alloc((size_t)nsrc * 8 + 8));
    }
    VGA_TRY(e_ref.alloc((size_t)max_chunk + 1));
    VGA_TRY(e_bin.alloc((size_t)max_chunk + 1));
    VGA_TRY(keys.alloc((size_t)max_chunk + 1));
    VGA_TRY(seg_off.alloc((size_t)nsrc + 1));

    int64_t i = 0;
    while (i < nsrc) {
        int64_t j = i + 1;
        while (j < nsrc && h_off[j + 1] - h_off[i] <= chunk_cap) j++;
        const int64_t ns = j - i;
        const uint64_t base = h_off[i], cnt_e = h_off[j] - h_off[i];
        if (ctx->cancel && ctx->cancel(ctx->user)) {
            set_error("cancelled");
            return VGA_ERR_CANCELLED;
        }
        SieveArgs e = a;
        e.src_begin = src_begin + i;
        e.ntasks = ns * 8;
        e.task_begin = 0;
        e.bigflag = bigflag.p + i * 8;
        e.cnt = cnt.p + i * 8;
        e.fillcnt = fillcnt.p + i * 4;
        e.row_off = row_off.p + i;
        e.chunk_base = base;
        e.e_ref = e_ref.p;
        e.e_bin = e_bin.p;
        mt.start();
        if (sort_tasks && ns * 8 > 32) {
            k_task_keys<<<blocks_for(ns * 8, 256), 256, 0, st>>>(e.cnt, ns, tk_in.p, tv_in.p);
            size_t tb = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tb, tk_in.p, tk_out.p, tv_in.p, tv_out.p, (int)(ns * 8), 0, 28, st);
            VGA_TRY(tsort_tmp.alloc(tb + 16));
            VGA_CUDA(cub::DeviceRadixSort::SortPairs(tsort_tmp.p, tb, tk_in.p, tk_out.p, tv_in.p, tv_out.p, (int)(ns * 8), 0, 28, st));
            tm.launches += 3;
            e.order = tv_out.p;
        }
        if (ctx->opt.sieve_mode == 1 && thread_cap == 1)
            k_sieve_thread<true, 16, 48><<<blocks_for(ns * 8, TT), TT, 0, st>>>(e, ns);
        else if (ctx->opt.sieve_mode == 1)
            k_sieve_thread<true, 8, 24><<<blocks_for(ns * 8, TT), TT, 0, st>>>(e, ns);
        else
            k_sieve<true><<<blocks_for(ns * 8, WARPS_PER_BLOCK), WARPS_PER_BLOCK * 32, smem_bytes, st>>>(e);
        tm.launches++;
        tm.main_launches++;
        VGA_CUDA(cudaGetLastError());
        if (h_over > 0) {
            // big tasks of this chunk: the overflow list holds global task ids; filter on host
            // (rare path) -- re-run them with global scratch
            std::vector<int64_t> all((size_t)h_over);
            VGA_CUDA(cudaMemcpyAsync(all.data(), overflow_list.p, sizeof(int64_t) * h_over, cudaMemcpyDeviceToHost, st));
            VGA_CUDA(cudaStreamSynchronize(st));
            std::vector<int64_t> mine;
            for (int64_t t : all)
                if (t >= i * 8 && t < j * 8) mine.push_back(t - i * 8);
            if (!mine.empty()) {
                DevBuf<int64_t> dl;
                VGA_TRY(dl.alloc(mine.size()));
                VGA_CUDA(cudaMemcpyAsync(dl.p, mine.data(), sizeof(int64_t) * mine.size(), cudaMemcpyHostToDevice, st));
                SieveArgs b = e;
                b.ntasks = (int64_t)mine.size();
                b.tasklist = dl.p;
                b.gcap = big_gcap;
                b.bcap = big_bcap;
                b.gscratch = (Zone *)big_scratch.p;
                k_sieve<true><<<ctx->sm_count, WARPS_PER_BLOCK * 32, 0, st>>>(b);
                tm.launches++;
                tm.main_launches++;
                VGA_CUDA(cudaGetLastError());
                VGA_CUDA(cudaStreamSynchronize(st));
            }
        }
        mt.stop();
        k_node_stats<<<blocks_for(ns * 32, NS_TPB), NS_TPB, 0, st>>>(g, src_begin, i, ns, row_off.p, base, cnt.p, e_ref.p, e_bin.p,
                                                         gr->connectivity.p, gr->sum_d.p, gr->sum_d2.p,
                                                         gr->far_dist.p, gr->bin_count.p, gr->gridconn.p);
        tm.launches++;
        VGA_CUDA(cudaGetLastError());
        if (cnt_e > 0) {
            k_make_keys<<<blocks_for((int64_t)cnt_e, 256), 256, 0, st>>>(g, e_ref.p, e_bin.p, cnt_e, keys.p);
            tm.launches++;
            k_rebase<<<blocks_for(ns + 1, 256), 256, 0, st>>>(row_off.p + i, ns + 1, base, seg_off.p);
            tm.launches++;
            const uint32_t total_v = (uint32_t)dg->cells;  // filled cells + ghosts
            const size_t rs_bytes = rank_sort_smem(total_v);
            if (ctx->opt.build_sort != 0 && rs_bytes <= ctx->smem_optin) {
                VGA_CUDA(cudaFuncSetAttribute(k_rank_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_bytes));
                const unsigned blocks = (unsigned)std::min<int64_t>(ns, (int64_t)ctx->sm_count * 4);  // one CTA per SM fits; each clears its bitmap once
                k_rank_sort<<<blocks, RS_TPB, rs_bytes, st>>>(keys.p, seg_off.p, ns, total_v, gr->adj.p + base);
                tm.launches++;
            } else {
                size_t tb = 0;
                cub::DeviceSegmentedSort::SortKeys(nullptr, tb, keys.p, gr->adj.p + base, (int)cnt_e, (int)ns, seg_off.p,
                                                   seg_off.p + 1, st);
                VGA_TRY(sort_tmp.alloc(tb + 16));
                VGA_CUDA(cub::DeviceSegmentedSort::SortKeys(sort_tmp.p, tb, keys.p, gr->adj.p + base, (int)cnt_e, (int)ns,
                                                            seg_off.p, seg_off.p + 1, st));
                tm.launches += 3;
            }
        }
        VGA_CUDA(cudaGetLastError());
        if (ctx->progress) ctx->progress(ctx->user, j, nsrc);
        i = j;
    }
    kt.stop();
    VGA_CUDA(cudaStreamSynchronize(st));
    VGA_CUDA(cudaGetLastError());
    // algorithmic bytes of construction: 8 B per accepted edge + 40 B per wall segment
    tm.algo_bytes = 8.0 * (double)total + 40.0 * (double)dg->nseg;
    *out = gr.release();
    return VGA_OK;
}

}  // namespace vga
