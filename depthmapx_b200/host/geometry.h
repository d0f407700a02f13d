// Host-side geometry value types of the GPU-backed PointMap (depthmapx_b200/host/pointmap.h).
//
// These mirror the *behaviour* of the reference's genlib/p2dpoly.{h,cpp} value types that the
// pre-steps of the hot path need: Line's normal form (bbox + parity,
// p2dpoly.cpp:291-336), the tolerance predicates intersect_region / intersect_line
// (p2dpoly.cpp:247-279, 350-363) and Line::crop (p2dpoly.cpp:626-667).  They run once per wall
// segment / per fill step on the host (O(#segments) and O(#cells)); everything per (source, target)
// pair runs on the GPU.  Compile without FMA contraction (-ffp-contract=off): the reference is built
// for plain x86-64 and every product/sum below must round exactly as it does there.
#pragma once

#include <cmath>

namespace dmx {

struct Point2f {
    double x = 0.0, y = 0.0;
    Point2f() = default;
    Point2f(double ax, double ay) : x(ax), y(ay) {}
};

struct Region {
    Point2f bl, tr;
    Region() = default;
    Region(const Point2f &a, const Point2f &b) : bl(a), tr(b) {}
    double width() const { return std::fabs(tr.x - bl.x); }
    double height() const { return std::fabs(tr.y - bl.y); }
};

// A segment kept as its bounding box plus `parity` (true: rises left to right; vertical segments
// are parity true).  a = left end, b = right end.
struct Line : Region {
    bool parity = false;
    Line() = default;
    Line(const Point2f &p, const Point2f &q) {
        const bool q_left = q.x < p.x;
        const Point2f &l = q_left ? q : p;
        const Point2f &r = q_left ? p : q;
        bl.x = l.x;
        tr.x = r.x;
        if (p.x == q.x) {
            parity = true;
            bl.y = p.y <= q.y ? p.y : q.y;
            tr.y = p.y <= q.y ? q.y : p.y;
        } else if (l.y <= r.y) {
            parity = true;
            bl.y = l.y;
            tr.y = r.y;
        } else {
            parity = false;
            bl.y = r.y;
            tr.y = l.y;
        }
    }
    double ax() const { return bl.x; }
    double bx() const { return tr.x; }
    double ay() const { return parity ? bl.y : tr.y; }
    double by() const { return parity ? tr.y : bl.y; }
    double &ay_ref() { return parity ? bl.y : tr.y; }
    double &by_ref() { return parity ? tr.y : bl.y; }
    double sign() const { return parity ? 1.0 : -1.0; }
    Point2f start() const { return Point2f(ax(), ay()); }
    Point2f end() const { return Point2f(bx(), by()); }

    // Clip to r; false if the segment lies outside.  Order: left, right, bottom, top, each step with
    // the current width/height.
    bool crop(const Region &r) {
        if (!(bx() >= r.bl.x)) return false;
        if (ax() < r.bl.x) {
            ay_ref() += sign() * (height() * (r.bl.x - ax()) / width());
            bl.x = r.bl.x;
        }
        if (!(ax() <= r.tr.x)) return false;
        if (bx() > r.tr.x) {
            by_ref() -= sign() * height() * (bx() - r.tr.x) / width();
            tr.x = r.tr.x;
        }
        if (!(tr.y >= r.bl.y)) return false;
        if (bl.y < r.bl.y) {
            const double d = width() * (r.bl.y - bl.y) / height();
            if (parity)
                bl.x += d;
            else
                tr.x -= d;
            bl.y = r.bl.y;
        }
        if (!(bl.y <= r.tr.y)) return false;
        if (tr.y > r.tr.y) {
            const double d = width() * (tr.y - r.tr.y) / height();
            if (parity)
                tr.x -= d;
            else
                bl.x += d;
            tr.y = r.tr.y;
        }
        return true;
    }
};

inline bool overlap_1d(double a_lo, double a_hi, double b_lo, double b_hi, double tol) {
    return (a_lo > b_lo) ? (b_hi >= a_lo - tol) : (a_hi >= b_lo - tol);
}

inline bool regions_touch(const Region &a, const Region &b, double tol) {
    return overlap_1d(a.bl.x, a.tr.x, b.bl.x, b.tr.x, tol) && overlap_1d(a.bl.y, a.tr.y, b.bl.y, b.tr.y, tol);
}

// the two signed-area products of the segment pair (touching counts when <= tol)
inline void area_products(const Line &a, const Line &b, double &pa, double &pb) {
    pa = ((a.ay() - a.by()) * (b.ax() - a.ax()) + (a.bx() - a.ax()) * (b.ay() - a.ay())) *
         ((a.ay() - a.by()) * (b.bx() - a.ax()) + (a.bx() - a.ax()) * (b.by() - a.ay()));
    pb = ((b.ay() - b.by()) * (a.ax() - b.ax()) + (b.bx() - b.ax()) * (a.ay() - b.ay())) *
         ((b.ay() - b.by()) * (a.bx() - b.ax()) + (b.bx() - b.ax()) * (a.by() - b.ay()));
}

inline bool lines_cross(const Line &a, const Line &b, double tol) {
    double pa, pb;
    area_products(a, b, pa, pb);
    return pa <= tol && pb <= tol;
}

inline bool lines_cross_no_touch(const Line &a, const Line &b, double tol = 0.0) {
    double pa, pb;
    area_products(a, b, pa, pb);
    return pa < -tol && pb < -tol;
}

// the combined test used by fill and by the sieve: bounding boxes first, then the products
inline bool blocks(const Line &sight, const Line &wall, double tol) {
    return regions_touch(sight, wall, tol) && lines_cross(sight, wall, tol);
}

}  // namespace dmx
