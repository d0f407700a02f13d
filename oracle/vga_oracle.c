/* TEST INFRASTRUCTURE ONLY -- see vga_oracle.h for scope, citations and the pinning statement.
 * Compile with -ffp-contract=off (oracle/Makefile): the reference is built without FMA. */
#include "vga_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ST_FILLED 0x0002 /* salalib/point.h:32 */
#define ST_CONTEXTFILLED 0x0008 /* semi-fill (GUI "context fill"): FILLED | CONTEXTFILLED, pointdata.cpp:435-441 */

/* ---------------------------------------------------------------- geometry (genlib/p2dpoly) */

typedef struct {
    double blx, bly, trx, try_;
    int parity;
} Ln;

/* Line(const Point2f&, const Point2f&) normal form -- genlib/p2dpoly.cpp:291-336 */
static Ln ln_make(double ax, double ay, double bx, double by) {
    Ln l;
    if (ax == bx) {
        l.blx = ax;
        l.trx = bx;
        l.parity = 1;
        if (ay <= by) {
            l.bly = ay;
            l.try_ = by;
        } else {
            l.bly = by;
            l.try_ = ay;
        }
    } else if (ax < bx) {
        l.blx = ax;
        l.trx = bx;
        if (ay <= by) {
            l.bly = ay;
            l.try_ = by;
            l.parity = 1;
        } else {
            l.bly = by;
            l.try_ = ay;
            l.parity = 0;
        }
    } else {
        l.blx = bx;
        l.trx = ax;
        if (by <= ay) {
            l.bly = by;
            l.try_ = ay;
            l.parity = 1;
        } else {
            l.bly = ay;
            l.try_ = by;
            l.parity = 0;
        }
    }
    return l;
}

/* accessors -- genlib/p2dpoly.h:445-452 */
static inline double ln_ax(const Ln *l) { return l->blx; }
static inline double ln_bx(const Ln *l) { return l->trx; }
static inline double ln_ay(const Ln *l) { return l->parity ? l->bly : l->try_; }
static inline double ln_by(const Ln *l) { return l->parity ? l->try_ : l->bly; }
static inline double ln_w(const Ln *l) { return fabs(l->trx - l->blx); }
static inline double ln_h(const Ln *l) { return fabs(l->try_ - l->bly); }

/* intersect_region -- genlib/p2dpoly.cpp:247-279 */
static int overlap1(double abl, double atr, double bbl, double btr, double tol) {
    if (abl > bbl) {
        if (btr >= abl - tol) return 1;
    } else {
        if (atr >= bbl - tol) return 1;
    }
    return 0;
}
static int isect_region(const Ln *a, const Ln *b, double tol) {
    return overlap1(a->blx, a->trx, b->blx, b->trx, tol) && overlap1(a->bly, a->try_, b->bly, b->try_, tol);
}

/* intersect_line -- genlib/p2dpoly.cpp:350-363 */
static int isect_line(const Ln *a, const Ln *b, double tol) {
    double aax = ln_ax(a), aay = ln_ay(a), abx = ln_bx(a), aby = ln_by(a);
    double bax = ln_ax(b), bay = ln_ay(b), bbx = ln_bx(b), bby = ln_by(b);
    if (((aay - aby) * (bax - aax) + (abx - aax) * (bay - aay)) * ((aay - aby) * (bbx - aax) + (abx - aax) * (bby - aay)) <=
            tol &&
        ((bay - bby) * (aax - bax) + (bbx - bax) * (aay - bay)) * ((bay - bby) * (abx - bax) + (bbx - bax) * (aby - bay)) <=
            tol)
        return 1;
    return 0;
}

/* Line::crop -- genlib/p2dpoly.cpp:626-667; region r = (rblx, rbly, rtrx, rtry) */
static int ln_crop(Ln *l, double rblx, double rbly, double rtrx, double rtry) {
    double sign = l->parity ? 1.0 : -1.0;
    if (l->trx >= rblx) {
        if (l->blx < rblx) {
            double d = sign * (ln_h(l) * (rblx - l->blx) / ln_w(l));
            if (l->parity)
                l->bly += d;
            else
                l->try_ += d;
            l->blx = rblx;
        }
        if (l->blx <= rtrx) {
            if (l->trx > rtrx) {
                double d = sign * ln_h(l) * (l->trx - rtrx) / ln_w(l);
                if (l->parity)
                    l->try_ -= d;
                else
                    l->bly -= d;
                l->trx = rtrx;
            }
            if (l->try_ >= rbly) {
                if (l->bly < rbly) {
                    double d = ln_w(l) * (rbly - l->bly) / ln_h(l);
                    if (l->parity)
                        l->blx += d;
                    else
                        l->trx -= d;
                    l->bly = rbly;
                }
                if (l->bly <= rtry) {
                    if (l->try_ > rtry) {
                        double d = ln_w(l) * (l->try_ - rtry) / ln_h(l);
                        if (l->parity)
                            l->trx -= d;
                        else
                            l->blx += d;
                        l->try_ = rtry;
                    }
                    return 1;
                }
            }
        }
    }
    return 0;
}

/* ---------------------------------------------------------------- sieve (salalib/sparksieve2) */

typedef struct {
    double start, end;
} Zone;

typedef struct {
    double cx, cy, maxdist;
    Zone *gaps;
    int ngaps, capgaps;
    Zone *blocks;
    int nblocks, capblocks;
} Sieve;

static void sieve_init(Sieve *s, double cx, double cy, double maxdist) {
    s->cx = cx;
    s->cy = cy;
    s->maxdist = maxdist;
    s->capgaps = 16;
    s->gaps = (Zone *)malloc(sizeof(Zone) * s->capgaps);
    s->ngaps = 1;
    s->gaps[0].start = 0.0;
    s->gaps[0].end = 1.0;
    s->capblocks = 16;
    s->blocks = (Zone *)malloc(sizeof(Zone) * s->capblocks);
    s->nblocks = 0;
}
static void sieve_free(Sieve *s) {
    free(s->gaps);
    free(s->blocks);
}

/* sparksieve2.cpp:143-173 -- IEEE division kept as is (x/0 = +-inf) */
static double tanify(const Sieve *s, double px, double py, int q) {
    switch (q) {
    case 0: return (py - s->cy) / (s->cx - px);
    case 1: return (py - s->cy) / (px - s->cx);
    case 2: return (s->cy - py) / (s->cx - px);
    case 3: return (s->cy - py) / (px - s->cx);
    case 4: return (s->cx - px) / (s->cy - py);
    case 5: return (px - s->cx) / (s->cy - py);
    case 6: return (s->cx - px) / (py - s->cy);
    case 7: return (px - s->cx) / (py - s->cy);
    }
    return -1.0;
}

/* ordering of sparkZone2 -- sparksieve2.h:64-79: start ascending, end descending */
static int zone_cmp(const void *pa, const void *pb) {
    const Zone *a = (const Zone *)pa, *b = (const Zone *)pb;
    if (a->start == b->start) {
        if (a->end > b->end) return -1;
        if (a->end < b->end) return 1;
        return 0;
    }
    return a->start < b->start ? -1 : 1;
}

/* sparksieve2.cpp:67-87 (sort + unique after appending) */
static void sieve_block(Sieve *s, const Ln *lines, int n, int q) {
    for (int i = 0; i < n; i++) {
        const Ln *l = &lines[i];
        double a = tanify(s, ln_ax(l), ln_ay(l), q);
        double b = tanify(s, ln_bx(l), ln_by(l), q);
        Zone z;
        if (a < b) {
            z.start = a - 1e-10;
            z.end = b + 1e-10;
        } else {
            z.start = b - 1e-10;
            z.end = a + 1e-10;
        }
        if (s->nblocks == s->capblocks) {
            s->capblocks *= 2;
            s->blocks = (Zone *)realloc(s->blocks, sizeof(Zone) * s->capblocks);
        }
        s->blocks[s->nblocks++] = z;
    }
    if (n == 0) return; /* sorting an already sorted+unique list is the identity */
    qsort(s->blocks, s->nblocks, sizeof(Zone), zone_cmp);
    int w = 0;
    for (int i = 0; i < s->nblocks; i++) {
        if (w == 0 || s->blocks[i].start != s->blocks[w - 1].start || s->blocks[i].end != s->blocks[w - 1].end)
            s->blocks[w++] = s->blocks[i];
    }
    s->nblocks = w;
}

/* sparksieve2.cpp:89-132 */
static void sieve_collectgarbage(Sieve *s) {
    int gi = 0, bi = 0;
    while (bi < s->nblocks && gi < s->ngaps) {
        Zone B = s->blocks[bi];
        Zone *G = &s->gaps[gi];
        if (B.end < G->start) {
            bi++;
            continue;
        }
        int create = 1;
        if (B.start <= G->start) {
            create = 0;
            if (B.end > G->start) G->start = B.end;
        }
        if (B.end >= G->end) {
            create = 0;
            if (B.start < G->end) G->end = B.start;
        }
        if (G->end <= G->start + 1e-10) {
            memmove(&s->gaps[gi], &s->gaps[gi + 1], sizeof(Zone) * (s->ngaps - gi - 1));
            s->ngaps--;
            continue;
        } else if (B.end > G->end) {
            gi++;
            continue;
        } else if (create) {
            if (s->ngaps == s->capgaps) {
                s->capgaps *= 2;
                s->gaps = (Zone *)realloc(s->gaps, sizeof(Zone) * s->capgaps);
                G = &s->gaps[gi];
            }
            memmove(&s->gaps[gi + 1], &s->gaps[gi], sizeof(Zone) * (s->ngaps - gi));
            s->ngaps++;
            s->gaps[gi].start = s->gaps[gi + 1].start;
            s->gaps[gi].end = B.start;
            s->gaps[gi + 1].start = B.end;
            gi++;
        }
        bi++;
    }
    s->nblocks = 0;
}

/* sparksieve2.cpp:45-63 */
static int sieve_testblock(const Sieve *s, double px, double py, const Ln *lines, int n, double tol) {
    Ln l = ln_make(s->cx, s->cy, px, py);
    if (s->maxdist != -1.0) {
        double len = sqrt((l.trx - l.blx) * (l.trx - l.blx) + (l.try_ - l.bly) * (l.try_ - l.bly));
        if (len > s->maxdist) return 1;
    }
    for (int i = 0; i < n; i++) {
        if (isect_region(&l, &lines[i], tol) && isect_line(&l, &lines[i], tol)) return 1;
    }
    return 0;
}

int vgao_sieve_kat(double cx, double cy, int q, const double *segs, int nsegs, double *gaps, int cap) {
    Sieve s;
    sieve_init(&s, cx, cy, -1.0);
    Ln *ls = (Ln *)malloc(sizeof(Ln) * (nsegs > 0 ? nsegs : 1));
    for (int i = 0; i < nsegs; i++) ls[i] = ln_make(segs[4 * i], segs[4 * i + 1], segs[4 * i + 2], segs[4 * i + 3]);
    sieve_block(&s, ls, nsegs, q);
    sieve_collectgarbage(&s);
    int n = s.ngaps;
    for (int i = 0; i < n && i < cap; i++) {
        gaps[2 * i] = s.gaps[i].start;
        gaps[2 * i + 1] = s.gaps[i].end;
    }
    free(ls);
    sieve_free(&s);
    return n;
}

/* ---------------------------------------------------------------- whichbin (pointdata.h:432-520) */

static int whichbin(double gx, double gy) {
    int bin = 0;
    double ratio;
    if (fabs(gy) > fabs(gx)) bin = 1;
    if (bin == 0) {
        ratio = fabs(gy) / fabs(gx);
        if (gx > 0.0)
            bin = (gy >= 0.0) ? 0 : -32;
        else
            bin = (gy >= 0.0) ? -16 : 16;
    } else {
        ratio = fabs(gx) / fabs(gy);
        if (gy > 0.0)
            bin = (gx >= 0.0) ? -8 : 8;
        else
            bin = (gx >= 0.0) ? 24 : -24;
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

/* ---------------------------------------------------------------- graph container */

typedef struct {
    int32_t *ref;
    uint8_t *bin;
    int64_t n, cap;
} EVec;

static void ev_push(EVec *v, int32_t ref, uint8_t bin) {
    if (v->n == v->cap) {
        v->cap = v->cap ? v->cap * 2 : 1024;
        v->ref = (int32_t *)realloc(v->ref, sizeof(int32_t) * v->cap);
        v->bin = (uint8_t *)realloc(v->bin, v->cap);
    }
    v->ref[v->n] = ref;
    v->bin[v->n] = bin;
    v->n++;
}

struct vgao_graph {
    int32_t cols, rows;
    int64_t n;         /* filled cells */
    int32_t *cellref;  /* [n] packed PixelRef */
    int32_t *ord;      /* [cols*rows] ordinal or -1 */
    const uint16_t *state_copy;
    uint16_t *state;
    int64_t src_begin, src_end;
    uint64_t *acc_ptr; /* [n+1] */
    EVec acc;
    uint64_t *it_ptr; /* [n+1] */
    EVec it;
    float *conn, *m1, *m2, *far;
    uint16_t *bincount;
    uint8_t *gridconn;
};

static inline int32_t pack_ref(int x, int y) { return (int32_t)(((uint32_t)x << 16) + ((uint32_t)y & 0xffff)); }
static inline int ref_x(int32_t r) { return (int)(int16_t)(r >> 16); }
static inline int ref_y(int32_t r) { return (int)(int16_t)(r & 0xffff); }

static vgao_graph *graph_alloc(const vgao_grid *g) {
    vgao_graph *gr = (vgao_graph *)calloc(1, sizeof(vgao_graph));
    gr->cols = g->cols;
    gr->rows = g->rows;
    int64_t cells = (int64_t)g->cols * g->rows;
    gr->ord = (int32_t *)malloc(sizeof(int32_t) * cells);
    gr->state = (uint16_t *)malloc(sizeof(uint16_t) * cells);
    memcpy(gr->state, g->state, sizeof(uint16_t) * cells);
    int64_t n = 0;
    for (int64_t i = 0; i < cells; i++) gr->ord[i] = (g->state[i] & ST_FILLED) ? (int32_t)n++ : -1;
    gr->n = n;
    gr->cellref = (int32_t *)malloc(sizeof(int32_t) * (n ? n : 1));
    for (int x = 0; x < g->cols; x++)
        for (int y = 0; y < g->rows; y++) {
            int32_t o = gr->ord[(int64_t)x * g->rows + y];
            if (o >= 0) gr->cellref[o] = pack_ref(x, y);
        }
    gr->acc_ptr = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
    gr->it_ptr = (uint64_t *)calloc(n + 1, sizeof(uint64_t));
    gr->conn = (float *)calloc(n ? n : 1, sizeof(float));
    gr->m1 = (float *)calloc(n ? n : 1, sizeof(float));
    gr->m2 = (float *)calloc(n ? n : 1, sizeof(float));
    gr->far = (float *)calloc((n ? n : 1) * 32, sizeof(float));
    gr->bincount = (uint16_t *)calloc((n ? n : 1) * 32, sizeof(uint16_t));
    gr->gridconn = (uint8_t *)calloc(n ? n : 1, 1);
    return gr;
}

void vgao_graph_free(vgao_graph *gr) {
    if (!gr) return;
    free(gr->cellref);
    free(gr->ord);
    free(gr->state);
    free(gr->acc_ptr);
    free(gr->acc.ref);
    free(gr->acc.bin);
    free(gr->it_ptr);
    free(gr->it.ref);
    free(gr->it.bin);
    free(gr->conn);
    free(gr->m1);
    free(gr->m2);
    free(gr->far);
    free(gr->bincount);
    free(gr->gridconn);
    free(gr);
}

/* ---------------------------------------------------------------- makegraph */

static int cmp_h(const void *pa, const void *pb) { /* PixelRefH: y then x -- ngraph.h:158-165 */
    int32_t a = *(const int32_t *)pa, b = *(const int32_t *)pb;
    int ay = ref_y(a), by = ref_y(b);
    if (ay != by) return ay < by ? -1 : 1;
    int ax = ref_x(a), bx = ref_x(b);
    return ax < bx ? -1 : (ax > bx ? 1 : 0);
}
static int cmp_v(const void *pa, const void *pb) { /* PixelRefV / PixelRef <: x then y */
    int32_t a = *(const int32_t *)pa, b = *(const int32_t *)pb;
    int ax = ref_x(a), bx = ref_x(b);
    if (ax != bx) return ax < bx ? -1 : 1;
    int ay = ref_y(a), by = ref_y(b);
    return ay < by ? -1 : (ay > by ? 1 : 0);
}

static void cell_lines(const vgao_grid *g, int x, int y, Ln **buf, int *cap, int *n) {
    int64_t c = (int64_t)x * g->rows + y;
    uint32_t a = g->line_off[c], b = g->line_off[c + 1];
    int cnt = (int)(b - a);
    if (cnt > *cap) {
        *cap = cnt * 2;
        *buf = (Ln *)realloc(*buf, sizeof(Ln) * *cap);
    }
    for (int i = 0; i < cnt; i++) {
        const double *p = g->lines + 5 * (int64_t)(a + i);
        (*buf)[i].blx = p[0];
        (*buf)[i].bly = p[1];
        (*buf)[i].trx = p[2];
        (*buf)[i].try_ = p[3];
        (*buf)[i].parity = p[4] != 0.0;
    }
    *n = cnt;
}

/* one source: sparkPixel2(curs, make=1) -- pointdata.cpp:1380-1510, with sieve2 :1512-1565 inlined */
static void spark_pixel(const vgao_grid *g, vgao_graph *gr, int cx, int cy, int64_t v) {
    double s = g->spacing;
    double c0x = g->bl_x + s * 1.0 * (double)cx, c0y = g->bl_y + s * 1.0 * (double)cy; /* depixelate */
    float far[32];
    for (int i = 0; i < 32; i++) far[i] = 0.0f;
    int nsize = 0;
    double total_dist = 0.0, total_dist_sqr = 0.0;
    EVec binlist[32];
    memset(binlist, 0, sizeof(binlist));
    Ln *lbuf = NULL;
    int lcap = 0, ln = 0;
    Ln *l0 = NULL;
    int l0cap = 0;
    int32_t *addlist = NULL;
    int addcap = 0;

    for (int q = 0; q < 8; q++) {
        Sieve sv;
        sieve_init(&sv, c0x, c0y, g->maxdist);
        double border = s * 1e-10;
        /* regionate(curs, 1e-10) -- pointdata.h:359-367 */
        double vblx = g->bl_x + s * ((double)cx - 0.5 - 1e-10), vbly = g->bl_y + s * ((double)cy - 0.5 - 1e-10);
        double vtrx = g->bl_x + s * ((double)cx + 0.5 + 1e-10), vtry = g->bl_y + s * ((double)cy + 0.5 + 1e-10);
        switch (q) { /* pointdata.cpp:1405-1438 */
        case 0: vtrx = c0x; vbly = c0y - border; break;
        case 6: vtrx = c0x + border; vbly = c0y; break;
        case 1: vblx = c0x; vbly = c0y - border; break;
        case 7: vblx = c0x - border; vbly = c0y; break;
        case 2: vtrx = c0x; vtry = c0y + border; break;
        case 4: vtrx = c0x + border; vtry = c0y; break;
        case 3: vblx = c0x; vtry = c0y + border; break;
        case 5: vblx = c0x - border; vtry = c0y; break;
        }
        cell_lines(g, cx, cy, &lbuf, &lcap, &ln);
        if (ln > l0cap) {
            l0cap = ln * 2;
            l0 = (Ln *)realloc(l0, sizeof(Ln) * l0cap);
        }
        int n0 = 0;
        for (int i = 0; i < ln; i++) {
            Ln l = lbuf[i];
            if (ln_crop(&l, vblx, vbly, vtrx, vtry)) l0[n0++] = l;
        }
        sieve_block(&sv, l0, n0, q);
        sieve_collectgarbage(&sv);

        for (int depth = 1; sv.ngaps > 0; depth++) {
            int nadd = 0;
            int hasgaps = 0;
            int firstind = 0;
            for (int gi = 0; gi < sv.ngaps; gi++) {
                double gs = sv.gaps[gi].start, ge = sv.gaps[gi].end;
                int lo = (int)ceil(gs * (depth - 0.5) - 0.5);
                int hi = (int)floor(ge * (depth + 0.5) + 0.5);
                for (int ind = lo; ind <= hi; ind++) {
                    if (ind < firstind) continue;
                    if (ind > depth) break;
                    firstind = ind;
                    int x = (q >= 4 ? ind : depth);
                    int y = (q >= 4 ? depth : ind);
                    int hx = cx + ((q % 2) ? x : -x);
                    int hy = cy + ((q <= 1 || q >= 6) ? y : -y);
                    if (hx >= 0 && hx < g->cols && hy >= 0 && hy < g->rows) {
                        hasgaps = 1;
                        int centregap = ((double)ind >= (gs * depth) && (double)ind <= (ge * depth));
                        int64_t hc = (int64_t)hx * g->rows + hy;
                        cell_lines(g, hx, hy, &lbuf, &lcap, &ln);
                        if (centregap && (g->state[hc] & ST_FILLED)) {
                            if ((ind != 0 || q == 0 || q == 1 || q == 5 || q == 6) && (ind != depth || q < 4)) {
                                double px = g->bl_x + s * 1.0 * (double)hx, py = g->bl_y + s * 1.0 * (double)hy;
                                if (!sieve_testblock(&sv, px, py, lbuf, ln, s * 1e-10)) {
                                    if (nadd == addcap) {
                                        addcap = addcap ? addcap * 2 : 256;
                                        addlist = (int32_t *)realloc(addlist, sizeof(int32_t) * addcap);
                                    }
                                    addlist[nadd++] = pack_ref(hx, hy);
                                }
                            }
                        }
                        sieve_block(&sv, lbuf, ln, q);
                    }
                }
            }
            sieve_collectgarbage(&sv);
            if (!hasgaps) break;
            for (int i = 0; i < nadd; i++) {
                int hx = ref_x(addlist[i]), hy = ref_y(addlist[i]);
                double px = g->bl_x + s * 1.0 * (double)hx, py = g->bl_y + s * 1.0 * (double)hy;
                int bin = whichbin(px - c0x, py - c0y);
                double dx = (double)(hx - cx), dy = (double)(hy - cy);
                double this_dist = sqrt(dx * dx + dy * dy) * s;
                if (this_dist > far[bin]) far[bin] = (float)this_dist;
                total_dist += this_dist;
                total_dist_sqr += this_dist * this_dist;
                nsize++;
                ev_push(&binlist[bin], addlist[i], (uint8_t)bin);
                ev_push(&gr->acc, addlist[i], (uint8_t)bin);
            }
        }
        sieve_free(&sv);
    }
    gr->acc_ptr[v + 1] = (uint64_t)gr->acc.n;

    /* Node::make / Bin::make -- ngraph.cpp:27-58, 234-304; iteration order :392-416 */
    for (int b = 0; b < 32; b++) {
        EVec *bl = &binlist[b];
        gr->far[v * 32 + b] = far[b];
        gr->bincount[v * 32 + b] = (uint16_t)bl->n;
        if (bl->n == 0) continue;
        if (b == 4 || b == 20 || b == 12 || b == 28) {
            int32_t st = bl->ref[0], en = bl->ref[0], back = bl->ref[bl->n - 1];
            if (ref_x(back) < ref_x(st)) st = back;
            if (ref_x(back) > ref_x(en)) en = back;
            int dy = (b == 4 || b == 20) ? 1 : -1;
            int x = ref_x(st), y = ref_y(st);
            for (; x <= ref_x(en); x++, y += dy) ev_push(&gr->it, pack_ref(x, y), (uint8_t)b);
        } else {
            int vertical = ((b > 4 && b < 12) || (b > 20 && b < 28));
            qsort(bl->ref, bl->n, sizeof(int32_t), vertical ? cmp_v : cmp_h);
            for (int64_t i = 0; i < bl->n; i++) {
                if (i > 0 && bl->ref[i] == bl->ref[i - 1]) continue; /* std::set */
                ev_push(&gr->it, bl->ref[i], (uint8_t)b);
            }
        }
        free(bl->ref);
        free(bl->bin);
    }
    gr->it_ptr[v + 1] = (uint64_t)gr->it.n;
    gr->conn[v] = (float)nsize;
    gr->m1[v] = (float)total_dist;
    gr->m2[v] = (float)total_dist_sqr;

    /* addGridConnections -- pointdata.cpp:1735-1768 */
    {
        static const int nx[8] = {1, 1, 0, -1, -1, -1, 0, 1};
        static const int ny[8] = {0, 1, 1, 1, 0, -1, -1, -1};
        uint8_t gc = 0;
        for (int i = 0; i < 8; i++) {
            int32_t want = pack_ref(cx + nx[i], cy + ny[i]);
            for (uint64_t e = gr->it_ptr[v]; e < gr->it_ptr[v + 1]; e++) {
                if (gr->it.bin[e] == i * 4 && gr->it.ref[e] == want) {
                    gc |= (uint8_t)(1 << i);
                    break;
                }
            }
        }
        gr->gridconn[v] = gc;
    }
    free(lbuf);
    free(l0);
    free(addlist);
}

vgao_graph *vgao_makegraph_range(const vgao_grid *g, int64_t src_begin, int64_t src_end) {
    vgao_graph *gr = graph_alloc(g);
    if (src_begin < 0) src_begin = 0;
    if (src_end > gr->n) src_end = gr->n;
    gr->src_begin = src_begin;
    gr->src_end = src_end;
    for (int64_t v = 0; v < gr->n; v++) {
        if (v >= src_begin && v < src_end) {
            spark_pixel(g, gr, ref_x(gr->cellref[v]), ref_y(gr->cellref[v]), v);
        } else {
            gr->acc_ptr[v + 1] = (uint64_t)gr->acc.n;
            gr->it_ptr[v + 1] = (uint64_t)gr->it.n;
        }
    }
    return gr;
}

vgao_graph *vgao_makegraph(const vgao_grid *g) { return vgao_makegraph_range(g, 0, INT64_MAX); }

vgao_graph *vgao_graph_from_edges(const vgao_grid *g, const uint64_t *rowptr, const int32_t *ref) {
    vgao_graph *gr = graph_alloc(g);
    gr->src_begin = 0;
    gr->src_end = gr->n;
    for (int64_t v = 0; v < gr->n; v++) {
        for (uint64_t e = rowptr[v]; e < rowptr[v + 1]; e++) ev_push(&gr->it, ref[e], 0);
        gr->it_ptr[v + 1] = (uint64_t)gr->it.n;
        gr->acc_ptr[v + 1] = 0;
    }
    return gr;
}

int64_t vgao_num_cells(const vgao_graph *gr) { return gr->n; }
int64_t vgao_num_acc(const vgao_graph *gr) { return gr->acc.n; }
int64_t vgao_num_iter(const vgao_graph *gr) { return gr->it.n; }
void vgao_cell_refs(const vgao_graph *gr, int32_t *ref) { memcpy(ref, gr->cellref, sizeof(int32_t) * gr->n); }
void vgao_acc_rows(const vgao_graph *gr, uint64_t *rowptr, int32_t *ref, uint8_t *bin) {
    memcpy(rowptr, gr->acc_ptr, sizeof(uint64_t) * (gr->n + 1));
    if (ref) memcpy(ref, gr->acc.ref, sizeof(int32_t) * gr->acc.n);
    if (bin) memcpy(bin, gr->acc.bin, gr->acc.n);
}
void vgao_iter_rows(const vgao_graph *gr, uint64_t *rowptr, int32_t *ref, uint8_t *bin) {
    memcpy(rowptr, gr->it_ptr, sizeof(uint64_t) * (gr->n + 1));
    if (ref) memcpy(ref, gr->it.ref, sizeof(int32_t) * gr->it.n);
    if (bin) memcpy(bin, gr->it.bin, gr->it.n);
}
void vgao_node_attrs(const vgao_graph *gr, float *connectivity, float *first_moment, float *second_moment,
                     float *far_bin_dists, uint16_t *bin_count, uint8_t *gridconn) {
    if (connectivity) memcpy(connectivity, gr->conn, sizeof(float) * gr->n);
    if (first_moment) memcpy(first_moment, gr->m1, sizeof(float) * gr->n);
    if (second_moment) memcpy(second_moment, gr->m2, sizeof(float) * gr->n);
    if (far_bin_dists) memcpy(far_bin_dists, gr->far, sizeof(float) * gr->n * 32);
    if (bin_count) memcpy(bin_count, gr->bincount, sizeof(uint16_t) * gr->n * 32);
    if (gridconn) memcpy(gridconn, gr->gridconn, gr->n);
}

/* ---------------------------------------------------------------- global BFS (vgavisualglobal.cpp) */

static inline int64_t cell_of_ref(const vgao_graph *gr, int32_t r) { return (int64_t)ref_x(r) * gr->rows + ref_y(r); }

/* a context-filled cell that is not "even" (PixelRef::iseven, pixelref.h:68-69) is skipped as a source and, when a
 * radius is set, counted but not expanded (vgavisualglobal.cpp:75, 108-110; vgavisuallocal.cpp:43) */
static inline int ctx_odd(const vgao_graph *gr, int32_t r) {
    int64_t c = (int64_t)ref_x(r) * gr->rows + ref_y(r);
    return (gr->state[c] & ST_CONTEXTFILLED) && !(ref_x(r) % 2 == 0 && ref_y(r) % 2 == 0);
}

int vgao_global(const vgao_graph *gr, int radius, int64_t src_begin, int64_t src_end, int32_t *total_nodes,
                int64_t *total_depth, int32_t *dist, int32_t maxl, int32_t *nlevels) {
    int64_t cells = (int64_t)gr->cols * gr->rows;
    /* seen[]: the only property of `miscs` that matters is ==0 / !=0; the `extents`
     * short-circuit is a pure optimisation */
    int32_t *seen = (int32_t *)malloc(sizeof(int32_t) * cells);
    int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * cells);
    int32_t *nxt = (int32_t *)malloc(sizeof(int32_t) * cells);
    for (int64_t i = 0; i < cells; i++) seen[i] = -1;
    int rc = 0;
    if (src_end > gr->n) src_end = gr->n;
    for (int64_t sidx = src_begin; sidx < src_end; sidx++) {
        int32_t stamp = (int32_t)(sidx - src_begin);
        int64_t out = sidx - src_begin;
        int64_t ncur = 0, nnxt = 0;
        if (ctx_odd(gr, gr->cellref[sidx])) { /* skipped source: the reference writes nothing, total_nodes -1 marks it */
            total_nodes[out] = -1;
            total_depth[out] = 0;
            if (dist) memset(dist + out * maxl, 0, sizeof(int32_t) * maxl);
            if (nlevels) nlevels[out] = 0;
            continue;
        }
        cur[ncur++] = gr->cellref[sidx];
        /* the source is pushed without being marked (misc==0) and finalised when popped; it can
         * never be re-pushed because its misc is ~0 before any other node is expanded */
        seen[cell_of_ref(gr, gr->cellref[sidx])] = stamp;
        int32_t tn = 0;
        int64_t td = 0;
        int level = 0;
        if (dist) memset(dist + out * maxl, 0, sizeof(int32_t) * maxl);
        int nl = 0;
        while (ncur > 0) {
            int32_t cnt = 0;
            nnxt = 0;
            for (int64_t i = 0; i < ncur; i++) {
                int64_t c = cell_of_ref(gr, cur[i]);
                if (!(gr->state[c] & ST_FILLED)) continue; /* p.filled() test :104 */
                td += level;
                tn += 1;
                cnt += 1;
                if (radius == -1 || (level < radius && !ctx_odd(gr, cur[i]))) {
                    int64_t u = gr->ord[c];
                    for (uint64_t e = gr->it_ptr[u]; e < gr->it_ptr[u + 1]; e++) {
                        int64_t wc = cell_of_ref(gr, gr->it.ref[e]);
                        if (seen[wc] != stamp) {
                            seen[wc] = stamp;
                            nxt[nnxt++] = gr->it.ref[e];
                        }
                    }
                }
            }
            if (level < maxl) {
                if (dist) dist[out * maxl + level] = cnt;
            } else if (cnt > 0) {
                rc = -1; /* maxl too small */
            }
            if (cnt > 0) nl = level + 1;
            int32_t *t = cur;
            cur = nxt;
            nxt = t;
            ncur = nnxt;
            level++;
        }
        total_nodes[out] = tn;
        total_depth[out] = td;
        if (nlevels) nlevels[out] = nl;
    }
    free(seen);
    free(cur);
    free(nxt);
    return rc;
}

/* genlib/pafmath.h:61-80 */
static const double M_1_LN2_ = 1.4426950408889634073599246810019; /* pafmath.h:44 */
static double paf_log2(double a) { return log(a) * M_1_LN2_; }
static double paf_dvalue(double k) { return 2.0 * (k * (paf_log2((k + 2.0) / 3.0) - 1.0) + 1.0) / ((k - 1.0) * (k - 2.0)); }
static double paf_pvalue(double k) { return 2.0 * (k - paf_log2(k) - 1.0) / ((k - 1.0) * (k - 2.0)); }
static double paf_teklinteg(double nodecount, double totaldepth) {
    return log(0.5 * (nodecount - 2.0)) / log((double)(totaldepth - nodecount + 1));
}

/* vgavisualglobal.cpp:131-193 */
void vgao_global_formulas(int64_t n, const int32_t *total_nodes, const int64_t *total_depth, const int32_t *dist,
                          int32_t maxl, const int32_t *nlevels, float *node_count, float *mean_depth,
                          float *integ_hh, float *integ_pv, float *integ_tk, float *entropy, float *rel_entropy) {
    for (int64_t i = 0; i < n; i++) {
        int tn = total_nodes[i];
        int td = (int)total_depth[i];
        /* untouched cells keep the attribute table's default -1 */
        float f_md = -1.0f, f_hh = -1.0f, f_pv = -1.0f, f_tk = -1.0f, f_en = -1.0f, f_re = -1.0f;
        if (tn > 1) {
            double md = (double)td / (double)(tn - 1);
            f_md = (float)md;
            if (tn > 2 && md > 1.0) {
                double ra = 2.0 * (md - 1.0) / (double)(tn - 2);
                double rra_d = ra / paf_dvalue(tn);
                double rra_p = ra / paf_pvalue(tn);
                double tk = paf_teklinteg(tn, td);
                f_hh = (float)(1.0 / rra_d);
                f_pv = (float)(1.0 / rra_p);
                f_tk = (td - tn + 1 > 1) ? (float)tk : -1.0f;
            }
            double en = 0.0, re = 0.0, factorial = 1.0;
            int nl = nlevels ? nlevels[i] : maxl;
            for (int k = 1; k < nl; k++) {
                int dk = dist[i * maxl + k];
                if (dk > 0) {
                    double prob = (double)dk / (double)(tn - 1);
                    en -= prob * paf_log2(prob);
                    factorial *= (double)(k + 1);
                    double q = (pow(md, (double)k) / (double)factorial) * exp(-md);
                    re += (float)prob * paf_log2(prob / q);
                }
            }
            f_en = (float)en;
            f_re = (float)re;
        }
        if (node_count) node_count[i] = (float)tn;
        if (mean_depth) mean_depth[i] = f_md;
        if (integ_hh) integ_hh[i] = f_hh;
        if (integ_pv) integ_pv[i] = f_pv;
        if (integ_tk) integ_tk[i] = f_tk;
        if (entropy) entropy[i] = f_en;
        if (rel_entropy) rel_entropy[i] = f_re;
    }
}

/* ---------------------------------------------------------------- local (vgavisuallocal.cpp:23-117) */

int vgao_local(const vgao_graph *gr, int64_t src_begin, int64_t src_end, int64_t *cluster, int32_t *k,
               int32_t *total, float *control) {
    int64_t cells = (int64_t)gr->cols * gr->rows;
    int64_t *in_hood = (int64_t *)malloc(sizeof(int64_t) * cells);
    int64_t *in_total = (int64_t *)malloc(sizeof(int64_t) * cells);
    for (int64_t i = 0; i < cells; i++) in_hood[i] = in_total[i] = -1;
    if (src_end > gr->n) src_end = gr->n;
    int32_t *hood = NULL;
    int64_t hcap = 0;
    for (int64_t v = src_begin; v < src_end; v++) {
        int64_t out = v - src_begin;
        if (ctx_odd(gr, gr->cellref[v])) { /* skipped cell (vgavisuallocal.cpp:43): k = -1 marks it */
            cluster[out] = 0;
            k[out] = -1;
            total[out] = 0;
            control[out] = 0.0f;
            continue;
        }
        /* contents(): iterated pixels, de-duplicated */
        int64_t nh = 0;
        for (uint64_t e = gr->it_ptr[v]; e < gr->it_ptr[v + 1]; e++) {
            int64_t c = cell_of_ref(gr, gr->it.ref[e]);
            if (in_hood[c] != v) {
                in_hood[c] = v;
                if (nh == hcap) {
                    hcap = hcap ? hcap * 2 : 1024;
                    hood = (int32_t *)realloc(hood, sizeof(int32_t) * hcap);
                }
                hood[nh++] = gr->it.ref[e];
            }
        }
        qsort(hood, nh, sizeof(int32_t), cmp_v); /* std::sort with PixelRef operator< */
        int64_t cl = 0;
        float ctl = 0.0f;
        int64_t ntotal = 0;
        for (int64_t i = 0; i < nh; i++) {
            int64_t c = cell_of_ref(gr, hood[i]);
            if (!(gr->state[c] & ST_FILLED)) continue; /* filled() && hasNode() */
            int64_t u = gr->ord[c];
            int isz = 0, rsz = 0;
            for (uint64_t e = gr->it_ptr[u]; e < gr->it_ptr[u + 1]; e++) {
                int64_t wc = cell_of_ref(gr, gr->it.ref[e]);
                rsz++;
                if (in_hood[wc] == v) isz++;
                if (in_total[wc] != v) {
                    in_total[wc] = v;
                    ntotal++;
                }
            }
            ctl += 1.0f / (float)rsz;
            cl += isz;
        }
        cluster[out] = cl;
        k[out] = (int32_t)nh;
        total[out] = (int32_t)ntotal;
        control[out] = ctl;
    }
    free(in_hood);
    free(in_total);
    free(hood);
    return 0;
}

void vgao_local_formulas(int64_t n, const int64_t *cluster, const int32_t *k, const int32_t *total,
                         const float *control, float *clustering, float *control_out, float *controllability) {
    for (int64_t i = 0; i < n; i++) {
        if (k[i] > 1) {
            size_t kk = (size_t)k[i];
            clustering[i] = (float)((int)cluster[i] / (double)(kk * (kk - 1.0)));
            control_out[i] = (float)control[i];
            controllability[i] = (float)((double)kk / (double)(size_t)total[i]);
        } else {
            clustering[i] = -1.0f;
            control_out[i] = -1.0f;
            controllability[i] = -1.0f;
        }
    }
}

/* ---------------------------------------------------------------- full-size checks over an ordinal CSR
 * The same level BFS (vgao_global) and local counts (vgao_local) over an adjacency handed over as CSR of vertex
 * ORDINALS (x-major index of the filled cell; ordinals >= n are unfilled "ghost" cells inside diagonal runs: never
 * counted or expanded, p.filled() test vgavisualglobal.cpp:104; members of neighbourhoods in vgavisuallocal.cpp:50-57
 * but without a node of their own).  No copy of the adjacency is made, so the full-size configurations (5.5e9 entries)
 * can be checked for sampled sources / cells.  Entries are `col[e] >> shift` (shift 6 = the library's packed form).
 * Single-threaded per call; the Python binding splits the independent sources / cells over threads. */
int vgao_global_csr(int64_t n, const uint64_t *rowptr, const uint32_t *col, int shift, int radius, const int64_t *src,
                    int64_t nsrc, int32_t *total_nodes, int64_t *total_depth, int32_t *dist, int32_t maxl) {
    int rc = 0;
    {
        uint8_t *seen = (uint8_t *)malloc((size_t)n);
        uint32_t *cur = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        uint32_t *nxt = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)n);
        for (int64_t s = 0; s < nsrc; s++) {
            memset(seen, 0, (size_t)n);
            int64_t ncur = 0, nnxt = 0;
            cur[ncur++] = (uint32_t)src[s];
            seen[src[s]] = 1;
            int32_t tn = 0;
            int64_t td = 0;
            int level = 0;
            if (dist) memset(dist + s * maxl, 0, sizeof(int32_t) * maxl);
            while (ncur > 0) {
                nnxt = 0;
                for (int64_t i = 0; i < ncur; i++) {
                    const uint32_t u = cur[i];
                    td += level;
                    tn += 1;
                    if (radius == -1 || level < radius) {
                        for (uint64_t e = rowptr[u]; e < rowptr[u + 1]; e++) {
                            const uint32_t w = col[e] >> shift;
                            if (w < (uint32_t)n && !seen[w]) {
                                seen[w] = 1;
                                nxt[nnxt++] = w;
                            }
                        }
                    }
                }
                if (level < maxl) {
                    if (dist) dist[s * maxl + level] = (int32_t)ncur;
                } else {
                    rc = -1;
                }
                uint32_t *t = cur;
                cur = nxt;
                nxt = t;
                ncur = nnxt;
                level++;
            }
            total_nodes[s] = tn;
            total_depth[s] = td;
        }
        free(seen);
        free(cur);
        free(nxt);
    }
    return rc;
}

/* rows must be sorted by ordinal (PixelRef order with ghosts last is NOT PixelRef order: `ref` gives the packed
 * PixelRef of every vertex, n cells then the ghosts, so that the float32 control sum runs in the reference's
 * sorted-PixelRef order, vgavisuallocal.cpp:50-52) */
static int cmp_refpair(const void *pa, const void *pb) {
    const int64_t a = *(const int64_t *)pa, b = *(const int64_t *)pb; /* (unsigned ref << 32) | ordinal: x then y */
    return (a > b) - (a < b);
}
int vgao_local_csr(int64_t n, int64_t nv, const uint64_t *rowptr, const uint32_t *col, int shift, const int32_t *ref,
                   const int64_t *cells, int64_t ncells, int64_t *cluster, int32_t *k, int32_t *total, float *control) {
    {
        int64_t *in_hood = (int64_t *)malloc(sizeof(int64_t) * (size_t)nv);
        int64_t *in_total = (int64_t *)malloc(sizeof(int64_t) * (size_t)nv);
        for (int64_t i = 0; i < nv; i++) in_hood[i] = in_total[i] = -1;
        int64_t *hood = NULL;
        int64_t hcap = 0;
        for (int64_t s = 0; s < ncells; s++) {
            const int64_t v = cells[s];
            int64_t nh = 0;
            for (uint64_t e = rowptr[v]; e < rowptr[v + 1]; e++) {
                const uint32_t w = col[e] >> shift;
                if (in_hood[w] != s) {
                    in_hood[w] = s;
                    if (nh == hcap) {
                        hcap = hcap ? hcap * 2 : 1024;
                        hood = (int64_t *)realloc(hood, sizeof(int64_t) * (size_t)hcap);
                    }
                    /* PixelRef operator< compares x then y as shorts; coordinates are non-negative here */
                    hood[nh++] = ((int64_t)(uint32_t)ref[w] << 32) | (int64_t)w;
                }
            }
            qsort(hood, (size_t)nh, sizeof(int64_t), cmp_refpair);
            int64_t cl = 0, ntotal = 0;
            float ctl = 0.0f;
            for (int64_t i = 0; i < nh; i++) {
                const uint32_t u = (uint32_t)(hood[i] & 0xffffffff);
                if (u >= (uint32_t)n) continue; /* not filled: no node */
                int64_t isz = 0, rsz = 0;
                for (uint64_t e = rowptr[u]; e < rowptr[u + 1]; e++) {
                    const uint32_t w = col[e] >> shift;
                    rsz++;
                    if (in_hood[w] == s) isz++;
                    if (in_total[w] != s) {
                        in_total[w] = s;
                        ntotal++;
                    }
                }
                ctl += 1.0f / (float)rsz;
                cl += isz;
            }
            cluster[s] = cl;
            k[s] = (int32_t)nh;
            total[s] = (int32_t)ntotal;
            control[s] = ctl;
        }
        free(in_hood);
        free(in_total);
        free(hood);
    }
    return 0;
}

/* ---------------------------------------------------------------- step depth (vgavisualglobaldepth.cpp:23-75)
 * BFS from a set of cells; depth[v] = level at which the filled cell v is first popped, -1 if never
 * (the column is reset to -1 by insertOrResetColumn).  No merges / context fill. */
int vgao_step_depth(const vgao_graph *gr, const int32_t *src, int64_t nsrc, int32_t *depth) {
    int64_t cells = (int64_t)gr->cols * gr->rows;
    uint8_t *seen = (uint8_t *)calloc(cells, 1);
    int32_t *cur = (int32_t *)malloc(sizeof(int32_t) * (cells + nsrc + 1));
    int32_t *nxt = (int32_t *)malloc(sizeof(int32_t) * (cells + nsrc + 1));
    for (int64_t v = 0; v < gr->n; v++) depth[v] = -1;
    int64_t ncur = 0;
    for (int64_t i = 0; i < nsrc; i++) cur[ncur++] = gr->cellref[src[i]];
    int level = 0;
    while (ncur > 0) {
        int64_t nnxt = 0;
        for (int64_t i = 0; i < ncur; i++) {
            int64_t c = cell_of_ref(gr, cur[i]);
            if (!(gr->state[c] & ST_FILLED)) continue;
            int64_t u = gr->ord[c];
            if (depth[u] != -1) continue; /* m_misc == ~0: already finalised */
            depth[u] = level;
            seen[c] = 1;
            if (level > 0 && ctx_odd(gr, cur[i])) continue; /* marked, not expanded (vgavisualglobaldepth.cpp:53) */
            for (uint64_t e = gr->it_ptr[u]; e < gr->it_ptr[u + 1]; e++) {
                int64_t wc = cell_of_ref(gr, gr->it.ref[e]);
                if (!seen[wc]) {
                    seen[wc] = 1;
                    nxt[nnxt++] = gr->it.ref[e];
                }
            }
        }
        int32_t *t = cur;
        cur = nxt;
        nxt = t;
        ncur = nnxt;
        level++;
    }
    free(seen);
    free(cur);
    free(nxt);
    return 0;
}

/* ---------------------------------------------------------------- metric / angular VGA
 * VGAMetric::run salalib/vgamodules/vgametric.cpp:25-136, VGAAngular::run vgaangular.cpp:22-133,
 * Node/Bin::extractMetric / extractAngular salalib/ngraph.cpp:67-87, 329-365, MetricTriple / AngularTriple ordering
 * salalib/pointdata.h:377-419, dist / angle salalib/pixelref.h:116-131, blockedAdjacent pointdata.cpp:1016-1066.
 *
 * A Dijkstra-like search per source over the ITERATED adjacency with a std::set<(key, pixel, lastpixel)> ordered by
 * (key, pixel) -- here a binary heap ordered by (key, pixel, insertion number): the set refuses a second element
 * with an equal (key, pixel), which then pops where the first one did; the heap pops the later duplicate right after
 * the first, when the pixel is already finalised (or is not a filled cell), so it is ignored either way.
 * Per-cell state of the reference (Point::m_misc / m_dist / m_cumangle) is kept for ALL cells, filled or not, because
 * the gaps of diagonal bins are relaxed and queued like any other pixel (they are never expanded).
 * Merge links: partner[v] = ordinal of the cell v is merged with (-1 none; NULL = no links): the partner of a finalised
 * cell is expanded from the same key with no last pixel and finalised without being counted.  radius -1.0 = unlimited. */
#define ST_BLOCKED 0x0004 /* salalib/point.h:33 */
typedef struct { float key; int32_t ref, last; uint64_t seq; } HeapE;
static inline int heap_less(const HeapE *a, const HeapE *b) {
    if (a->key != b->key) return a->key < b->key;
    int ax = ref_x(a->ref), bx = ref_x(b->ref);
    if (ax != bx) return ax < bx; /* PixelRef operator<, pixelref.h:95-98 */
    int ay = ref_y(a->ref), by = ref_y(b->ref);
    if (ay != by) return ay < by;
    return a->seq < b->seq;
}
typedef struct { HeapE *e; int64_t n, cap; uint64_t seq; } Heap;
static void heap_push(Heap *h, float key, int32_t ref, int32_t last) {
    if (h->n == h->cap) {
        h->cap = h->cap ? 2 * h->cap : 1024;
        h->e = (HeapE *)realloc(h->e, sizeof(HeapE) * h->cap);
    }
    HeapE x = {key, ref, last, h->seq++};
    int64_t i = h->n++;
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (!heap_less(&x, &h->e[p])) break;
        h->e[i] = h->e[p];
        i = p;
    }
    h->e[i] = x;
}
static HeapE heap_pop(Heap *h) {
    HeapE top = h->e[0], x = h->e[--h->n];
    int64_t i = 0;
    for (;;) {
        int64_t c = 2 * i + 1;
        if (c >= h->n) break;
        if (c + 1 < h->n && heap_less(&h->e[c + 1], &h->e[c])) c++;
        if (!heap_less(&h->e[c], &x)) break;
        h->e[i] = h->e[c];
        i = c;
    }
    if (h->n > 0) h->e[i] = x;
    return top;
}
static inline int cell_blocked(const vgao_graph *gr, int x, int y) {
    return x >= 0 && x < gr->cols && y >= 0 && y < gr->rows && (gr->state[(int64_t)x * gr->rows + y] & ST_BLOCKED) != 0;
}
/* Node::extractMetric / extractAngular expand a popped pixel only if its key is 0 or it is blocked or has a blocked cell
 * among its 8 neighbours (ngraph.cpp:71, 82) */
static int expands(const vgao_graph *gr, int32_t r) {
    int x = ref_x(r), y = ref_y(r);
    for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
            if (cell_blocked(gr, x + dx, y + dy)) return 1;
    return 0;
}
static inline double ref_dist(int32_t a, int32_t b) { /* pixelref.h:116-119: sqr() of ints is an int */
    int dx = ref_x(a) - ref_x(b), dy = ref_y(a) - ref_y(b);
    return sqrt((double)(dx * dx + dy * dy));
}
#ifndef M_PI
#define M_PI 3.14159265358979323846 /* glibc math.h value, the one the reference compiles against */
#endif
static const int32_t NO_PIXEL = -1; /* PixelRef(-1,-1) packs to -1 */
static inline double ref_angle(int32_t a, int32_t b, int32_t c) { /* pixelref.h:121-131 */
    if (c == NO_PIXEL) return 0.0;
    int abx = ref_x(a) - ref_x(b), aby = ref_y(a) - ref_y(b), bcx = ref_x(b) - ref_x(c), bcy = ref_y(b) - ref_y(c);
    return acos((double)(abx * bcx + aby * bcy) /
                (sqrt((double)(abx * abx + aby * aby)) * sqrt((double)(bcx * bcx + bcy * bcy)) + 1e-12));
}

int vgao_metric(const vgao_graph *gr, const int32_t *partner, double spacing, double radius, int64_t src_begin, int64_t src_end,
                float *mspa, float *mspl, float *msld, float *count) {
    int64_t cells = (int64_t)gr->cols * gr->rows;
    uint8_t *misc = (uint8_t *)malloc(cells);
    float *mdist = (float *)malloc(sizeof(float) * cells), *cum = (float *)malloc(sizeof(float) * cells);
    Heap h = {0, 0, 0, 0};
    if (src_end > gr->n) src_end = gr->n;
    for (int64_t s = src_begin; s < src_end; s++) {
        int32_t curs = gr->cellref[s];
        for (int64_t i = 0; i < cells; i++) {
            misc[i] = 0;
            mdist[i] = -1.0f;
            cum[i] = 0.0f;
        }
        float euclid_depth = 0.0f, total_depth = 0.0f, total_angle = 0.0f;
        int total_nodes = 0;
        h.n = 0;
        h.seq = 0;
        heap_push(&h, 0.0f, curs, NO_PIXEL);
        while (h.n > 0) {
            HeapE here = heap_pop(&h);
            if (radius != -1.0 && (here.key * spacing) > radius) break;
            int64_t c = cell_of_ref(gr, here.ref);
            if ((gr->state[c] & ST_FILLED) && !misc[c]) {
                if (here.key == 0.0f || expands(gr, here.ref)) {
                    int64_t u = gr->ord[c];
                    for (uint64_t e = gr->it_ptr[u]; e < gr->it_ptr[u + 1]; e++) {
                        int32_t pix = gr->it.ref[e];
                        int64_t pc = cell_of_ref(gr, pix);
                        if (!misc[pc] && (mdist[pc] == -1.0 || (here.key + ref_dist(pix, here.ref) < mdist[pc]))) {
                            mdist[pc] = here.key + (float)ref_dist(pix, here.ref);
                            cum[pc] = cum[c] + (here.last == NO_PIXEL ? 0.0f : (float)(ref_angle(pix, here.ref, here.last) / (M_PI * 0.5)));
                            heap_push(&h, mdist[pc], pix, here.ref);
                        }
                    }
                }
                misc[c] = 1;
                if (partner && partner[gr->ord[c]] >= 0) { /* vgametric.cpp:96-104 */
                    int64_t u2 = partner[gr->ord[c]];
                    int32_t r2 = gr->cellref[u2];
                    int64_t c2 = cell_of_ref(gr, r2);
                    if (!misc[c2]) {
                        cum[c2] = cum[c];
                        if (here.key == 0.0f || expands(gr, r2)) {
                            for (uint64_t e = gr->it_ptr[u2]; e < gr->it_ptr[u2 + 1]; e++) {
                                int32_t pix = gr->it.ref[e];
                                int64_t pc = cell_of_ref(gr, pix);
                                if (!misc[pc] && (mdist[pc] == -1.0 || (here.key + ref_dist(pix, r2) < mdist[pc]))) {
                                    mdist[pc] = here.key + (float)ref_dist(pix, r2);
                                    cum[pc] = cum[c2] + 0.0f; /* lastpixel == NoPixel */
                                    heap_push(&h, mdist[pc], pix, r2);
                                }
                            }
                        }
                        misc[c2] = 1;
                    }
                }
                total_depth += (float)(here.key * spacing);
                total_angle += cum[c];
                euclid_depth += (float)(spacing * ref_dist(here.ref, curs));
                total_nodes += 1;
            }
        }
        int64_t o = s - src_begin;
        mspa[o] = (float)((double)total_angle / (double)total_nodes);
        mspl[o] = (float)((double)total_depth / (double)total_nodes);
        msld[o] = (float)((double)euclid_depth / (double)total_nodes);
        count[o] = (float)total_nodes;
    }
    free(h.e);
    free(misc);
    free(mdist);
    free(cum);
    return 0;
}

/* mean_depth is only written when total_nodes > 0 (always true: the source counts itself) */
int vgao_angular(const vgao_graph *gr, const int32_t *partner, double radius, int64_t src_begin, int64_t src_end,
                 float *mean_depth, float *total_depth, float *count) {
    int64_t cells = (int64_t)gr->cols * gr->rows;
    uint8_t *misc = (uint8_t *)malloc(cells);
    float *cum = (float *)malloc(sizeof(float) * cells);
    Heap h = {0, 0, 0, 0};
    if (src_end > gr->n) src_end = gr->n;
    for (int64_t s = src_begin; s < src_end; s++) {
        int32_t curs = gr->cellref[s];
        for (int64_t i = 0; i < cells; i++) {
            misc[i] = 0;
            cum[i] = -1.0f;
        }
        float total_angle = 0.0f;
        int total_nodes = 0;
        h.n = 0;
        h.seq = 0;
        heap_push(&h, 0.0f, curs, NO_PIXEL);
        cum[cell_of_ref(gr, curs)] = 0.0f;
        while (h.n > 0) {
            HeapE here = heap_pop(&h);
            if (radius != -1.0 && here.key > radius) break;
            int64_t c = cell_of_ref(gr, here.ref);
            if ((gr->state[c] & ST_FILLED) && !misc[c]) {
                if (here.key == 0.0f || expands(gr, here.ref)) {
                    int64_t u = gr->ord[c];
                    for (uint64_t e = gr->it_ptr[u]; e < gr->it_ptr[u + 1]; e++) {
                        int32_t pix = gr->it.ref[e];
                        int64_t pc = cell_of_ref(gr, pix);
                        if (!misc[pc]) {
                            float ang = (here.last == NO_PIXEL) ? 0.0f : (float)(ref_angle(pix, here.ref, here.last) / (M_PI * 0.5));
                            if (cum[pc] == -1.0 || here.key + ang < cum[pc]) {
                                cum[pc] = cum[c] + ang;
                                heap_push(&h, cum[pc], pix, here.ref);
                            }
                        }
                    }
                }
                misc[c] = 1;
                if (partner && partner[gr->ord[c]] >= 0) { /* vgaangular.cpp:91-99 */
                    int64_t u2 = partner[gr->ord[c]];
                    int32_t r2 = gr->cellref[u2];
                    int64_t c2 = cell_of_ref(gr, r2);
                    if (!misc[c2]) {
                        cum[c2] = cum[c];
                        if (here.key == 0.0f || expands(gr, r2)) {
                            for (uint64_t e = gr->it_ptr[u2]; e < gr->it_ptr[u2 + 1]; e++) {
                                int32_t pix = gr->it.ref[e];
                                int64_t pc = cell_of_ref(gr, pix);
                                if (!misc[pc]) {
                                    float ang = 0.0f; /* lastpixel == NoPixel */
                                    if (cum[pc] == -1.0 || here.key + ang < cum[pc]) {
                                        cum[pc] = cum[c2] + ang;
                                        heap_push(&h, cum[pc], pix, r2);
                                    }
                                }
                            }
                        }
                        misc[c2] = 1;
                    }
                }
                total_angle += cum[c];
                total_nodes += 1;
            }
        }
        int64_t o = s - src_begin;
        mean_depth[o] = total_nodes > 0 ? (float)((double)total_angle / (double)total_nodes) : -1.0f;
        total_depth[o] = total_angle;
        count[o] = (float)total_nodes;
    }
    free(h.e);
    free(misc);
    free(cum);
    return 0;
}
