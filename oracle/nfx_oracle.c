/*
 * nfx_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Plain-C restatement of the transect-flux hot path of pletzer/nemoflux.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library; the shipped package (nemoflux_b200/) never does.
 *
 * PARITY STATUS
 *   * K2 (edge-flux assembly, field.py:145-234) is pinned: tests/golden/ holds outputs produced by
 *     the reference's own Field.readField / Field.computeIntegratedFlux / geo.getArcLengthArray
 *     code run in the build container (tests/golden/make_golden.py).
 *   * K1/K3 restate the third-party C++ library `mint` (conda-forge python-mint>=1.24.4, the only
 *     pin is /root/reference/README.md:12).  Its source is NOT in /root/reference, so this part is a
 *     restatement of the published algorithm from memory: PARITY UNPINNED against mint itself.  It is
 *     anchored on the reference's call sites (field.py:44-49, field.py:102, fluxplot.py:56), on the
 *     known answers recorded in README.md:39,56,68 and pictures/ PNG title bars (360, 0.5, ~0 for closed loops)
 *     and on the exactness invariants (node-to-node paths give psi(B)-psi(A)).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off: no FMA contraction, so every + - * / is the
 * IEEE-754 operation written here and the CUDA path can be compared bit for bit).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_EPS (10.0 * DBL_EPSILON)
#define ORC_EPS100 (100.0 * ORC_EPS)
#define ORC_NEWTON_MAXIT 20
#define ORC_NEWTON_TOL 1.0e-15
/* candidate filter margin in degrees: a pure superset filter, acceptance is decided by collect_lambdas */
#define ORC_FILTER_MARGIN 1.0e-7

typedef struct {
    int64_t cell;   /* cell id (row-major j*nx+i in nemoflux, horizgrid.py:17-22) */
    int32_t seg;    /* polyline segment index 0..npts-2 */
    int32_t img;    /* x-period image -1,0,+1 (segment shifted by img*periodX) */
    double ta, tb;  /* sub-segment parameter range on the segment */
    double coeff;   /* duplicity coefficient */
    double xia[2], xib[2]; /* cell parametric coordinates of the sub-segment ends */
    double w[4];    /* weights of edges 0=S 1=E 2=N 3=W (diagram field.py:198-206) */
} orc_subseg;

/* ---------------------------------------------------------------------------------------------
 * mint vmtCellLocator::containsPoint restated: point is inside when the z component of
 * (p - v_i) x (p - v_{i+1}) is >= -tol for the 4 counter-clockwise edges.
 * ------------------------------------------------------------------------------------------- */
static int contains_point(const double vx[4], const double vy[4], double px, double py, double tol) {
    int inside = 1;
    for (int i0 = 0; i0 < 4; ++i0) {
        int i1 = (i0 + 1) & 3;
        double d0x = px - vx[i0];
        double d0y = py - vy[i0];
        double d1x = px - vx[i1];
        double d1y = py - vy[i1];
        double cross = d0x * d1y - d0y * d1x;
        if (cross < -tol) inside = 0;
    }
    return inside;
}

/* ---------------------------------------------------------------------------------------------
 * mint vmtCellLocator::collectIntersectionPoints + LineLineIntersector restated.
 * Segment a->b against one quad.  Returns the number of lambdas (sorted ascending) in lam[<=10].
 * ------------------------------------------------------------------------------------------- */
static int collect_lambdas(const double vx[4], const double vy[4],
                           double ax, double ay, double bx, double by, double lam[10]) {
    int n = 0;
    if (contains_point(vx, vy, ax, ay, ORC_EPS)) lam[n++] = 0.0;
    if (contains_point(vx, vy, bx, by, ORC_EPS)) lam[n++] = 1.0;
    const double m0 = bx - ax;
    const double m2 = by - ay;
    for (int i0 = 0; i0 < 4; ++i0) {
        int i1 = (i0 + 1) & 3;
        double q0x = vx[i0], q0y = vy[i0], q1x = vx[i1], q1y = vy[i1];
        double m1 = q0x - q1x;
        double m3 = q0y - q1y;
        double r0 = q0x - ax;
        double r1 = q0y - ay;
        double det = m0 * m3 - m1 * m2;
        double s0 = m3 * r0 - m1 * r1; /* lambda * det */
        double s1 = m0 * r1 - m2 * r0; /* mu * det     */
        if (fabs(det) > ORC_EPS) {
            double l = s0 / det;
            double mu = s1 / det;
            if (l >= -ORC_EPS100 && l <= 1.0 + ORC_EPS100 && mu >= -ORC_EPS100 && mu <= 1.0 + ORC_EPS100)
                lam[n++] = l;
        } else if (fabs(s0) < ORC_EPS && fabs(s1) < ORC_EPS) {
            /* the segment and the edge lie on the same line: keep the overlap's end parameters */
            double u2 = m0 * m0 + m2 * m2;
            double t0x = q1x - ax;
            double t0y = q1y - ay;
            double lA = (r0 * m0 + r1 * m2) / u2;
            double lB = (t0x * m0 + t0y * m2) / u2;
            double lmin = lA < lB ? lA : lB;
            double lmax = lA < lB ? lB : lA;
            if (!(lmin > 1.0 + ORC_EPS || lmax < -ORC_EPS)) {
                double la = lmin > 0.0 ? lmin : 0.0;
                double lb = lmax < 1.0 ? lmax : 1.0;
                if (fabs(lb - la) > ORC_EPS) {
                    lam[n++] = la;
                    lam[n++] = lb;
                }
            }
        }
    }
    /* insertion sort, ascending */
    for (int i = 1; i < n; ++i) {
        double x = lam[i];
        int j = i - 1;
        while (j >= 0 && lam[j] > x) { lam[j + 1] = lam[j]; --j; }
        lam[j + 1] = x;
    }
    return n;
}

/* ---------------------------------------------------------------------------------------------
 * Inverse of the bilinear map of a quad (what vtkQuad::EvaluatePosition does for mint): Newton
 * from (1/2,1/2), fixed operation order, at most ORC_NEWTON_MAXIT steps, stop when both updates
 * are below ORC_NEWTON_TOL.
 * ------------------------------------------------------------------------------------------- */
static void param_coords(const double vx[4], const double vy[4], double px, double py, double xi[2]) {
    const double bx = vx[1] - vx[0], by = vy[1] - vy[0];
    const double cx = vx[3] - vx[0], cy = vy[3] - vy[0];
    const double dx = (vx[0] - vx[1]) + (vx[2] - vx[3]);
    const double dy = (vy[0] - vy[1]) + (vy[2] - vy[3]);
    double s = 0.5, t = 0.5;
    for (int it = 0; it < ORC_NEWTON_MAXIT; ++it) {
        double st = s * t;
        double fx = (((vx[0] + bx * s) + cx * t) + dx * st) - px;
        double fy = (((vy[0] + by * s) + cy * t) + dy * st) - py;
        double j00 = bx + dx * t;
        double j01 = cx + dx * s;
        double j10 = by + dy * t;
        double j11 = cy + dy * s;
        double det = j00 * j11 - j01 * j10;
        double ds = (fy * j01 - fx * j11) / det;
        double dt = (fx * j10 - fy * j00) / det;
        s = s + ds;
        t = t + dt;
        if (fabs(ds) < ORC_NEWTON_TOL && fabs(dt) < ORC_NEWTON_TOL) break;
    }
    xi[0] = s;
    xi[1] = t;
}

/* conservative "segment misses box" test (margin-expanded); 1 = certainly no contact */
static int seg_box_reject(double ax, double ay, double bx, double by, const double box[4]) {
    const double mg = ORC_FILTER_MARGIN;
    double sxmin = ax < bx ? ax : bx, sxmax = ax < bx ? bx : ax;
    double symin = ay < by ? ay : by, symax = ay < by ? by : ay;
    if (sxmax < box[0] - mg || sxmin > box[1] + mg || symax < box[2] - mg || symin > box[3] + mg) return 1;
    double dx = bx - ax, dy = by - ay;
    double mt = mg * (fabs(dx) + fabs(dy)) + mg * mg;
    double c0 = dx * (box[2] - ay) - dy * (box[0] - ax);
    double c1 = dx * (box[2] - ay) - dy * (box[1] - ax);
    double c2 = dx * (box[3] - ay) - dy * (box[1] - ax);
    double c3 = dx * (box[3] - ay) - dy * (box[0] - ax);
    if (c0 > mt && c1 > mt && c2 > mt && c3 > mt) return 1;
    if (c0 < -mt && c1 < -mt && c2 < -mt && c3 < -mt) return 1;
    return 0;
}

static int subseg_cmp(const void* pa, const void* pb) {
    const orc_subseg* a = (const orc_subseg*)pa;
    const orc_subseg* b = (const orc_subseg*)pb;
    if (a->ta < b->ta) return -1;
    if (a->ta > b->ta) return 1;
    if (a->cell < b->cell) return -1;
    if (a->cell > b->cell) return 1;
    if (a->img < b->img) return -1;
    if (a->img > b->img) return 1;
    return 0;
}

typedef struct {
    int64_t ncell;
    const double* pts; /* borrowed (ncell,4,3), horizgrid.py:17-22 */
    int64_t nl1, nl2;  /* bbox hierarchy: 64 cells / 4096 cells per node */
    double* box1;      /* (nl1,4) xmin xmax ymin ymax */
    double* box2;
} orc_grid;

#define ORC_L1 64
#define ORC_L2 4096

void* orc_grid_new(const double* pts, int64_t ncell) {
    orc_grid* g = (orc_grid*)calloc(1, sizeof(orc_grid));
    g->ncell = ncell;
    g->pts = pts;
    g->nl1 = (ncell + ORC_L1 - 1) / ORC_L1;
    g->nl2 = (ncell + ORC_L2 - 1) / ORC_L2;
    g->box1 = (double*)malloc(sizeof(double) * 4 * (size_t)(g->nl1 > 0 ? g->nl1 : 1));
    g->box2 = (double*)malloc(sizeof(double) * 4 * (size_t)(g->nl2 > 0 ? g->nl2 : 1));
    for (int64_t n = 0; n < g->nl1; ++n) {
        double b[4] = {DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX};
        int64_t c1 = (n + 1) * ORC_L1 < ncell ? (n + 1) * ORC_L1 : ncell;
        for (int64_t c = n * ORC_L1; c < c1; ++c)
            for (int v = 0; v < 4; ++v) {
                double x = pts[(c * 4 + v) * 3 + 0], y = pts[(c * 4 + v) * 3 + 1];
                if (x < b[0]) b[0] = x;
                if (x > b[1]) b[1] = x;
                if (y < b[2]) b[2] = y;
                if (y > b[3]) b[3] = y;
            }
        memcpy(g->box1 + 4 * n, b, sizeof b);
    }
    for (int64_t n = 0; n < g->nl2; ++n) {
        double b[4] = {DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX};
        int64_t k1 = (n + 1) * (ORC_L2 / ORC_L1) < g->nl1 ? (n + 1) * (ORC_L2 / ORC_L1) : g->nl1;
        for (int64_t k = n * (ORC_L2 / ORC_L1); k < k1; ++k) {
            const double* s = g->box1 + 4 * k;
            if (s[0] < b[0]) b[0] = s[0];
            if (s[1] > b[1]) b[1] = s[1];
            if (s[2] < b[2]) b[2] = s[2];
            if (s[3] > b[3]) b[3] = s[3];
        }
        memcpy(g->box2 + 4 * n, b, sizeof b);
    }
    return g;
}

void orc_grid_del(void* h) {
    orc_grid* g = (orc_grid*)h;
    if (!g) return;
    free(g->box1);
    free(g->box2);
    free(g);
}

typedef struct { orc_subseg* v; int64_t n, cap; } subseg_vec;

static void vec_push(subseg_vec* a, const orc_subseg* s) {
    if (a->n == a->cap) {
        a->cap = a->cap ? 2 * a->cap : 256;
        a->v = (orc_subseg*)realloc(a->v, sizeof(orc_subseg) * (size_t)a->cap);
    }
    a->v[a->n++] = *s;
}

static void try_cell(const orc_grid* g, int64_t c, int seg, int img, double ax, double ay, double bx, double by,
                     subseg_vec* out) {
    double vx[4], vy[4];
    for (int v = 0; v < 4; ++v) {
        vx[v] = g->pts[(c * 4 + v) * 3 + 0];
        vy[v] = g->pts[(c * 4 + v) * 3 + 1];
    }
    double lam[10];
    int n = collect_lambdas(vx, vy, ax, ay, bx, by, lam);
    if (n < 2) return;
    double ta = lam[0], tb = lam[n - 1];
    if (fabs(tb - ta) <= ORC_EPS100) return; /* PolysegmentIter: zero-length sub-segments dropped */
    orc_subseg s;
    memset(&s, 0, sizeof s);
    s.cell = c;
    s.seg = seg;
    s.img = img;
    s.ta = ta;
    s.tb = tb;
    vec_push(out, &s);
}

/* ---------------------------------------------------------------------------------------------
 * mint PolylineIntegral::computeWeights restated (call site field.py:44-49: periodX=360,
 * counterclock=False, enableFolding=False).
 *   filter_mode 0: brute force over every cell (no candidate filter)
 *   filter_mode 1: two-level bounding-box filter (superset of the accepted cells)
 * Output: malloc'ed array of sub-segments ordered by (segment, ta, cell, image); caller frees with
 * orc_free.  Returns 0 on success.
 * ------------------------------------------------------------------------------------------- */
int orc_pli_compute_weights(void* hgrid, const double* xyz, int npts, double periodX, int counterclock,
                            int filter_mode, orc_subseg** out, int64_t* nout) {
    orc_grid* g = (orc_grid*)hgrid;
    subseg_vec all = {0, 0, 0};
    if (!g || npts < 0) return 1;
    for (int seg = 0; seg + 1 < npts; ++seg) {
        const double p0x = xyz[3 * seg + 0], p0y = xyz[3 * seg + 1];
        const double p1x = xyz[3 * seg + 3], p1y = xyz[3 * seg + 4];
        if (p0x == p1x && p0y == p1y) continue; /* zero-length segment: nothing to integrate */
        subseg_vec sv = {0, 0, 0};
        int img0 = periodX > 0.0 ? -1 : 0, img1 = periodX > 0.0 ? 1 : 0;
        for (int img = img0; img <= img1; ++img) {
            double shift = (double)img * periodX;
            double ax = p0x + shift, bx = p1x + shift, ay = p0y, by = p1y;
            if (filter_mode == 0) {
                for (int64_t c = 0; c < g->ncell; ++c) try_cell(g, c, seg, img, ax, ay, bx, by, &sv);
            } else {
                for (int64_t n2 = 0; n2 < g->nl2; ++n2) {
                    if (seg_box_reject(ax, ay, bx, by, g->box2 + 4 * n2)) continue;
                    int64_t k1 = (n2 + 1) * (ORC_L2 / ORC_L1) < g->nl1 ? (n2 + 1) * (ORC_L2 / ORC_L1) : g->nl1;
                    for (int64_t n1 = n2 * (ORC_L2 / ORC_L1); n1 < k1; ++n1) {
                        if (seg_box_reject(ax, ay, bx, by, g->box1 + 4 * n1)) continue;
                        int64_t c1 = (n1 + 1) * ORC_L1 < g->ncell ? (n1 + 1) * ORC_L1 : g->ncell;
                        for (int64_t c = n1 * ORC_L1; c < c1; ++c) {
                            double b[4] = {DBL_MAX, -DBL_MAX, DBL_MAX, -DBL_MAX};
                            for (int v = 0; v < 4; ++v) {
                                double x = g->pts[(c * 4 + v) * 3 + 0], y = g->pts[(c * 4 + v) * 3 + 1];
                                if (x < b[0]) b[0] = x;
                                if (x > b[1]) b[1] = x;
                                if (y < b[2]) b[2] = y;
                                if (y > b[3]) b[3] = y;
                            }
                            if (seg_box_reject(ax, ay, bx, by, b)) continue;
                            try_cell(g, c, seg, img, ax, ay, bx, by, &sv);
                        }
                    }
                }
            }
        }
        /* PolysegmentIter: order by ta (tie-break: cell id, image) */
        qsort(sv.v, (size_t)sv.n, sizeof(orc_subseg), subseg_cmp);
        /* duplicity coefficients: 1 - overlap with the next sub-segment / own length */
        for (int64_t i = 0; i < sv.n; ++i) {
            orc_subseg* s = &sv.v[i];
            s->coeff = 1.0;
            if (i + 1 < sv.n) {
                const orc_subseg* q = &sv.v[i + 1];
                double hi = s->tb < q->tb ? s->tb : q->tb;
                double lo = s->ta > q->ta ? s->ta : q->ta;
                double ov = hi - lo;
                if (ov < 0.0) ov = 0.0;
                s->coeff = 1.0 - ov / (s->tb - s->ta);
            }
            /* parametric coordinates of both ends, then the 4 edge weights */
            double vx[4], vy[4];
            for (int v = 0; v < 4; ++v) {
                vx[v] = g->pts[(s->cell * 4 + v) * 3 + 0];
                vy[v] = g->pts[(s->cell * 4 + v) * 3 + 1];
            }
            double shift = (double)s->img * periodX;
            double ax = p0x + shift, bx = p1x + shift;
            double ddx = bx - ax, ddy = p1y - p0y;
            double pax = ax + s->ta * ddx, pay = p0y + s->ta * ddy;
            double pbx = ax + s->tb * ddx, pby = p0y + s->tb * ddy;
            param_coords(vx, vy, pax, pay, s->xia);
            param_coords(vx, vy, pbx, pby, s->xib);
            double dxi0 = s->xib[0] - s->xia[0];
            double dxi1 = s->xib[1] - s->xia[1];
            double xm0 = 0.5 * (s->xia[0] + s->xib[0]);
            double xm1 = 0.5 * (s->xia[1] + s->xib[1]);
            /* edges 0 (S, 0->1) and 1 (E, 1->2) run along +xi0/+xi1 in both orientations; with
             * counterclock=False (field.py:48) so do 2 (N, 3->2) and 3 (W, 0->3); the
             * counter-clockwise orientation flips those two */
            double sgn = counterclock ? -1.0 : 1.0;
            s->w[0] = (dxi0 * (1.0 - xm1)) * s->coeff;
            s->w[1] = (dxi1 * xm0) * s->coeff;
            s->w[2] = sgn * ((dxi0 * xm1) * s->coeff);
            s->w[3] = sgn * ((dxi1 * (1.0 - xm0)) * s->coeff);
            vec_push(&all, s);
        }
        free(sv.v);
    }
    *out = all.v;
    *nout = all.n;
    return 0;
}

void orc_free(void* p) { free(p); }

/* ---------------------------------------------------------------------------------------------
 * mint PolylineIntegral::getIntegral(data, CELL_BY_CELL_DATA) restated (field.py:102,
 * fluxplot.py:56): sum of weight * data[cell*4+edge].
 *   order 0 ("list"): sequential over the emission list (sub-segment order, edges 0..3)
 *   order 1 ("map") : mint keeps weights in a std::map keyed (cellId, edgeIndex): contributions to
 *                     the same key are added in emission order, the integral then runs over the
 *                     keys in ascending order.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int64_t key; int64_t pos; double w; } map_item;
static int map_cmp(const void* pa, const void* pb) {
    const map_item* a = (const map_item*)pa;
    const map_item* b = (const map_item*)pb;
    if (a->key < b->key) return -1;
    if (a->key > b->key) return 1;
    if (a->pos < b->pos) return -1;
    if (a->pos > b->pos) return 1;
    return 0;
}

/* builds the merged map; returns number of unique keys; keys[i]=cell*4+edge ascending */
int64_t orc_pli_merge(const orc_subseg* s, int64_t n, int64_t* keys, double* wsum) {
    map_item* it = (map_item*)malloc(sizeof(map_item) * (size_t)(4 * n + 1));
    for (int64_t i = 0; i < n; ++i)
        for (int e = 0; e < 4; ++e) {
            it[4 * i + e].key = s[i].cell * 4 + e;
            it[4 * i + e].pos = 4 * i + e;
            it[4 * i + e].w = s[i].w[e];
        }
    qsort(it, (size_t)(4 * n), sizeof(map_item), map_cmp);
    int64_t m = 0;
    for (int64_t i = 0; i < 4 * n; ++i) {
        if (m > 0 && keys[m - 1] == it[i].key) {
            wsum[m - 1] = wsum[m - 1] + it[i].w;
        } else {
            keys[m] = it[i].key;
            wsum[m] = it[i].w;
            ++m;
        }
    }
    free(it);
    return m;
}

double orc_pli_get_integral(const orc_subseg* s, int64_t n, const double* data, int order) {
    double res = 0.0;
    if (order == 0) {
        for (int64_t i = 0; i < n; ++i)
            for (int e = 0; e < 4; ++e) res = res + s[i].w[e] * data[s[i].cell * 4 + e];
        return res;
    }
    int64_t* keys = (int64_t*)malloc(sizeof(int64_t) * (size_t)(4 * n + 1));
    double* ws = (double*)malloc(sizeof(double) * (size_t)(4 * n + 1));
    int64_t m = orc_pli_merge(s, n, keys, ws);
    for (int64_t i = 0; i < m; ++i) res = res + ws[i] * data[keys[i]];
    free(keys);
    free(ws);
    return res;
}

/* ---------------------------------------------------------------------------------------------
 * K2 restated sequentially (field.py:145-163 readField, field.py:183-228 computeIntegratedFlux):
 *   U[c] = sum_k thickness[k] * nan_to_zero(uo[k,c])   (k ascending, plain mul then add)
 *   eU = U*arc1 ; eV = -V*arc2 ; optional *6371000/1e6 ; scatter into iV (ncell,4) with the
 *   south/west neighbour copies, x-periodic west, row-0 south edges left at 0.
 * One time step.  u,v are (nz,ny,nx).  iV may be NULL.  Threads: OpenMP over cells when built
 * with -fopenmp (the sum order per cell does not depend on the thread count).
 * ------------------------------------------------------------------------------------------- */
void orc_edgeflux_step(const double* u, const double* v, const double* thickness, const double* arc1,
                       const double* arc2, int nz, int ny, int nx, int sverdrup, double fill, double* eU,
                       double* eV, double* iV) {
    const int64_t ncell = (int64_t)ny * nx;
    const double scale = 6371000.0 / 1.e6; /* field.py:12,226 */
    const int has_fill = !(fill != fill);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncell; ++c) {
        double su = 0.0, sv = 0.0;
        for (int k = 0; k < nz; ++k) {
            double a = u[(int64_t)k * ncell + c];
            double b = v[(int64_t)k * ncell + c];
            if (a != a || (has_fill && a == fill)) a = 0.0;
            if (b != b || (has_fill && b == fill)) b = 0.0;
            su = su + thickness[k] * a;
            sv = sv + thickness[k] * b;
        }
        double fu = su * arc1[c];
        double fv = (-sv) * arc2[c];
        if (sverdrup) {
            fu = fu * scale;
            fv = fv * scale;
        }
        eU[c] = fu;
        eV[c] = fv;
    }
    if (iV) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < ncell; ++c) {
            int64_t j = c / nx, i = c % nx;
            iV[4 * c + 1] = eU[c];
            iV[4 * c + 2] = eV[c];
            iV[4 * c + 0] = j >= 1 ? eV[c - nx] : 0.0;
            iV[4 * c + 3] = i >= 1 ? eU[c - 1] : eU[c + nx - 1];
        }
    }
}

/* per-column vertical scale factors e3u/e3v (nz,ncell) instead of thickness[k]; not in the
 * reference (field.py:51 has a 1-D thickness only) -- this defines the extension: missing e3 counts as 0 too */
void orc_edgeflux_step_e3(const double* u, const double* v, const double* e3u, const double* e3v, const double* arc1,
                          const double* arc2, int nz, int ny, int nx, int sverdrup, double fill, double* eU, double* eV) {
    const int64_t ncell = (int64_t)ny * nx;
    const double scale = 6371000.0 / 1.e6;
    const int has_fill = !(fill != fill);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncell; ++c) {
        double su = 0.0, sv = 0.0;
        for (int k = 0; k < nz; ++k) {
            double a = u[(int64_t)k * ncell + c], b = v[(int64_t)k * ncell + c];
            double ea = e3u[(int64_t)k * ncell + c], eb = e3v[(int64_t)k * ncell + c];
            if (a != a || (has_fill && a == fill)) a = 0.0;
            if (b != b || (has_fill && b == fill)) b = 0.0;
            if (ea != ea || (has_fill && ea == fill)) ea = 0.0;
            if (eb != eb || (has_fill && eb == fill)) eb = 0.0;
            su = su + ea * a;
            sv = sv + eb * b;
        }
        double fu = su * arc1[c];
        double fv = (-sv) * arc2[c];
        if (sverdrup) {
            fu = fu * scale;
            fv = fv * scale;
        }
        eU[c] = fu;
        eV[c] = fv;
    }
}

/* float32-storage variant (real NEMO files store uo/vo as float32; accumulation stays fp64) */
void orc_edgeflux_step_f32(const float* u, const float* v, const double* thickness, const double* arc1,
                           const double* arc2, int nz, int ny, int nx, int sverdrup, float fill, double* eU,
                           double* eV, double* iV) {
    const int64_t ncell = (int64_t)ny * nx;
    const double scale = 6371000.0 / 1.e6;
    const int has_fill = !(fill != fill);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncell; ++c) {
        double su = 0.0, sv = 0.0;
        for (int k = 0; k < nz; ++k) {
            float a = u[(int64_t)k * ncell + c];
            float b = v[(int64_t)k * ncell + c];
            if (a != a || (has_fill && a == fill)) a = 0.0f;
            if (b != b || (has_fill && b == fill)) b = 0.0f;
            su = su + thickness[k] * (double)a;
            sv = sv + thickness[k] * (double)b;
        }
        double fu = su * arc1[c];
        double fv = (-sv) * arc2[c];
        if (sverdrup) {
            fu = fu * scale;
            fv = fv * scale;
        }
        eU[c] = fu;
        eV[c] = fv;
    }
    if (iV) {
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < ncell; ++c) {
            int64_t j = c / nx, i = c % nx;
            iV[4 * c + 1] = eU[c];
            iV[4 * c + 2] = eV[c];
            iV[4 * c + 0] = j >= 1 ? eV[c - nx] : 0.0;
            iV[4 * c + 3] = i >= 1 ? eU[c - 1] : eU[c + nx - 1];
        }
    }
}

/* K3 restated against the compact [eU | eV] layout: flux index of (cell, edge) per
 * field.py:209-223; -1 = the never-written south edge of row 0 (stays 0, field.py:61,219). */
int64_t orc_flux_index(int64_t cell, int edge, int ny, int nx) {
    const int64_t ncell = (int64_t)ny * nx;
    int64_t j = cell / nx, i = cell % nx;
    switch (edge) {
        case 1: return cell;
        case 3: return i >= 1 ? cell - 1 : cell + nx - 1;
        case 2: return ncell + cell;
        default: return j >= 1 ? ncell + cell - nx : -1;
    }
}

/* ---------------------------------------------------------------------------------------------
 * mint VectorInterp.findPoints + getFaceVectors restated (call sites field.py:90-95,119-120; viz only).
 * findPoints: the containing cell of every point (lowest cell id among the cells whose 4 CCW edge tests pass
 * with tolerance tol; images of the point shifted by 0, -periodX, +periodX in that order) and its
 * parametric coordinates.  getFaceVectors(data (ncell,4), placement 0): with r_xi, r_eta the tangent vectors
 * of the bilinear map and J = (r_xi x r_eta).z,
 *     v = [ (d3 (1-xi) + d1 xi) r_xi  -  (d0 (1-eta) + d2 eta) r_eta ] / J
 * Sign convention pinned by pictures/simple.png (psi = x: arrows point south) and pictures/singular.png
 * (arrows point away from the singularity).
 * ------------------------------------------------------------------------------------------- */
int orc_vinterp_find_points(void* hgrid, const double* xyz, int64_t npts, double periodX, double tol, int64_t* cell,
                            double* xi) {
    orc_grid* g = (orc_grid*)hgrid;
    if (!g) return 1;
    const double shifts[3] = {0.0, -periodX, periodX};
    const int nimg = periodX > 0.0 ? 3 : 1;
    for (int64_t n = 0; n < npts; ++n) {
        cell[n] = -1;
        xi[2 * n] = xi[2 * n + 1] = 0.0;
        for (int im = 0; im < nimg && cell[n] < 0; ++im) {
            const double px = xyz[3 * n] + shifts[im], py = xyz[3 * n + 1];
            for (int64_t c = 0; c < g->ncell; ++c) {
                double vx[4], vy[4];
                for (int v = 0; v < 4; ++v) {
                    vx[v] = g->pts[(c * 4 + v) * 3 + 0];
                    vy[v] = g->pts[(c * 4 + v) * 3 + 1];
                }
                if (contains_point(vx, vy, px, py, tol)) {
                    cell[n] = c;
                    param_coords(vx, vy, px, py, xi + 2 * n);
                    break;
                }
            }
        }
    }
    return 0;
}

void orc_vinterp_face_vectors(void* hgrid, const int64_t* cell, const double* xi, int64_t npts, const double* data,
                              double* vec) {
    orc_grid* g = (orc_grid*)hgrid;
    for (int64_t n = 0; n < npts; ++n) {
        vec[3 * n] = vec[3 * n + 1] = vec[3 * n + 2] = 0.0;
        const int64_t c = cell[n];
        if (c < 0) continue;
        double vx[4], vy[4];
        for (int v = 0; v < 4; ++v) {
            vx[v] = g->pts[(c * 4 + v) * 3 + 0];
            vy[v] = g->pts[(c * 4 + v) * 3 + 1];
        }
        const double s = xi[2 * n], t = xi[2 * n + 1];
        const double s1 = 1.0 - s, t1 = 1.0 - t;
        const double rxx = (vx[1] - vx[0]) * t1 + (vx[2] - vx[3]) * t;
        const double rxy = (vy[1] - vy[0]) * t1 + (vy[2] - vy[3]) * t;
        const double rex = (vx[3] - vx[0]) * s1 + (vx[2] - vx[1]) * s;
        const double rey = (vy[3] - vy[0]) * s1 + (vy[2] - vy[1]) * s;
        const double jac = rxx * rey - rxy * rex;
        const double* d = data + 4 * c;
        const double a = d[3] * s1 + d[1] * s;
        const double b = d[0] * t1 + d[2] * t;
        vec[3 * n + 0] = (a * rxx - b * rex) / jac;
        vec[3 * n + 1] = (a * rxy - b * rey) / jac;
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
