// K4: vector interpolation on the transect (viz only in the reference).
//
// Replaces mint.VectorInterp.findPoints(points, tol2) + getFaceVectors(data, placement=0)
// (/root/reference/nemoflux/field.py:90-95,119-120).
//   find_points : one warp per point walks the grid's bounding-box hierarchy (the locator K1 uses), lanes test
//                 32 boxes / 32 cells at a time; the containing cell with the LOWEST id wins (points on shared
//                 edges belong to several cells); x-period images are tried in the order 0, -period, +period;
//                 lane 0 inverts the bilinear map (Newton, same routine as K1).
//   face_vectors: one thread per point, v = [(d3 (1-xi) + d1 xi) r_xi - (d0 (1-eta) + d2 eta) r_eta] / J with the
//                 cell-by-cell edge data d0..d3 (south, east, north, west).  Sign convention pinned by
//                 pictures/simple.png and pictures/singular.png of the reference.
// Fixed operation order without FMA contraction: bit-identical to the C oracle.
#include <math_constants.h>

#include "nfx_common.cuh"

namespace nfx {

namespace {

constexpr double kEps = 10.0 * 2.220446049250313e-16;
constexpr double kMargin = 1.0e-7;

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ bool contains_point(const double2 (&v)[4], double px, double py, double tol) {
    bool inside = true;
#pragma unroll
    for (int i0 = 0; i0 < 4; ++i0) {
        const int i1 = (i0 + 1) & 3;
        const double d0x = dsub(px, v[i0].x), d0y = dsub(py, v[i0].y);
        const double d1x = dsub(px, v[i1].x), d1y = dsub(py, v[i1].y);
        if (dsub(dmul(d0x, d1y), dmul(d0y, d1x)) < -tol) inside = false;
    }
    return inside;
}

__device__ __forceinline__ void param_coords(const double2 (&v)[4], double px, double py, double& xi0, double& xi1) {
    const double bx = dsub(v[1].x, v[0].x), by = dsub(v[1].y, v[0].y);
    const double cx = dsub(v[3].x, v[0].x), cy = dsub(v[3].y, v[0].y);
    const double dx = dadd(dsub(v[0].x, v[1].x), dsub(v[2].x, v[3].x));
    const double dy = dadd(dsub(v[0].y, v[1].y), dsub(v[2].y, v[3].y));
    double s = 0.5, t = 0.5;
    for (int it = 0; it < 20; ++it) {
        const double st = dmul(s, t);
        const double fx = dsub(dadd(dadd(dadd(v[0].x, dmul(bx, s)), dmul(cx, t)), dmul(dx, st)), px);
        const double fy = dsub(dadd(dadd(dadd(v[0].y, dmul(by, s)), dmul(cy, t)), dmul(dy, st)), py);
        const double j00 = dadd(bx, dmul(dx, t)), j01 = dadd(cx, dmul(dx, s));
        const double j10 = dadd(by, dmul(dy, t)), j11 = dadd(cy, dmul(dy, s));
        const double det = dsub(dmul(j00, j11), dmul(j01, j10));
        const double ds = ddiv(dsub(dmul(fy, j01), dmul(fx, j11)), det);
        const double dt = ddiv(dsub(dmul(fx, j10), dmul(fy, j00)), det);
        s = dadd(s, ds);
        t = dadd(t, dt);
        if (fabs(ds) < 1.0e-15 && fabs(dt) < 1.0e-15) break;
    }
    xi0 = s;
    xi1 = t;
}

__device__ __forceinline__ bool box_has(const double4& b, double px, double py) {
    return px >= b.x - kMargin && px <= b.y + kMargin && py >= b.z - kMargin && py <= b.w + kMargin;
}

__global__ void __launch_bounds__(128)
k4_find_points(const double2* __restrict__ verts, int64_t ncell, const double4* __restrict__ box1, int64_t nl1,
               const double4* __restrict__ box2, int64_t nl2, const double4* __restrict__ box3, int64_t nl3,
               const double* __restrict__ xyz, int64_t npts,
               double period_x, double tol, int32_t* __restrict__ cell, double* __restrict__ xi) {
    const int lane = threadIdx.x & 31;
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= npts) return;
    const double py = xyz[3 * n + 1];
    const int nimg = period_x > 0.0 ? 3 : 1;
    int64_t found = -1;
    double fpx = 0.0;
    for (int im = 0; im < nimg && found < 0; ++im) {
        const double shift = im == 0 ? 0.0 : (im == 1 ? -period_x : period_x);
        const double px = dadd(xyz[3 * n], shift);
        int64_t best = INT64_MAX;
        // descend the three box levels; cell ids grow with the node index at every level, so the first hit in
        // traversal order is the lowest containing cell id and everything after it can be skipped
        for (int64_t n3b = 0; n3b < nl3 && best == INT64_MAX; n3b += 32) {
            const int64_t n3 = n3b + lane;
            unsigned mask3 = __ballot_sync(0xffffffffu, n3 < nl3 && box_has(box3[n3 < nl3 ? n3 : 0], px, py));
            for (; mask3 && best == INT64_MAX; mask3 &= mask3 - 1) {
                const int64_t node3 = n3b + (__ffs(mask3) - 1);
                const int64_t n2 = node3 * kFan + lane;
                unsigned mask2 = __ballot_sync(0xffffffffu, n2 < nl2 && box_has(box2[n2 < nl2 ? n2 : 0], px, py));
                for (; mask2 && best == INT64_MAX; mask2 &= mask2 - 1) {
                    const int64_t node2 = node3 * kFan + (__ffs(mask2) - 1);
                    const int64_t n1 = node2 * kFan + lane;
                    unsigned mask1 = __ballot_sync(0xffffffffu, n1 < nl1 && box_has(box1[n1 < nl1 ? n1 : 0], px, py));
                    for (; mask1 && best == INT64_MAX; mask1 &= mask1 - 1) {
                        const int64_t first = (node2 * kFan + (__ffs(mask1) - 1)) * kFan;
                        const int64_t c = first + lane;
                        bool in = false;
                        if (c < ncell) {
                            double2 v[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) v[k] = verts[c * 4 + k];
                            in = contains_point(v, px, py, tol);
                        }
                        const unsigned m = __ballot_sync(0xffffffffu, in);
                        if (m) best = first + (__ffs(m) - 1);
                    }
                }
            }
        }
        if (best != INT64_MAX) {
            found = best;
            fpx = px;
        }
    }
    if (lane == 0) {
        double s = 0.0, t = 0.0;
        if (found >= 0) {
            double2 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = verts[found * 4 + k];
            param_coords(v, fpx, py, s, t);
        }
        cell[n] = (int32_t)found;
        xi[2 * n] = s;
        xi[2 * n + 1] = t;
    }
}

__global__ void k4_face_vectors(const double2* __restrict__ verts, const int32_t* __restrict__ cell,
                                const double* __restrict__ xi, int64_t npts, const double* __restrict__ data,
                                double* __restrict__ vec) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= npts) return;
    double vx = 0.0, vy = 0.0;
    const int64_t c = cell[n];
    if (c >= 0) {
        double2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = verts[c * 4 + k];
        const double s = xi[2 * n], t = xi[2 * n + 1];
        const double s1 = dsub(1.0, s), t1 = dsub(1.0, t);
        const double rxx = dadd(dmul(dsub(v[1].x, v[0].x), t1), dmul(dsub(v[2].x, v[3].x), t));
        const double rxy = dadd(dmul(dsub(v[1].y, v[0].y), t1), dmul(dsub(v[2].y, v[3].y), t));
        const double rex = dadd(dmul(dsub(v[3].x, v[0].x), s1), dmul(dsub(v[2].x, v[1].x), s));
        const double rey = dadd(dmul(dsub(v[3].y, v[0].y), s1), dmul(dsub(v[2].y, v[1].y), s));
        const double jac = dsub(dmul(rxx, rey), dmul(rxy, rex));
        const double* d = data + 4 * c;
        const double a = dadd(dmul(d[3], s1), dmul(d[1], s));
        const double b = dadd(dmul(d[0], t1), dmul(d[2], t));
        vx = ddiv(dsub(dmul(a, rxx), dmul(b, rex)), jac);
        vy = ddiv(dsub(dmul(a, rxy), dmul(b, rey)), jac);
    }
    vec[3 * n] = vx;
    vec[3 * n + 1] = vy;
    vec[3 * n + 2] = 0.0;
}

}  // namespace

void vinterp_find_points(VInterpDev& vi, int64_t npts, const double* xyz_host, double tol2, cudaStream_t s) {
    NFX_REQUIRE(vi.grid && vi.grid->locator_built, "findPoints: setGrid / buildLocator was not called");
    NFX_REQUIRE(npts >= 0 && (npts == 0 || xyz_host), "findPoints: bad arguments");
    GridDev& g = *vi.grid;
    vi.npts = npts;
    vi.cell.ensure((size_t)std::max<int64_t>(npts, 1));
    vi.xi.ensure((size_t)std::max<int64_t>(2 * npts, 1));
    if (npts == 0) return;
    DevBuf<double> d_xyz;
    d_xyz.alloc((size_t)3 * npts);
    NFX_CUDA(cudaMemcpyAsync(d_xyz.p, xyz_host, sizeof(double) * 3 * npts, cudaMemcpyHostToDevice, s));
    const double tol = tol2 > kEps ? tol2 : kEps;
    k4_find_points<<<(unsigned)((npts * 32 + 127) / 128), 128, 0, s>>>(g.verts.p, g.ncell, g.box1.p, g.nl1, g.box2.p, g.nl2, g.box3.p, g.nl3,
                                                                       d_xyz.p, npts, vi.period_x, tol, vi.cell.p, vi.xi.p);
    count_launch();
    NFX_CUDA(cudaGetLastError());
    NFX_CUDA(cudaStreamSynchronize(s));
}

void vinterp_face_vectors(VInterpDev& vi, const double* data_dev, double* vec_dev, cudaStream_t s) {
    NFX_REQUIRE(vi.grid, "getFaceVectors: setGrid was not called");
    if (vi.npts == 0) return;
    k4_face_vectors<<<(unsigned)((vi.npts + 127) / 128), 128, 0, s>>>(vi.grid->verts.p, vi.cell.p, vi.xi.p, vi.npts, data_dev,
                                                                      vec_dev);
    count_launch();
    NFX_CUDA(cudaGetLastError());
}

}  // namespace nfx
