// K3: batched sparse weight * edge-flux reduction (sm_100a).
//
// Replaces the double loop of /root/reference/nemoflux/fluxplot.py:51-59 (for every time step, for
// every transect, mint PolylineIntegral.getIntegral(integratedVelocity, CELL_BY_CELL_DATA),
// field.py:102): series[t, m] = sum_n w[n] * data[t*stride_t + idx[n]] over row m of a CSR built by
// K1.  idx < 0 marks the never-written south edge of row 0 (always 0, field.py:61,219).
//
// One warp per (transect, time step): lanes stride over the row, each lane adds its terms in row
// order, then a fixed xor-shuffle tree combines the 32 partial sums -> deterministic run to run.
#include "nfx_common.cuh"

namespace nfx {

namespace {

__global__ void __launch_bounds__(256)
k3_integrate(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ idx, const double* __restrict__ w,
             int ntransects, const double* __restrict__ data, int64_t stride_t, int nt, double* __restrict__ series,
             int64_t out_stride_t) {
    const int lane = threadIdx.x & 31;
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (wid >= (int64_t)nt * ntransects) return;
    const int64_t t = wid / ntransects;
    const int m = (int)(wid - t * ntransects);
    const int64_t r0 = rowptr[m], r1 = rowptr[m + 1];
    const double* d = data + t * stride_t;
    double acc = 0.0;
    for (int64_t n = r0 + lane; n < r1; n += 32) {
        const int32_t k = idx[n];
        const double f = k >= 0 ? d[k] : 0.0;
        acc = fma(w[n], f, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) series[t * out_stride_t + m] = acc;
}

// series[t, m] = sum over panels (ascending, sequential -> deterministic) of partial[t, panel, m]
__global__ void k_reduce_panels(const double* __restrict__ partial, int nt, int npanels, int ntransects,
                                double* __restrict__ series) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nt * ntransects) return;
    const int64_t t = i / ntransects;
    const int m = (int)(i - t * ntransects);
    double acc = 0.0;
    for (int p = 0; p < npanels; ++p) acc += partial[(t * npanels + p) * ntransects + m];
    series[i] = acc;
}

}  // namespace

void csr_integrate(const Csr& c, int ntransects, const double* data, int64_t stride_t, int nt, double* series,
                   cudaStream_t s) {
    NFX_REQUIRE(data && series, "integrate: NULL pointer");
    if (nt <= 0 || ntransects <= 0) return;
    const int64_t nw = (int64_t)nt * ntransects;
    const unsigned blocks = (unsigned)((nw * 32 + 255) / 256);
    k3_integrate<<<blocks, 256, 0, s>>>(c.rowptr.p, c.idx.p, c.w, ntransects, data, stride_t, nt, series, ntransects);
    count_launch();
    NFX_CUDA(cudaGetLastError());
}

// the same kernel on an arbitrary block of CSR rows (one panel of the fast path): out[t*out_stride_t + row]
void rows_integrate(const int64_t* rowptr, const int32_t* idx, const double* w, int nrows, const double* data,
                    int64_t stride_t, int nt, double* out, int64_t out_stride_t, cudaStream_t s) {
    if (nt <= 0 || nrows <= 0) return;
    const int64_t nw = (int64_t)nt * nrows;
    const unsigned blocks = (unsigned)((nw * 32 + 255) / 256);
    k3_integrate<<<blocks, 256, 0, s>>>(rowptr, idx, w, nrows, data, stride_t, nt, out, out_stride_t);
    count_launch();
    NFX_CUDA(cudaGetLastError());
}

void reduce_panels(const double* partial, int nt, int npanels, int ntransects, double* series, cudaStream_t s) {
    const int64_t n = (int64_t)nt * ntransects;
    if (n <= 0) return;
    k_reduce_panels<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(partial, nt, npanels, ntransects, series);
    count_launch();
    NFX_CUDA(cudaGetLastError());
}

}  // namespace nfx
