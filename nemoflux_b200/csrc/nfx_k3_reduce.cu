// K3: batched sparse weight * edge-flux reduction (sm_100a).
//
// Replaces the double loop of /root/reference/nemoflux/fluxplot.py:51-59 (for every time step, for
// every transect, mint PolylineIntegral.getIntegral(integratedVelocity, CELL_BY_CELL_DATA),
// field.py:102): series[t, m] = sum_n w[n] * data[t*stride_t + idx[n]] over row m of a CSR built by
// K1.  idx < 0 marks the never-written south edge of row 0 (always 0, field.py:61,219).
//
// One CTA per (transect, time step): threads stride over the row, each adds its terms in row order, a fixed
// xor-shuffle tree combines the lanes of a warp, the 8 warp sums are added in order -> deterministic.
#include "nfx_common.cuh"

namespace nfx {

namespace {

constexpr int kK3Block = 256;
constexpr int kK3Warps = kK3Block / 32;

// G warps share one (transect, time step) row, 8/G rows per CTA: thread-strided partial sums, shuffle tree per
// warp, the group's warp sums added in warp order -> deterministic; the same scheme as the fused pass.
__global__ void __launch_bounds__(kK3Block)
k3_integrate(const int64_t* __restrict__ rowptr, const int32_t* __restrict__ idx, const double* __restrict__ w,
             int ntransects, const double* __restrict__ data, int64_t stride_t, int nt, double* __restrict__ series,
             int64_t out_stride_t, int G) {
    __shared__ double s_part[kK3Warps];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int rows_per_block = kK3Warps / G;
    const int64_t rid = (int64_t)blockIdx.x * rows_per_block + wid / G;
    const bool valid = rid < (int64_t)nt * ntransects;
    const int64_t t = valid ? rid / ntransects : 0;
    const int m = valid ? (int)(rid - t * ntransects) : 0;
    double acc = 0.0;
    if (valid) {
        const int64_t r0 = rowptr[m], r1 = rowptr[m + 1];
        const double* d = data + t * stride_t;
        const int stride = G * 32;
#pragma unroll 4
        for (int64_t n = r0 + (wid % G) * 32 + lane; n < r1; n += stride) {
            const int32_t k = idx[n];
            const double f = k >= 0 ? d[k] : 0.0;
            acc = fma(w[n], f, acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) s_part[wid] = acc;
    __syncthreads();
    if (valid && lane == 0 && (wid % G) == 0) {
        double tot = s_part[wid];
        for (int k = 1; k < G; ++k) tot += s_part[wid + k];
        series[t * out_stride_t + m] = tot;
    }
}

// series[t, m] = sum of the partial sums of the transect's sub-rows, ascending and sequential -> deterministic
__global__ void k_reduce_subrows(const double* __restrict__ partial, int nt, int64_t nsr, const int64_t* __restrict__ tr_ptr,
                                 const int64_t* __restrict__ tr_sr, int ntransects, double* __restrict__ series) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)nt * ntransects) return;
    const int64_t t = i / ntransects;
    const int m = (int)(i - t * ntransects);
    double acc = 0.0;
    for (int64_t j = tr_ptr[m]; j < tr_ptr[m + 1]; ++j) acc += partial[t * nsr + tr_sr[j]];
    series[i] = acc;
}

}  // namespace

void csr_integrate(const Csr& c, int ntransects, const double* data, int64_t stride_t, int nt, double* series,
                   cudaStream_t s) {
    NFX_REQUIRE(data && series, "integrate: NULL pointer");
    if (nt <= 0 || ntransects <= 0) return;
    const int64_t nrows = (int64_t)nt * ntransects;
    const int G = k3_group_for(c.nnz, ntransects);
    const int64_t blocks = (nrows + kK3Warps / G - 1) / (kK3Warps / G);
    NFX_REQUIRE(blocks < 2147483647ll, "integrate: nt * ntransects too large");
    k3_integrate<<<(unsigned)blocks, kK3Block, 0, s>>>(c.rowptr.p, c.idx.p, c.w, ntransects, data, stride_t, nt, series,
                                                      ntransects, G);
    count_launch();
    NFX_CUDA(cudaGetLastError());
}

// the same kernel on an arbitrary block of CSR rows (one panel of the fast path): out[t*out_stride_t + row]
void rows_integrate(const int64_t* rowptr, const int32_t* idx, const double* w, int nrows, int64_t nnz,
                    const double* data, int64_t stride_t, int nt, double* out, int64_t out_stride_t, cudaStream_t s) {
    if (nt <= 0 || nrows <= 0) return;
    const int64_t n = (int64_t)nt * nrows;
    const int G = k3_group_for(nnz, nrows);
    const int64_t blocks = (n + kK3Warps / G - 1) / (kK3Warps / G);
    NFX_REQUIRE(blocks < 2147483647ll, "integrate: nt * rows too large");
    k3_integrate<<<(unsigned)blocks, kK3Block, 0, s>>>(rowptr, idx, w, nrows, data, stride_t, nt, out, out_stride_t, G);
    count_launch();
    NFX_CUDA(cudaGetLastError());
}

void reduce_subrows(const double* partial, int nt, int64_t nsr, const int64_t* tr_ptr, const int64_t* tr_sr,
                    int ntransects, double* series, cudaStream_t s) {
    const int64_t n = (int64_t)nt * ntransects;
    if (n <= 0) return;
    k_reduce_subrows<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(partial, nt, nsr, tr_ptr, tr_sr, ntransects, series);
    count_launch();
    NFX_CUDA(cudaGetLastError());
}

}  // namespace nfx
