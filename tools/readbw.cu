// Pure-read HBM ceiling on this GPU (what K2 can at best reach): grid-stride 256-bit loads, a few shapes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/readbw tools/readbw.cu
#include <cstdio>
#include <cuda_runtime.h>
struct __align__(32) d4 { double x, y, z, w; };
__device__ __forceinline__ d4 ld256(const d4* p) {
    d4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
template <int U>
__global__ void rd(const d4* __restrict__ a, size_t n, double* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    double s = 0;
    for (; i + (U - 1) * blockDim.x < n; i += stride) {
        d4 v[U];
#pragma unroll
        for (int q = 0; q < U; ++q) v[q] = ld256(a + i + q * blockDim.x);
#pragma unroll
        for (int q = 0; q < U; ++q) s += v[q].x + v[q].y + v[q].z + v[q].w;
    }
    if (s == 1.2345e300) *out = s;
}
// K2-like: every CTA walks nz planes (stride plane) for a tile of columns, two arrays
template <int U>
__global__ void rd_planes(const d4* __restrict__ a, const d4* __restrict__ b, size_t plane4, int nz, double* out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plane4) return;
    const d4* pa = a + (size_t)blockIdx.y * nz * plane4 + c;
    const d4* pb = b + (size_t)blockIdx.y * nz * plane4 + c;
    double s = 0;
    for (int k = 0; k + U <= nz; k += U) {
        d4 v[U], w[U];
#pragma unroll
        for (int q = 0; q < U; ++q) { v[q] = ld256(pa + (size_t)(k + q) * plane4); w[q] = ld256(pb + (size_t)(k + q) * plane4); }
#pragma unroll
        for (int q = 0; q < U; ++q) s += v[q].x + v[q].y + v[q].z + v[q].w + w[q].x + w[q].y + w[q].z + w[q].w;
    }
    if (s == 1.2345e300) *out = s;
}
__global__ void fill_rand(double* a, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long x = i * 0x9E3779B97F4A7C15ull; x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        a[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
    }
}
// K2 math on top of the plane walk: MODE 0 = mul+add+nan select (what K2 does), 1 = fma + nan select, 2 = fma only
template <int U, int MODE>
__global__ void rd_planes_math(const d4* __restrict__ a, const d4* __restrict__ b, size_t plane4, int nz, double* out) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plane4) return;
    const d4* pa = a + (size_t)blockIdx.y * nz * plane4 + c;
    const d4* pb = b + (size_t)blockIdx.y * nz * plane4 + c;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k + U <= nz; k += U) {
        d4 v[U], w[U];
#pragma unroll
        for (int q = 0; q < U; ++q) { v[q] = ld256(pa + (size_t)(k + q) * plane4); w[q] = ld256(pb + (size_t)(k + q) * plane4); }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const double d = 1.0 + 0.25 * (k + q);
            double x[8] = {v[q].x, v[q].y, v[q].z, v[q].w, w[q].x, w[q].y, w[q].z, w[q].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                double y = x[e];
                if (MODE <= 1) y = (y != y) ? 0.0 : y;
                if (MODE == 0) s[e] = __dadd_rn(s[e], __dmul_rn(d, y)); else s[e] = fma(d, y, s[e]);
            }
        }
    }
    double t = 0;
    for (int e = 0; e < 8; ++e) t += s[e];
    if (t == 1.2345e300) *out = t;
}
// + dz in shared memory (DZ) and + arc loads / eflux stores (EPI): what is left between rd_planes_math and K2
template <int ST>
__device__ __forceinline__ void st256(d4* p, const d4& v) {
    if (ST == 0) *p = v;
    if (ST == 1) asm volatile("st.global.L1::no_allocate.L2::evict_last.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    if (ST == 2) asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
    if (ST == 3) { __stcs((double2*)p, make_double2(v.x, v.y)); __stcs((double2*)p + 1, make_double2(v.z, v.w)); }
    if (ST == 4) { __stwt((double2*)p, make_double2(v.x, v.y)); __stwt((double2*)p + 1, make_double2(v.z, v.w)); }
    if (ST == 5) asm volatile("st.global.L1::no_allocate.L2::evict_normal.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
template <int U, bool DZ, bool EPI, int ST = 0, int RING = 0>
__global__ void rd_planes_full(const d4* __restrict__ a, const d4* __restrict__ b, size_t plane4, int nz, const double* dzg,
                               const d4* __restrict__ arc1, const d4* __restrict__ arc2, d4* __restrict__ eflux, double* out) {
    __shared__ double s_dz[128];
    if (DZ) {
        for (int k = threadIdx.x; k < nz; k += blockDim.x) s_dz[k] = dzg[k];
        __syncthreads();
    }
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plane4) return;
    const d4* pa = a + (size_t)blockIdx.y * nz * plane4 + c;
    const d4* pb = b + (size_t)blockIdx.y * nz * plane4 + c;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = 0; k + U <= nz; k += U) {
        d4 v[U], w[U];
#pragma unroll
        for (int q = 0; q < U; ++q) { v[q] = ld256(pa + (size_t)(k + q) * plane4); w[q] = ld256(pb + (size_t)(k + q) * plane4); }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            const double d = DZ ? s_dz[k + q] : 1.0 + 0.25 * (k + q);
            double x[8] = {v[q].x, v[q].y, v[q].z, v[q].w, w[q].x, w[q].y, w[q].z, w[q].w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                double y = x[e];
                y = (y != y) ? 0.0 : y;
                s[e] = __dadd_rn(s[e], __dmul_rn(d, y));
            }
        }
    }
    if (EPI) {
        const d4 r1 = arc1[c], r2 = arc2[c];
        d4 o1 = {s[0] * r1.x, s[1] * r1.y, s[2] * r1.z, s[3] * r1.w};
        d4 o2 = {-s[4] * r2.x, -s[5] * r2.y, -s[6] * r2.z, -s[7] * r2.w};
        const size_t slot = RING ? (blockIdx.y % RING) : blockIdx.y;   // RING: re-use a few L2-resident time-step slots
        st256<ST>(eflux + slot * 2 * plane4 + c, o1);
        st256<ST>(eflux + slot * 2 * plane4 + plane4 + c, o2);
    } else {
        double t = 0;
        for (int e = 0; e < 8; ++e) t += s[e];
        if (t == 1.2345e300) *out = t;
    }
}
int main() {
    const size_t bytes = (size_t)24 << 30;
    d4* a; double* out;
    cudaMalloc(&a, bytes); cudaMalloc(&out, 8);
    cudaMemset(a, 0, bytes);
    const size_t n = bytes / 32;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto time = [&](auto launch, const char* name) {
        for (int i = 0; i < 2; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-40s %8.3f ms  %8.1f GB/s  (%s)\n", name, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32, 148 * 64}) {
        char nm[64];
        snprintf(nm, 64, "gridstride U=4 256thr grid=%d", blocks); time([&] { rd<4><<<blocks, 256>>>(a, n, out); }, nm);
        snprintf(nm, 64, "gridstride U=8 256thr grid=%d", blocks); time([&] { rd<8><<<blocks, 256>>>(a, n, out); }, nm);
    }
    time([&] { rd<4><<<148 * 8, 512>>>(a, n, out); }, "gridstride U=4 512thr grid=1184");
    time([&] { rd<8><<<148 * 8, 1024>>>(a, n, out); }, "gridstride U=8 1024thr grid=1184");
    // K2-like geometry: plane of 120184 doubles (C3) = 30046 d4, nz=75, two arrays of 12 GB each
    {
        const size_t plane4 = 30046; const int nz = 75;
        const int nt = (int)((bytes / 2) / (32 * plane4 * nz));
        const d4* b = a + (size_t)nt * nz * plane4;
        const double used = 2.0 * nt * nz * plane4 * 32;
        for (int i = 0; i < 2; ++i) rd_planes<5><<<dim3((plane4 + 255) / 256, nt), 256>>>(a, b, plane4, nz, out);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) rd_planes<5><<<dim3((plane4 + 255) / 256, nt), 256>>>(a, b, plane4, nz, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-40s %8.3f ms  %8.1f GB/s\n", "K2-like planes U=5 (no math)", ms, used / ms / 1e6);
        for (int i = 0; i < 2; ++i) rd_planes<15><<<dim3((plane4 + 255) / 256, nt), 256>>>(a, b, plane4, nz, out);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) rd_planes<15><<<dim3((plane4 + 255) / 256, nt), 256>>>(a, b, plane4, nz, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-40s %8.3f ms  %8.1f GB/s\n", "K2-like planes U=15 (no math)", ms, used / ms / 1e6);
    }
    // the same walks on random (non-zero) data, and with K2's arithmetic
    {
        fill_rand<<<148 * 16, 256>>>((double*)a, bytes / 8);
        cudaDeviceSynchronize();
        time([&] { rd<8><<<148 * 4, 256>>>(a, n, out); }, "gridstride U=8 grid=592, random data");
        const size_t plane4 = 30046; const int nz = 75;
        const int nt = (int)((bytes / 2) / (32 * plane4 * nz));
        const d4* b = a + (size_t)nt * nz * plane4;
        const double used = 2.0 * nt * nz * plane4 * 32;
        dim3 grid((plane4 + 255) / 256, nt);
        auto timeK = [&](auto launch, const char* name) {
            for (int i = 0; i < 2; ++i) launch();
            cudaEventRecord(e0);
            for (int i = 0; i < 5; ++i) launch();
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
            printf("%-40s %8.3f ms  %8.1f GB/s\n", name, ms, used / ms / 1e6);
        };
        timeK([&] { rd_planes<5><<<grid, 256>>>(a, b, plane4, nz, out); }, "planes U=5 no math, random data");
        timeK([&] { rd_planes_math<5, 0><<<grid, 256>>>(a, b, plane4, nz, out); }, "planes U=5 mul+add+nan, random");
        timeK([&] { rd_planes_math<5, 1><<<grid, 256>>>(a, b, plane4, nz, out); }, "planes U=5 fma+nan, random");
        timeK([&] { rd_planes_math<5, 2><<<grid, 256>>>(a, b, plane4, nz, out); }, "planes U=5 fma only, random");
        timeK([&] { rd_planes_math<15, 0><<<grid, 256>>>(a, b, plane4, nz, out); }, "planes U=15 mul+add+nan, random");
        double* dzg; cudaMalloc(&dzg, 1024); cudaMemset(dzg, 0, 1024);
        d4 *arc1, *arc2, *ef;
        cudaMalloc(&arc1, plane4 * 32); cudaMalloc(&arc2, plane4 * 32); cudaMalloc(&ef, (size_t)nt * 2 * plane4 * 32);
        cudaMemset(arc1, 0, plane4 * 32); cudaMemset(arc2, 0, plane4 * 32);
        timeK([&] { rd_planes_full<5, true, false><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "planes U=5 math + smem dz");
        timeK([&] { rd_planes_full<5, false, true><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "planes U=5 math + epilogue stores");
        timeK([&] { rd_planes_full<5, true, true><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "planes U=5 math + dz + epilogue (=K2)");
        timeK([&] { rd_planes_full<5, true, true, 1><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  stores L2::evict_last");
        timeK([&] { rd_planes_full<5, true, true, 2><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  stores L2::evict_first");
        timeK([&] { rd_planes_full<5, true, true, 3><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  stores __stcs");
        timeK([&] { rd_planes_full<5, true, true, 4><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  stores __stwt");
        timeK([&] { rd_planes_full<5, true, true, 0, 8><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  default stores into a ring of 8 steps");
        timeK([&] { rd_planes_full<5, true, true, 1, 8><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  evict_last stores into a ring of 8 steps");
        timeK([&] { rd_planes_full<5, true, true, 0, 32><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  default stores into a ring of 32 steps");
        timeK([&] { rd_planes_full<5, true, true, 5><<<grid, 256>>>(a, b, plane4, nz, dzg, arc1, arc2, ef, out); }, "  stores L2::evict_normal no_allocate");
    }
    // plain cudaMemcpy D2D for reference (read+write bytes)
    {
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) cudaMemcpyAsync((char*)a + bytes / 2, a, bytes / 2, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
        printf("%-40s %8.3f ms  %8.1f GB/s (read+write)\n", "cudaMemcpy D2D 12 GB", ms, bytes / ms / 1e6);
    }
    return 0;
}
