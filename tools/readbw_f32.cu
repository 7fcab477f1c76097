// What bounds the float32-storage edge-flux stream?  K2-like plane walks over float data (C3 plane, nz = 75, two
// arrays) with the pieces of the kernel added one at a time: loads only, + conversion flavours, + block shapes.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/readbw_f32 tools/readbw_f32.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../nemoflux_b200/csrc/nfx_stream_ops.cuh"

using namespace nfx::dev;

// the second conversion flavour that was A/B'd against clean_scaled and dropped (same bits, same speed in k23_fused):
// one ordered compare against the marker (or +inf), hi and lo word from one signed 32 x 32 -> 64 multiply by 2^29
__device__ __forceinline__ double clean_scaled_v2(float x, float marker_or_inf, float& amax) {
    amax = fmaxf(amax, fabsf(x));
    uint32_t b;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ne.f32 p, %1, %2;\n\t"
        "selp.b32 %0, %3, 0, p;\n\t}"
        : "=r"(b)
        : "f"(x), "f"(marker_or_inf), "r"(__float_as_uint(x)));
    long long wide;
    asm("mul.wide.s32 %0, %1, 0x20000000;" : "=l"(wide) : "r"(b));
    const uint32_t hi = (uint32_t)((unsigned long long)wide >> 32) & 0x8fffffffu;
    return __hiloint2double((int)hi, (int)(uint32_t)wide);
}

// MATH: 0 = float adds only, 1 = (double)x (F2F on the XU pipe) + dmul + dadd, 2 = clean_scaled (bit shuffle) + dmul + dadd,
//       3 = clean_scaled_v2, 4 = bit shuffle without any NaN/marker handling (lower bound of the ALU cost)
template <int BLOCK, int VEC, int U, int MATH, int MINB>
__global__ void __launch_bounds__(BLOCK, MINB)
walk(const float* __restrict__ a, const float* __restrict__ b, size_t plane, int nz, double* out) {
    using P = Pack<float, VEC>;
    using V = typename P::type;
    const size_t c = ((size_t)blockIdx.x * BLOCK + threadIdx.x) * VEC;
    if (c >= plane) return;
    const float* pa = a + (size_t)blockIdx.y * nz * plane + c;
    const float* pb = b + (size_t)blockIdx.y * nz * plane + c;
    const uint64_t pol = l2_evict_first_policy();
    double su[VEC], sv[VEC];
    float fu[VEC], fv[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) su[e] = sv[e] = 0.0, fu[e] = fv[e] = 0.f;
    float amax = 0.f;
    const float fill = __int_as_float(0x7fc00000);
    for (int k = 0; k + U <= nz; k += U) {
        V ru[U], rv[U];
#pragma unroll
        for (int q = 0; q < U; ++q) {
            ru[q] = ld_stream(reinterpret_cast<const V*>(pa + (size_t)(k + q) * plane), pol);
            rv[q] = ld_stream(reinterpret_cast<const V*>(pb + (size_t)(k + q) * plane), pol);
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            pin(ru[q]);
            pin(rv[q]);
        }
#pragma unroll
        for (int q = 0; q < U; ++q) {
            float x[VEC], y[VEC];
            P::unpack(ru[q], x);
            P::unpack(rv[q], y);
            const double d = 1.0 + 0.25 * (k + q);
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                if (MATH == 0) {
                    fu[e] += x[e];
                    fv[e] += y[e];
                } else if (MATH == 1) {
                    su[e] = __dadd_rn(su[e], __dmul_rn(d, clean<float>(x[e], fill, false)));
                    sv[e] = __dadd_rn(sv[e], __dmul_rn(d, clean<float>(y[e], fill, false)));
                } else if (MATH == 2) {
                    su[e] = __dadd_rn(su[e], __dmul_rn(d, clean_scaled(x[e], fill, false, amax)));
                    sv[e] = __dadd_rn(sv[e], __dmul_rn(d, clean_scaled(y[e], fill, false, amax)));
                } else if (MATH == 3) {
                    su[e] = __dadd_rn(su[e], __dmul_rn(d, clean_scaled_v2(x[e], fill, amax)));
                    sv[e] = __dadd_rn(sv[e], __dmul_rn(d, clean_scaled_v2(y[e], fill, amax)));
                } else {
                    const uint32_t bu = __float_as_uint(x[e]), bv = __float_as_uint(y[e]);
                    su[e] = __dadd_rn(su[e], __dmul_rn(d, __hiloint2double((int)(((int32_t)bu >> 3) & 0x8fffffffu), (int)(bu << 29))));
                    sv[e] = __dadd_rn(sv[e], __dmul_rn(d, __hiloint2double((int)(((int32_t)bv >> 3) & 0x8fffffffu), (int)(bv << 29))));
                }
            }
        }
    }
    double t = amax;
#pragma unroll
    for (int e = 0; e < VEC; ++e) t += su[e] + sv[e] + fu[e] + fv[e];
    if (t == 1.2345e300) *out = t;
}


// ---- the pieces of the fused pass added one at a time on top of walk<256, 4, 5, 2, 4> --------------------------
// EPI: arc loads + edge fluxes stored with evict_last into a ring of RING time steps; DZ: dz * 2^896 from shared memory;
// PERSIST: persistent CTAs taking tiles from an atomic counter (two __syncthreads per tile, as k23_fused does)
template <bool EPI, bool DZ, bool PERSIST, int RING>
__global__ void __launch_bounds__(256, 4)
walk2(const float* __restrict__ a, const float* __restrict__ b, size_t plane, int nz, int nt, const double* __restrict__ dzg,
      const double* __restrict__ arc1, const double* __restrict__ arc2, double* __restrict__ ring, int* counter, double* out) {
    constexpr int VEC = 4, U = 5;
    using P = Pack<float, VEC>;
    using V = typename P::type;
    __shared__ double s_dz[128];
    __shared__ int s_item;
    if (DZ) {
        for (int k = threadIdx.x; k < nz; k += 256) s_dz[k] = dzg[k];
        __syncthreads();
    }
    const uint64_t pol = l2_evict_first_policy();
    const uint64_t pol_el = l2_evict_last_policy();
    const float fill = __int_as_float(0x7fc00000);
    const int ntiles = (int)((plane / VEC + 255) / 256);
    int item = PERSIST ? 0 : (int)(blockIdx.y * ntiles + blockIdx.x);
    for (;;) {
        if (PERSIST) {
            if (threadIdx.x == 0) s_item = atomicAdd(counter, 1);
            __syncthreads();
            item = s_item;
            __syncthreads();
            if (item >= ntiles * nt) break;
        }
        const int t = item / ntiles, r = item - t * ntiles;
        const size_t c = ((size_t)r * 256 + threadIdx.x) * VEC;
        if (c < plane) {
            const float* pa = a + (size_t)t * nz * plane + c;
            const float* pb = b + (size_t)t * nz * plane + c;
            double su[VEC], sv[VEC];
#pragma unroll
            for (int e = 0; e < VEC; ++e) su[e] = sv[e] = 0.0;
            float amax = 0.f;
            for (int k = 0; k + U <= nz; k += U) {
                V ru[U], rv[U];
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    ru[q] = ld_stream(reinterpret_cast<const V*>(pa + (size_t)(k + q) * plane), pol);
                    rv[q] = ld_stream(reinterpret_cast<const V*>(pb + (size_t)(k + q) * plane), pol);
                }
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    pin(ru[q]);
                    pin(rv[q]);
                }
#pragma unroll
                for (int q = 0; q < U; ++q) {
                    float x[VEC], y[VEC];
                    P::unpack(ru[q], x);
                    P::unpack(rv[q], y);
                    const double d = DZ ? s_dz[k + q] : 1.0 + 0.25 * (k + q);
#pragma unroll
                    for (int e = 0; e < VEC; ++e) {
                        su[e] = __dadd_rn(su[e], __dmul_rn(d, clean_scaled(x[e], fill, false, amax)));
                        sv[e] = __dadd_rn(sv[e], __dmul_rn(d, clean_scaled(y[e], fill, false, amax)));
                    }
                }
            }
            if (EPI) {
                double* ou = ring + (size_t)(t % RING) * 2 * plane + c;
                double* ov = ou + plane;
#pragma unroll
                for (int e = 0; e < VEC; e += 2) {
                    st_stream2(ou + e, su[e] * arc1[c + e], su[e + 1] * arc1[c + e + 1], 1, pol_el);
                    st_stream2(ov + e, -sv[e] * arc2[c + e], -sv[e + 1] * arc2[c + e + 1], 1, pol_el);
                }
                if (amax == 1.2345f) *out = amax;
            } else {
                double tt = amax;
#pragma unroll
                for (int e = 0; e < VEC; ++e) tt += su[e] + sv[e];
                if (tt == 1.2345e300) *out = tt;
            }
        }
        if (!PERSIST) break;
        __syncthreads();
    }
}

__global__ void fill_rand(float* a, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned long long x = i * 0x9E3779B97F4A7C15ull;
        x ^= x >> 29;
        x *= 0xBF58476D1CE4E5B9ull;
        x ^= x >> 32;
        a[i] = (x % 10 < 3) ? __int_as_float(0x7fc00000) : (float)(x >> 40) * (1.0f / 16777216.0f) - 0.5f;
    }
}

int main() {
    const size_t plane = 120184;   // C3: 362 x 332 cells, divisible by 8
    const int nz = 75;
    const int nt = 660;            // 2 x 660 x 75 x 120184 x 4 B = 47.6 GB
    const size_t n1 = (size_t)nt * nz * plane;
    float *a, *b;
    double* out;
    cudaMalloc(&a, n1 * 4);
    cudaMalloc(&b, n1 * 4);
    cudaMalloc(&out, 8);
    fill_rand<<<148 * 16, 256>>>(a, n1);
    fill_rand<<<148 * 16, 256>>>(b, n1);
    cudaDeviceSynchronize();
    const double used = 2.0 * n1 * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto timeK = [&](auto launch, const char* name) {
        for (int i = 0; i < 2; ++i) launch();
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        ms /= 5;
        printf("%-56s %8.3f ms  %8.1f GB/s  (%s)\n", name, ms, used / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    };
#define RUN(BLOCK, VEC, U, MATH, MINB, label)                                                                          \
    timeK([&] {                                                                                                         \
        walk<BLOCK, VEC, U, MATH, MINB><<<dim3((plane / VEC + BLOCK - 1) / BLOCK, nt), BLOCK>>>(a, b, plane, nz, out); \
    }, label)
    RUN(256, 4, 5, 0, 4, "128-bit x5, block 256 (4 KB rows), float adds only");
    RUN(512, 4, 5, 0, 2, "128-bit x5, block 512 (8 KB rows), float adds only");
    RUN(256, 8, 5, 0, 3, "256-bit x5, block 256 (8 KB rows), float adds only");
    RUN(256, 8, 3, 0, 4, "256-bit x3, block 256 (8 KB rows), float adds only");
    RUN(128, 8, 5, 0, 6, "256-bit x5, block 128 (4 KB rows), float adds only");
    RUN(256, 4, 5, 1, 4, "128-bit x5, block 256, F2F + dmul + dadd");
    RUN(256, 4, 5, 2, 4, "128-bit x5, block 256, clean_scaled + dmul + dadd");
    RUN(256, 4, 5, 3, 4, "128-bit x5, block 256, clean_scaled_v2 + dmul + dadd");
    RUN(256, 4, 5, 4, 4, "128-bit x5, block 256, raw bit shuffle + dmul + dadd");
    RUN(512, 4, 5, 2, 2, "128-bit x5, block 512, clean_scaled + dmul + dadd");
    RUN(256, 8, 5, 2, 3, "256-bit x5, block 256, clean_scaled (3 CTAs/SM)");
    RUN(256, 8, 3, 2, 4, "256-bit x3, block 256, clean_scaled (4 CTAs/SM)");
    RUN(256, 8, 3, 4, 4, "256-bit x3, block 256, raw bit shuffle (4 CTAs/SM)");
    RUN(256, 4, 3, 2, 5, "128-bit x3, block 256, clean_scaled (5 CTAs/SM)");
    RUN(256, 4, 5, 2, 3, "128-bit x5, block 256, clean_scaled (3 CTAs/SM)");
    RUN(256, 4, 15, 2, 2, "128-bit x15, block 256, clean_scaled (2 CTAs/SM)");
    {
        double *dzg, *arc1, *arc2, *ring;
        int* counter;
        cudaMalloc(&dzg, 1024);
        cudaMemset(dzg, 0, 1024);
        cudaMalloc(&arc1, plane * 8);
        cudaMalloc(&arc2, plane * 8);
        cudaMemset(arc1, 0, plane * 8);
        cudaMemset(arc2, 0, plane * 8);
        cudaMalloc(&ring, (size_t)8 * 2 * plane * 8);
        cudaMalloc(&counter, 4);
        const dim3 grid((plane / 4 + 255) / 256, nt);
#define RUN2(EPI, DZ, PERSIST, label)                                                                                    \
    timeK([&] {                                                                                                          \
        if (PERSIST) cudaMemsetAsync(counter, 0, 4);                                                                     \
        walk2<EPI, DZ, PERSIST, 8><<<PERSIST ? dim3(592) : grid, 256>>>(a, b, plane, nz, nt, dzg, arc1, arc2, ring, counter, out); \
    }, label)
        RUN2(false, false, false, "walk2: plain launch, no epilogue");
        RUN2(false, true, false, "walk2: + dz from shared memory");
        RUN2(true, true, false, "walk2: + arc loads, evict_last stores to a ring of 8 steps");
        RUN2(false, true, true, "walk2: persistent (atomic tile counter), no epilogue");
        RUN2(true, true, true, "walk2: persistent + epilogue (= K2 side of k23_fused)");
    }
    return 0;
}
