// Device helpers shared by the edge-flux kernels (nfx_k2_edgeflux.cu, nfx_k23_fused.cu): streaming loads of the
// u/v stream, register fences, vector packs, missing-value cleaning, eflux store flavours.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nfx {
namespace dev {

__device__ __forceinline__ double2 ld_stream(const double2* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p));
    return r;
}
// 256-bit global load (sm_100+): 4 doubles per instruction, with the L2 evict-first hint that only
// this width accepts -- the u/v stream is read exactly once
struct __align__(32) double4x {
    double x, y, z, w;
};
struct __align__(32) float8x {
    float a[8];
};
__device__ __forceinline__ double4x ld_stream(const double4x* p) {
    double4x r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.f64 {%0, %1, %2, %3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float8x ld_stream(const float8x* p) {
    float8x r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.a[0]), "=f"(r.a[1]), "=f"(r.a[2]), "=f"(r.a[3]), "=f"(r.a[4]), "=f"(r.a[5]), "=f"(r.a[6]),
                   "=f"(r.a[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ double ld_stream(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}

// the same loads with an explicit L2 evict-first policy (createpolicy): narrower than 256 bits the instruction
// takes the hint only through .L2::cache_hint.  The u/v stream is read once: it must not displace the edge-flux
// ring (evict-last) or the CSR lists in L2.
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ double2 ld_stream(const double2* p, uint64_t pol) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
                 : "=d"(r.x), "=d"(r.y)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ld_stream(const double* p, uint64_t pol) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ld_stream(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ld_stream(const float* p, uint64_t pol) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double4x ld_stream(const double4x* p, uint64_t) { return ld_stream(p); }
__device__ __forceinline__ float8x ld_stream(const float8x* p, uint64_t) { return ld_stream(p); }

// Register fence: every load of a batch is issued (volatile asm keeps program order) before the
// first value is consumed, so UNROLL*2 wide loads are in flight per thread instead of one.
__device__ __forceinline__ void pin(double& x) { asm volatile("" : "+d"(x)); }
__device__ __forceinline__ void pin(float& x) { asm volatile("" : "+f"(x)); }
__device__ __forceinline__ void pin(double2& r) { asm volatile("" : "+d"(r.x), "+d"(r.y)); }
__device__ __forceinline__ void pin(float4& r) { asm volatile("" : "+f"(r.x), "+f"(r.y), "+f"(r.z), "+f"(r.w)); }
__device__ __forceinline__ void pin(double4x& r) { asm volatile("" : "+d"(r.x), "+d"(r.y), "+d"(r.z), "+d"(r.w)); }
__device__ __forceinline__ void pin(float8x& r) {
    asm volatile(""
                 : "+f"(r.a[0]), "+f"(r.a[1]), "+f"(r.a[2]), "+f"(r.a[3]), "+f"(r.a[4]), "+f"(r.a[5]), "+f"(r.a[6]),
                   "+f"(r.a[7]));
}

// eflux is written once and read later by K3: streaming (evict-first) stores.  Measured with
// tools/readbw.cu: default write-back stores cost 10 % of the read stream (DRAM read/write turnarounds for
// 1.3 % of the traffic), evict-first stores 5 %.
// keep_l2: the caller re-reads eflux right away from a small ring that it re-uses (nfx_flux_series fast path):
// evict-last keeps the dirty lines in L2, they are overwritten there and never reach DRAM (7 439 GB/s in
// tools/readbw.cu, the no-store ceiling is 7 470).
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b, int keep_l2 = 0, uint64_t pol = 0) {
    if (keep_l2)
        asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(a), "d"(b), "l"(pol) : "memory");
    else
        __stcs(reinterpret_cast<double2*>(p), make_double2(a, b));
}
__device__ __forceinline__ void st_stream1(double* p, double a, int keep_l2 = 0, uint64_t pol = 0) {
    if (keep_l2)
        asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(a), "l"(pol) : "memory");
    else
        __stcs(p, a);
}

template <typename T, int VEC>
struct Pack;
template <>
struct Pack<double, 2> {
    using type = double2;
    __device__ static void unpack(const double2& p, double (&o)[2]) {
        o[0] = p.x;
        o[1] = p.y;
    }
};
template <>
struct Pack<double, 4> {
    using type = double4x;
    __device__ static void unpack(const double4x& p, double (&o)[4]) {
        o[0] = p.x;
        o[1] = p.y;
        o[2] = p.z;
        o[3] = p.w;
    }
};
template <>
struct Pack<float, 8> {
    using type = float8x;
    __device__ static void unpack(const float8x& p, float (&o)[8]) {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = p.a[i];
    }
};
template <>
struct Pack<double, 1> {
    using type = double;
    __device__ static void unpack(const double& p, double (&o)[1]) { o[0] = p; }
};
template <>
struct Pack<float, 4> {
    using type = float4;
    __device__ static void unpack(const float4& p, float (&o)[4]) {
        o[0] = p.x;
        o[1] = p.y;
        o[2] = p.z;
        o[3] = p.w;
    }
};
template <>
struct Pack<float, 1> {
    using type = float;
    __device__ static void unpack(const float& p, float (&o)[1]) { o[0] = p; }
};

// land / missing values count as zero: NaN (xarray-decoded _FillValue, field.py:157) or == fill
template <typename T>
__device__ __forceinline__ double clean(T x, T fill, bool has_fill) {
    const bool bad = (x != x) || (has_fill && x == fill);
    return bad ? 0.0 : (double)x;
}

// float32 storage: (double)x is an F2F.F64.F32 on the XU pipe, which saturates at ~5.2 TB/s of input (ncu:
// profiles/r1_k2_f32_ncu.md).  clean_scaled() avoids the conversion: the float's bits, shifted into the double
// layout WITHOUT re-biasing the exponent (sign | exponent | mantissa << 29), are exactly x * 2^-896 for every
// finite float -- normals, denormals (they become double denormals with the same value) and +-0.  The caller
// multiplies by dz[k] * 2^896 instead of dz[k]: both factors are scaled exactly, so the rounded product has the
// same bits as dz[k] * (double)x.  NaN and the missing-value marker give 0 as in clean().  Infinities have no
// image (exponent 255 would read as a finite number): they contribute 0 here and are recorded in `amax`, and a
// thread that saw one recomputes its columns with clean() (see k2_edgeflux_ldg).  Three integer operations
// replace the F2F.
constexpr double kScaleUp = 0x1p896;      // dz[k] is multiplied by this
constexpr double kScaleLimit = 0x1p100;   // |dz[k]| below this: dz[k] * 2^896 cannot overflow

__device__ __forceinline__ double clean_scaled(float x, float fill, bool has_fill, float& amax) {
    const float ax = fabsf(x);
    amax = fmaxf(amax, ax);                                        // NaN is ignored by fmaxf, inf sticks
    // without a marker the caller passes fill = NaN, and x != NaN holds for every x: no has_fill test needed.
    // One chained predicate (FSETP ... AND P) and one select; the compiler's own code had two of each.
    uint32_t b;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.neu.f32 p, %1, %2;\n\t"
        "setp.le.and.f32 p, %3, 0f7F7FFFFF, p;\n\t"
        "selp.b32 %0, %1, 0, p;\n\t}"
        : "=r"(b)
        : "r"(__float_as_uint(x)), "r"(__float_as_uint(fill)), "r"(__float_as_uint(ax)));
    const uint32_t hi = (uint32_t)((int32_t)b >> 3) & 0x8fffffffu;
    const uint32_t lo = b << 29;
    return __hiloint2double((int)hi, (int)lo);
}
__device__ __forceinline__ double clean_scaled(double x, double fill, bool has_fill, float&) {
    return clean<double>(x, fill, has_fill);                       // never used: ALU_MASK is 0 for float64 storage
}

}  // namespace dev
}  // namespace nfx
