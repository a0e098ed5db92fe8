// arith.cuh -- per-atom arithmetic contract of the B200 SEEKR2 step kernels (sm_100a).
//
// Every floating-point operation of the Langevin update is spelled with an explicit rounding
// intrinsic, in the order nvcc emits for the reference kernels
// (seekr2plugin/platforms/cuda/src/kernels/langevin.cu:13-27 and :71-109, read from their PTX):
//
//   kick :  sq = SQRT(w); fw = fs*w; b = ns*sq
//           v' = fma(b, (mixed)R, fma(vs, v, fw*(mixed)F));   delta = dt*v'
//   drift:  X = (x + corr) + delta;  v = (mixed)(inv_dt*(double)delta)
//           x = (real)X;  corr = (real)(X - (mixed)x)
//
// so the results are bit-identical to the reference kernels and to the CPU oracle in all three
// CudaPrecision modes, independent of compiler contraction (-fmad is irrelevant here).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace s2 {

enum { SINGLE = 0, MIXED = 1, DOUBLE = 2 };

template <int P> struct Prec;
template <> struct Prec<SINGLE> {
    using real = float;  using mixed = float;  using real4 = float4;  using mixed4 = float4;
    static constexpr bool has_corr = false;
};
template <> struct Prec<MIXED> {
    using real = float;  using mixed = double; using real4 = float4;  using mixed4 = double4;
    static constexpr bool has_corr = true;
};
template <> struct Prec<DOUBLE> {
    using real = double; using mixed = double; using real4 = double4; using mixed4 = double4;
    static constexpr bool has_corr = false;
};

// ---- memory access: 128-bit and 256-bit vector transactions ------------------------------------
// In-place arrays (posq, posqCorrection, velm) are read once and written once by the owning
// thread; they bypass L1 allocation but do not use the non-coherent path.  Read-only inputs
// (force, random) use ld.global.nc.

__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double4 ld_stream(const double4* p) {   // LDG.E.256 on sm_100a
    double4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_ro(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double4 ld_ro(const double4* p) {
    double4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ long long ld_ro(const long long* p) {
    long long r;
    asm volatile("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(r) : "l"(p));
    return r;
}
// ---- L2-pinned accesses of the decide blocks ------------------------------------------------------
// The few KB a decide block needs every step (slot metadata, slab entries, the group atoms' force
// words) are tagged evict_last so the ~250 MB that stream through L2 per launch do not push them
// out: the latency-critical gathers then hit L2 instead of queueing at HBM.
__device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ int ld_keep_i32(const int* p, unsigned long long pol) {
    int r;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double ld_keep_f64(const double* p, unsigned long long pol) {
    double r;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ long long ld_keep_s64(const long long* p, unsigned long long pol) {
    long long r;
    asm volatile("ld.global.nc.L2::cache_hint.s64 %0, [%1], %2;" : "=l"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float4 ld_keep(const float4* p, unsigned long long pol) {
    float4 r;
    asm volatile("ld.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ double4 ld_keep(const double4* p, unsigned long long pol) {
    double4 r;
    asm volatile("ld.global.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                 : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void st_keep(float4* p, const float4& v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep(double4* p, const double4& v, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.f64 [%0], {%1,%2,%3,%4}, %5;"
                 :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(double4* p, const double4& v) {   // STG.E.256
    asm volatile("st.global.L1::no_allocate.v4.f64 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

// ---- rounding-explicit scalar ops ----------------------------------------------------------------
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float  sub_rn(float a, float b)   { return __fsub_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float  fma_rn(float a, float b, float c)    { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_rn(double a, double b, double c) { return __fma_rn(a, b, c); }

template <class M> __device__ __forceinline__ M from_ll(long long v);
template <> __device__ __forceinline__ float  from_ll<float>(long long v)  { return __ll2float_rn(v); }
template <> __device__ __forceinline__ double from_ll<double>(long long v) { return __ll2double_rn(v); }

// SQRT of the OpenMM prelude: sqrtf unless full double precision (langevin.cu:22)
__device__ __forceinline__ float sqrt_no_call(float x);   // kernels.cuh
template <int P> __device__ __forceinline__ typename Prec<P>::mixed sqrt_inv_mass(typename Prec<P>::mixed w) {
    if constexpr (P == DOUBLE) return __dsqrt_rn(w);
#ifdef SEEKR2_SQRT_INV_MASS_NO_CALL                              // experimental builds: the call-free square root here too (bit-identical)
    else return (typename Prec<P>::mixed) sqrt_no_call((float) w);
#else
    else return (typename Prec<P>::mixed) __fsqrt_rn((float) w);
#endif
}

// parameters of one step, prepared on the host in the precision the reference kernel computes in
template <int P> struct StepParams {
    typename Prec<P>::mixed vscale;      // params[VelScale]
    typename Prec<P>::mixed fs;          // params[ForceScale] / 2^32   (langevin.cu:14)
    typename Prec<P>::mixed noisescale;  // params[NoiseScale]
    typename Prec<P>::mixed dt;          // dt[0].y
    double inv_dt;                       // 1.0 / dt[0].y, always double (langevin.cu:71)
    // LangevinMiddle scheme (langevinMiddle.cu:11,37,73)
    typename Prec<P>::mixed fs_mid;      // dt[0].y / 2^32
    typename Prec<P>::mixed halfdt;      // 0.5f * dt[0].y
    typename Prec<P>::mixed inv_dt_m;    // 1 / dt[0].y in MIXED precision
};

template <int P> struct Vec3M { typename Prec<P>::mixed x, y, z; };

// langevin.cu:19-27.  v holds the start-of-step velocity on entry and the kicked one on exit.
template <int P>
__device__ __forceinline__ void kick(const StepParams<P>& p, typename Prec<P>::mixed4& v, long long fx,
                                     long long fy, long long fz, const float4& r, Vec3M<P>& delta) {
    using M = typename Prec<P>::mixed;
    const M sq = sqrt_inv_mass<P>(v.w);
    const M fw = mul_rn(p.fs, v.w);
    const M b = mul_rn(p.noisescale, sq);
    v.x = fma_rn(b, (M) r.x, fma_rn(p.vscale, v.x, mul_rn(fw, from_ll<M>(fx))));
    v.y = fma_rn(b, (M) r.y, fma_rn(p.vscale, v.y, mul_rn(fw, from_ll<M>(fy))));
    v.z = fma_rn(b, (M) r.z, fma_rn(p.vscale, v.z, mul_rn(fw, from_ll<M>(fz))));
    delta.x = mul_rn(p.dt, v.x);
    delta.y = mul_rn(p.dt, v.y);
    delta.z = mul_rn(p.dt, v.z);
}

// langevin.cu:81-109.  x, c: position and correction before the step; xn, cn after it; v gets
// delta/dt.  c / cn are ignored unless P == MIXED.
template <int P>
__device__ __forceinline__ void drift(const StepParams<P>& p, const typename Prec<P>::real4& x,
                                      const typename Prec<P>::real4& c, const Vec3M<P>& d,
                                      typename Prec<P>::real4& xn, typename Prec<P>::real4& cn,
                                      typename Prec<P>::mixed4& v) {
    if constexpr (P == MIXED) {
        const double X = __dadd_rn(__dadd_rn((double) x.x, (double) c.x), d.x);
        const double Y = __dadd_rn(__dadd_rn((double) x.y, (double) c.y), d.y);
        const double Z = __dadd_rn(__dadd_rn((double) x.z, (double) c.z), d.z);
        xn.x = __double2float_rn(X); xn.y = __double2float_rn(Y); xn.z = __double2float_rn(Z); xn.w = x.w;
        cn.x = __double2float_rn(__dsub_rn(X, (double) xn.x));
        cn.y = __double2float_rn(__dsub_rn(Y, (double) xn.y));
        cn.z = __double2float_rn(__dsub_rn(Z, (double) xn.z));
        cn.w = 0.0f;
        v.x = __dmul_rn(p.inv_dt, d.x); v.y = __dmul_rn(p.inv_dt, d.y); v.z = __dmul_rn(p.inv_dt, d.z);
    } else if constexpr (P == SINGLE) {
        xn.x = __fadd_rn(x.x, d.x); xn.y = __fadd_rn(x.y, d.y); xn.z = __fadd_rn(x.z, d.z); xn.w = x.w;
        cn = c;
        v.x = __double2float_rn(__dmul_rn(p.inv_dt, (double) d.x));
        v.y = __double2float_rn(__dmul_rn(p.inv_dt, (double) d.y));
        v.z = __double2float_rn(__dmul_rn(p.inv_dt, (double) d.z));
    } else {
        xn.x = __dadd_rn(x.x, d.x); xn.y = __dadd_rn(x.y, d.y); xn.z = __dadd_rn(x.z, d.z); xn.w = x.w;
        cn = c;
        v.x = __dmul_rn(p.inv_dt, d.x); v.y = __dmul_rn(p.inv_dt, d.y); v.z = __dmul_rn(p.inv_dt, d.z);
    }
}

// ---- LangevinMiddle scheme (langevinMiddle.cu), fma placement as nvcc emits for the reference ------
// Part1 (:7-21): v1 = fma(fs*w, (mixed)F, v)
template <int P>
__device__ __forceinline__ void middle_kick(const StepParams<P>& p, typename Prec<P>::mixed4& v, long long fx,
                                            long long fy, long long fz) {
    using M = typename Prec<P>::mixed;
    const M fw = mul_rn(p.fs_mid, v.w);
    v.x = fma_rn(fw, from_ll<M>(fx), v.x);
    v.y = fma_rn(fw, from_ll<M>(fy), v.y);
    v.z = fma_rn(fw, from_ll<M>(fz), v.z);
}
// Part2 (:28-61): v2 = fma(vs, v1, (ns*SQRT(w))*R); delta = fma(h, v1, h*v2).  v: v1 in, v2 out.
template <int P>
__device__ __forceinline__ void middle_ou_drift(const StepParams<P>& p, typename Prec<P>::mixed4& v, const float4& r,
                                                Vec3M<P>& delta) {
    using M = typename Prec<P>::mixed;
    const M b = mul_rn(p.noisescale, sqrt_inv_mass<P>(v.w));
    const M nx = fma_rn(p.vscale, v.x, mul_rn(b, (M) r.x));
    const M ny = fma_rn(p.vscale, v.y, mul_rn(b, (M) r.y));
    const M nz = fma_rn(p.vscale, v.z, mul_rn(b, (M) r.z));
    delta.x = fma_rn(p.halfdt, v.x, mul_rn(p.halfdt, nx));
    delta.y = fma_rn(p.halfdt, v.y, mul_rn(p.halfdt, ny));
    delta.z = fma_rn(p.halfdt, v.z, mul_rn(p.halfdt, nz));
    v.x = nx; v.y = ny; v.z = nz;
}
// Part3 (:68-101): v3 = fma(invDt, delta - oldDelta, v2); X = delta + (x + corr); split.
template <int P>
__device__ __forceinline__ void middle_finish(const StepParams<P>& p, const typename Prec<P>::real4& x,
                                              const typename Prec<P>::real4& c, const Vec3M<P>& d, const Vec3M<P>& d_old,
                                              typename Prec<P>::real4& xn, typename Prec<P>::real4& cn,
                                              typename Prec<P>::mixed4& v) {
    v.x = fma_rn(p.inv_dt_m, sub_rn(d.x, d_old.x), v.x);
    v.y = fma_rn(p.inv_dt_m, sub_rn(d.y, d_old.y), v.y);
    v.z = fma_rn(p.inv_dt_m, sub_rn(d.z, d_old.z), v.z);
    if constexpr (P == MIXED) {
        const double X = __dadd_rn(d.x, __dadd_rn((double) x.x, (double) c.x));
        const double Y = __dadd_rn(d.y, __dadd_rn((double) x.y, (double) c.y));
        const double Z = __dadd_rn(d.z, __dadd_rn((double) x.z, (double) c.z));
        xn.x = __double2float_rn(X); xn.y = __double2float_rn(Y); xn.z = __double2float_rn(Z); xn.w = x.w;
        cn.x = __double2float_rn(__dsub_rn(X, (double) xn.x));
        cn.y = __double2float_rn(__dsub_rn(Y, (double) xn.y));
        cn.z = __double2float_rn(__dsub_rn(Z, (double) xn.z));
        cn.w = 0.0f;
    } else {
        xn.x = add_rn(d.x, x.x); xn.y = add_rn(d.y, x.y); xn.z = add_rn(d.z, x.z); xn.w = x.w;
        cn = c;
    }
}

template <class V4> __device__ __forceinline__ V4 negate_xyz(const V4& v) {
    V4 r; r.x = -v.x; r.y = -v.y; r.z = -v.z; r.w = v.w; return r;
}

}  // namespace s2
