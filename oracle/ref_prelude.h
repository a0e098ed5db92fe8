/*
 * oracle/ref_prelude.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * The reference's CUDA kernels (seekr2plugin/platforms/cuda/src/kernels/langevin.cu) are not
 * self-contained: OpenMM's CudaContext::createModule() prepends a block of typedefs and macros
 * (real/mixed vector types, make_* constructors, SQRT ...) before handing the source to NVRTC.
 * OpenMM is absent from this environment, so this header restates that injected prelude for the
 * three CudaPrecision modes.  It contains no reference code; the reference .cu is compiled from
 * where it lies under /root/reference by oracle/build_ref.sh.
 *
 * Precision is selected by exactly one of:
 *   -DREF_PRECISION_SINGLE   real=float  mixed=float
 *   -DREF_PRECISION_MIXED    real=float  mixed=double   (+ USE_MIXED_PRECISION, as OpenMM defines)
 *   -DREF_PRECISION_DOUBLE   real=double mixed=double   (+ USE_DOUBLE_PRECISION)
 * SQRT follows OpenMM: sqrtf unless full double precision.
 */
#pragma once
#include <cuda_runtime.h>

#if defined(REF_PRECISION_DOUBLE)
#define USE_DOUBLE_PRECISION 1
typedef double real;
typedef double2 real2;
typedef double3 real3;
typedef double4 real4;
typedef double mixed;
typedef double2 mixed2;
typedef double3 mixed3;
typedef double4 mixed4;
#define make_real2 make_double2
#define make_real3 make_double3
#define make_real4 make_double4
#define make_mixed2 make_double2
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#define SQRT sqrt
#elif defined(REF_PRECISION_MIXED)
#define USE_MIXED_PRECISION 1
typedef float real;
typedef float2 real2;
typedef float3 real3;
typedef float4 real4;
typedef double mixed;
typedef double2 mixed2;
typedef double3 mixed3;
typedef double4 mixed4;
#define make_real2 make_float2
#define make_real3 make_float3
#define make_real4 make_float4
#define make_mixed2 make_double2
#define make_mixed3 make_double3
#define make_mixed4 make_double4
#define SQRT sqrtf
#elif defined(REF_PRECISION_SINGLE)
typedef float real;
typedef float2 real2;
typedef float3 real3;
typedef float4 real4;
typedef float mixed;
typedef float2 mixed2;
typedef float3 mixed3;
typedef float4 mixed4;
#define make_real2 make_float2
#define make_real3 make_float3
#define make_real4 make_float4
#define make_mixed2 make_float2
#define make_mixed3 make_float3
#define make_mixed4 make_float4
#define SQRT sqrtf
#else
#error "define one of REF_PRECISION_SINGLE / REF_PRECISION_MIXED / REF_PRECISION_DOUBLE"
#endif
