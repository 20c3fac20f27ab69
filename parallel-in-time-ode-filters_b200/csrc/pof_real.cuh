// Scalar type of the register-resident kernel family.  The same sources are compiled twice: fp64 (namespace pof, the
// default and the parity path) and, with POF_F32 defined, fp32 (namespace pof32, C ABI entry points *_f32) -- the
// optional reduced-precision mode, reported separately (BASELINE.json north_star).  The large-state tile family, the
// one-thread sequential EKS and the host simulator are fp64 only.
#pragma once
#include <cuda_runtime.h>

#ifdef POF_F32
#define POF_NS pof32
#define POF_SUFFIX(name) name##_f32
namespace pof32 {
typedef float real;
typedef float2 real2;
__host__ __device__ __forceinline__ real2 make_real2(real a, real b) { return make_float2(a, b); }
}  // namespace pof32
#else
#define POF_NS pof
#define POF_SUFFIX(name) name##_f64
namespace pof {
typedef double real;
typedef double2 real2;
__host__ __device__ __forceinline__ real2 make_real2(real a, real b) { return make_double2(a, b); }
}  // namespace pof
#endif

namespace POF_NS {
constexpr bool REAL_IS_F64 = sizeof(real) == 8;
// reciprocal / reciprocal square root from the hardware approximation plus Newton steps (the IEEE division / sqrt
// subroutines cost registers and a long dependent chain inside the fully unrolled Householder sweeps)
__device__ __forceinline__ real fast_rcp(real x) {
  if constexpr (REAL_IS_F64) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"((double)x));
    double e = fma(-(double)x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-(double)x, r, 1.0);
    return (real)fma(r, e, r);
  } else {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)x));
    const float e = fmaf(-(float)x, r, 1.0f);
    return (real)fmaf(r, e, r);
  }
}
__device__ __forceinline__ real fast_rsqrt(real x) {
  if constexpr (REAL_IS_F64) {
    double r;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"((double)x));
    const double hx = 0.5 * (double)x;
    double e = fma(-hx * r, r, 0.5);
    r = fma(r, e, r);
    e = fma(-hx * r, r, 0.5);
    return (real)fma(r, e, r);
  } else {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((float)x));
    const float hx = 0.5f * (float)x;
    const float e = fmaf(-hx * r, r, 0.5f);
    return (real)fmaf(r, e, r);
  }
}
// squared norms below this are treated as zero by the Householder generator: rsqrt.approx.ftz flushes subnormal
// inputs to zero (-> inf -> NaN in the Newton step)
__device__ __forceinline__ real tiny_norm2() { return REAL_IS_F64 ? (real)0x1p-1000 : (real)0x1p-100f; }
}  // namespace POF_NS
