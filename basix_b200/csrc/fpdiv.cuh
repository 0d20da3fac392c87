// Correctly rounded division by a divisor whose reciprocal is known (shared by geometry.cu and interpolate.cu).
#pragma once

namespace bx
{
namespace fpdiv
{
// a / b given rb = RN(1 / b): two FMA correction steps on a * rb.  The residual a - q b of a quotient that is within
// an ulp is exact in one FMA, and adding residual * rb rounds to the correctly rounded quotient (Markstein); the first
// step brings a * rb (up to 2 ulp off) within an ulp, the second lands on RN(a / b) -- the value the reference's
// `acc / detJ` (maps.h:122-123) has.  tools/div_check.cu compares the two over 2 * 10^10 random and adversarial pairs
// per type (profiles/r02_div_check.json: 0 mismatches).  When the quotient leaves a wide exponent window (zeros,
// infinities, NaNs -- the caller passes rb = NaN for a divisor outside its own window) the IEEE division decides.
// (out of line on purpose: inlined, the compiler turns the rare branch into an unconditional division and a select)
static __device__ __noinline__ double ieee_div(double a, double b) { return a / b; }
static __device__ __noinline__ float ieee_div(float a, float b) { return a / b; }

__device__ __forceinline__ double div_core(double a, double b, double rb)
{
  double q = __dmul_rn(a, rb);
  double r = fma(-q, b, a);
  q = fma(r, rb, q);
  r = fma(-q, b, a);
  return fma(r, rb, q);
}
__device__ __forceinline__ float div_core(float a, float b, float rb)
{
  float q = __fmul_rn(a, rb);
  float r = fmaf(-q, b, a);
  q = fmaf(r, rb, q);
  r = fmaf(-q, b, a);
  return fmaf(r, rb, q);
}
// quotient outside the window in which div_core is trusted: biased exponent 1023 +- 500 (float: 127 +- 60)
__device__ __forceinline__ bool outside_window(double q)
{
  return ((unsigned(__double2hiint(q)) >> 20) & 0x7ffu) - 523u > 1000u;
}
__device__ __forceinline__ bool outside_window(float q) { return ((__float_as_uint(q) >> 23) & 0xffu) - 67u > 120u; }
// N quotients by the same divisor; one test (and one rarely taken branch) for all of them
template <typename T, int N>
__device__ __forceinline__ void div_by(T (&a)[N], T b, T rb)
{
  T q[N];
  bool bad = false;
#pragma unroll
  for (int i = 0; i < N; ++i)
  {
    q[i] = div_core(a[i], b, rb);
    bad = bad or outside_window(q[i]);
  }
  if (bad)
  {
#pragma unroll
    for (int i = 0; i < N; ++i)
      q[i] = ieee_div(a[i], b);
  }
#pragma unroll
  for (int i = 0; i < N; ++i)
    a[i] = q[i];
}
// reciprocal for div_by: NaN unless the divisor is finite and within 2^+-300 (float: 2^+-40), so that every product and
// residual in div_by stays in range whenever its quotient test passes
__device__ __forceinline__ double safe_reciprocal(double b)
{
  const unsigned ex = (unsigned(__double2hiint(b)) >> 20) & 0x7ffu;
  return ex - 723u > 600u ? __longlong_as_double(0x7ff8000000000000ll) : 1.0 / b;
}
__device__ __forceinline__ float safe_reciprocal(float b)
{
  const unsigned ex = (__float_as_uint(b) >> 23) & 0xffu;
  return ex - 87u > 80u ? __uint_as_float(0x7fc00000u) : 1.0f / b;
}
} // namespace fpdiv
} // namespace bx
