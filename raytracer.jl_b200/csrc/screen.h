// screen.h -- the algebraic screen every relax / push / tightness kernel runs before the exact candidate evaluation.
// Host + device so that its safety (a "skip" never changes a result) is property-tested on the CPU against the exact
// arithmetic (tests/test_screen.py); the kernels inline the very same functions.
//
// Is  d_from + w >= bound  guaranteed, where the edge weight is w = 2*sqrt(d2)/ssum (2-D: (2.0*len)/(Ui+Uj); 3-D:
// len*(1/|Ui+Uj|)*2, both within 3 roundings of the real value)?   w >= t  <=>  d2 >= (t*ssum/2)^2, t = bound - d_from.
// Slack: 4e-15*bound absorbs the rounding of (bound - d_from) and of the final fl(d_from + w); the factor (1 + 1e-9)
// absorbs the roundings of the products and of w itself.  Float32 mode (every operation of the exact path rounded to
// Float32): the final fl32(d_from + w) moves by <= 6e-8 relative and the Float32 weight differs from the real one by
// < 5 roundings of 6e-8 (d2 and ssum passed here are fp64 values computed from the Float32 inputs), hence 1.3e-7 and
// (1 + 2e-6).  A `true` answer is exact-safe: the candidate can be skipped without changing any result; `false` means
// "evaluate exactly".  The screen is free to contract (explicit FMAs): it has to be conservative, not bit-reproducible.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define RT_HD __host__ __device__ __forceinline__
#else
#define RT_HD inline
#endif

RT_HD double rt_fma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// Callers guarantee d_from < bound (so t > 0; bound = Inf gives ts = Inf and never skips).
template <bool F32>
RT_HD bool screen_cannot_improve_t(double bound, double d_from, double d2, double ssum) {
  const double tt = rt_fma(bound, F32 ? 1.3e-7 : 4e-15, bound - d_from);
  const double ts = tt * ssum;
  return (d2 > ts * ts * (F32 ? 0.25 * (1.0 + 2e-6) : 0.25 * (1.0 + 1e-9))) && (ssum > 0.0);
}
RT_HD bool screen_cannot_improve(double bound, double d_from, double d2, double ssum) {
  return screen_cannot_improve_t<false>(bound, d_from, d2, ssum);
}

// Group screen: can ANY source of a released group improve a target t whose incumbent is `bound`?  The group's sources
// lie in the disc (centre c, radius rho), their travel times are >= dmin (callers guarantee dmin < bound) and their
// velocities <= umax, so for every source s:  d_s + w(s, t) >= dmin + 2 (|t - c| - rho) / (u_t + umax).  That lower bound
// is run through the same inequality as screen_cannot_improve_t, without the square root:
//   |t - c| >= rho + (bound - dmin + slack) (u_t + umax) / 2   =>   no source can improve t.
// d2c = |t - c|^2, ssum = u_t + umax (dual velocity: the larger of the two values on both sides).  `true` is exact-safe.
template <bool F32>
RT_HD bool group_cannot_improve_t(double bound, double dmin, double d2c, double rho, double ssum) {
  const double tt = rt_fma(bound, F32 ? 1.3e-7 : 4e-15, bound - dmin);
  const double rhs = rt_fma(tt * ssum, F32 ? 0.5 * (1.0 + 1e-6) : 0.5 * (1.0 + 1e-9), rho);
  return (d2c > rhs * rhs * (1.0 + 1e-12)) && (ssum > 0.0);
}

// Can fl(d_from + w) == target hold?  false => certainly not tight.
template <bool F32>
RT_HD bool screen_maybe_tight_t(double target, double d_from, double d2, double ssum) {
  const double t = target - d_from;
  if (!(t >= 0.0)) return false;
  if (!(ssum > 0.0)) return true;
  const double slack = target * (F32 ? 1.3e-7 : 4e-15);
  const double hi = (t + slack) * ssum * 0.5;
  if (d2 > hi * hi * (F32 ? 1.0 + 2e-6 : 1.0 + 1e-9)) return false;
  const double tl = t - slack;
  if (tl > 0.0) {
    const double lo = tl * ssum * 0.5;
    if (d2 < lo * lo * (F32 ? 1.0 - 2e-6 : 1.0 - 1e-9)) return false;
  }
  return true;
}
RT_HD bool screen_maybe_tight(double target, double d_from, double d2, double ssum) {
  return screen_maybe_tight_t<false>(target, d_from, d2, ssum);
}
