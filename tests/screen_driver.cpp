// screen_driver.cpp -- TEST INFRASTRUCTURE: evaluates the algebraic screen (raytracer.jl_b200/csrc/screen.h) and the exact
// candidate expression (exact.h) -- the very functions the kernels inline -- next to the candidate value in the
// reference's operation order computed in plain `double` / genuine `float` arithmetic, so that tests/test_screen.py can
// check on the CPU that a "skip" can never change a result and that the fp64-with-rounding emulation of Float32 is
// exact.  Build with -ffp-contract=off.
#include <cmath>

#include "../raytracer.jl_b200/csrc/exact.h"  // pulls in screen.h

namespace {
// exact candidate, 2-D: dj + (2*sqrt(dx^2 + dz^2)) / (Ui + Uj)   (bfm.jl:186); T = float is genuine Float32 arithmetic
template <typename T>
T cand2(T dj, T xi, T zi, T Ui, T xj, T zj, T Uj) {
  T dx = xi - xj, dz = zi - zj;
  T d2 = dx * dx + dz * dz;
  T len2 = T(2) * std::sqrt(d2);
  T w = len2 / (Ui + Uj);
  return dj + w;
}
// exact candidate, 3-D: dj + sqrt(dx^2 + dy^2 + dz^2) * (1 / |Ui + Uj|) * 2   (weights.jl:20)
template <typename T>
T cand3(T dj, T xi, T yi, T zi, T Ui, T xj, T yj, T zj, T Uj) {
  T dx = xi - xj, dy = yi - yj, dz = zi - zj;
  T d = std::sqrt(dx * dx + dy * dy + dz * dz);
  T w = d * (T(1) / std::fabs(Ui + Uj)) * T(2);
  return dj + w;
}
}  // namespace

extern "C" {

// in: 9 arrays of n doubles (Float32 mode: values must be Float32 numbers); out: skip flags, tight flags (for
// target[i]), exact candidate values
void screen2d_batch(long n, int f32, const double* bound, const double* dj, const double* xi, const double* zi,
                    const double* Ui, const double* xj, const double* zj, const double* Uj, const double* target,
                    unsigned char* skip, unsigned char* maybe_tight, double* delta, double* delta_kernel) {
  for (long i = 0; i < n; ++i) {
    delta_kernel[i] = f32 ? exact_cand2<true>(dj[i], xi[i], zi[i], Ui[i], xj[i], zj[i], Uj[i])
                          : exact_cand2<false>(dj[i], xi[i], zi[i], Ui[i], xj[i], zj[i], Uj[i]);
    // what the kernels pass: fp64 differences of the (possibly Float32-valued) inputs, FMA'd square sum
    const double dx = xi[i] - xj[i], dz = zi[i] - zj[i];
    const double d2 = rt_fma(dx, dx, dz * dz);
    const double ssum = Ui[i] + Uj[i];
    if (f32) {
      skip[i] = screen_cannot_improve_t<true>(bound[i], dj[i], d2, ssum);
      maybe_tight[i] = screen_maybe_tight_t<true>(target[i], dj[i], d2, ssum);
      delta[i] = (double)cand2<float>((float)dj[i], (float)xi[i], (float)zi[i], (float)Ui[i], (float)xj[i],
                                      (float)zj[i], (float)Uj[i]);
    } else {
      skip[i] = screen_cannot_improve_t<false>(bound[i], dj[i], d2, ssum);
      maybe_tight[i] = screen_maybe_tight_t<false>(target[i], dj[i], d2, ssum);
      delta[i] = cand2<double>(dj[i], xi[i], zi[i], Ui[i], xj[i], zj[i], Uj[i]);
    }
  }
}

void screen3d_batch(long n, int f32, const double* bound, const double* dj, const double* xi, const double* yi,
                    const double* zi, const double* Ui, const double* xj, const double* yj, const double* zj,
                    const double* Uj, const double* target, unsigned char* skip, unsigned char* maybe_tight,
                    double* delta, double* delta_kernel) {
  for (long i = 0; i < n; ++i) {
    delta_kernel[i] = f32 ? exact_cand3<true>(dj[i], xi[i], yi[i], zi[i], Ui[i], xj[i], yj[i], zj[i], Uj[i])
                          : exact_cand3<false>(dj[i], xi[i], yi[i], zi[i], Ui[i], xj[i], yj[i], zj[i], Uj[i]);
    const double dx = xi[i] - xj[i], dy = yi[i] - yj[i], dz = zi[i] - zj[i];
    const double d2 = rt_fma(dx, dx, rt_fma(dy, dy, dz * dz));
    const double ssum = std::fabs(Ui[i] + Uj[i]);
    if (f32) {
      skip[i] = screen_cannot_improve_t<true>(bound[i], dj[i], d2, ssum);
      maybe_tight[i] = screen_maybe_tight_t<true>(target[i], dj[i], d2, ssum);
      delta[i] = (double)cand3<float>((float)dj[i], (float)xi[i], (float)yi[i], (float)zi[i], (float)Ui[i],
                                      (float)xj[i], (float)yj[i], (float)zj[i], (float)Uj[i]);
    } else {
      skip[i] = screen_cannot_improve_t<false>(bound[i], dj[i], d2, ssum);
      maybe_tight[i] = screen_maybe_tight_t<false>(target[i], dj[i], d2, ssum);
      delta[i] = cand3<double>(dj[i], xi[i], yi[i], zi[i], Ui[i], xj[i], yj[i], zj[i], Uj[i]);
    }
  }
}

// group screen (screen.h: group_cannot_improve_t): ng groups of ns sources each, one target per group.  Outputs the
// screen's answer and the smallest exact candidate over the group's sources (kernel arithmetic).
void group2d_batch(long ng, int ns, int f32, const double* bound, const double* ds, const double* xs, const double* zs,
                   const double* Us, const double* xt, const double* zt, const double* Ut, unsigned char* skip,
                   double* best_exact) {
  for (long g = 0; g < ng; ++g) {
    double dmin = INFINITY, xmin = INFINITY, xmax = -INFINITY, zmin = INFINITY, zmax = -INFINITY, umax = 0.0;
    double be = INFINITY;
    for (int k = 0; k < ns; ++k) {
      const long q = g * ns + k;
      dmin = std::fmin(dmin, ds[q]);
      xmin = std::fmin(xmin, xs[q]);
      xmax = std::fmax(xmax, xs[q]);
      zmin = std::fmin(zmin, zs[q]);
      zmax = std::fmax(zmax, zs[q]);
      umax = std::fmax(umax, Us[q]);
      const double e = f32 ? exact_cand2<true>(ds[q], xs[q], zs[q], Us[q], xt[g], zt[g], Ut[g])
                           : exact_cand2<false>(ds[q], xs[q], zs[q], Us[q], xt[g], zt[g], Ut[g]);
      be = std::fmin(be, e);
    }
    // exactly what the kernels compute per released item
    const double cx = 0.5 * (xmin + xmax), cz = 0.5 * (zmin + zmax);
    const double hx = xmax - xmin, hz = zmax - zmin;
    const double rho = 0.5 * std::sqrt(hx * hx + hz * hz) * (1.0 + 1e-12) + 1e-300;
    const double dx = xt[g] - cx, dz = zt[g] - cz;
    const double d2c = rt_fma(dx, dx, dz * dz);
    skip[g] = dmin < bound[g] && (f32 ? group_cannot_improve_t<true>(bound[g], dmin, d2c, rho, Ut[g] + umax)
                                      : group_cannot_improve_t<false>(bound[g], dmin, d2c, rho, Ut[g] + umax));
    best_exact[g] = be;
  }
}
}
