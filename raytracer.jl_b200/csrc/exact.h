// exact.h -- the exact candidate value of the relaxation, in the reference's operation order, one rounding per
// operation (round-to-nearest intrinsics on the device, plain IEEE operations on the host: build host code with
// -ffp-contract=off).  Host + device so that the expression the kernels inline is checked on the CPU against numpy
// Float64 and genuine Float32 arithmetic (tests/test_screen.py).
//
// Float32 mode (precision = 32; the reference's Float32 path src/SSSP/bfm_gpu.jl:170-205, 487-526): values stay in fp64
// registers / storage but every arithmetic result is rounded to Float32.  For +, -, *, / and sqrt of Float32 operands,
// rounding the correctly rounded fp64 result to Float32 equals the correctly rounded Float32 result (53 >= 2*24 + 2:
// double rounding is innocuous), so the travel times are bit-identical to genuine Float32 arithmetic.
#pragma once
#include <cmath>

#include "screen.h"

#if defined(__CUDA_ARCH__)
#define RT_DADD(a, b) __dadd_rn(a, b)
#define RT_DSUB(a, b) __dsub_rn(a, b)
#define RT_DMUL(a, b) __dmul_rn(a, b)
#define RT_DDIV(a, b) __ddiv_rn(a, b)
#define RT_DSQRT(a) __dsqrt_rn(a)
#define RT_DRCP(a) __drcp_rn(a)
#else
#define RT_DADD(a, b) ((a) + (b))
#define RT_DSUB(a, b) ((a) - (b))
#define RT_DMUL(a, b) ((a) * (b))
#define RT_DDIV(a, b) ((a) / (b))
#define RT_DSQRT(a) std::sqrt(a)
#define RT_DRCP(a) (1.0 / (a))
#endif

template <bool F32>
RT_HD double rnd(double v) {
  if constexpr (F32) {
#if defined(__CUDA_ARCH__)
    return (double)__double2float_rn(v);
#else
    return (double)(float)v;
#endif
  }
  return v;
}

// 2-D: d_from + (2.0 * sqrt(dx*dx + dz*dz)) / (Ui + Uj)   (src/SSSP/bfm.jl:186, src/GridAnnulus.jl:808-815;
// Float32: src/SSSP/bfm_gpu.jl:505-512).  Bitwise symmetric in (i, j).
template <bool F32>
RT_HD double exact_cand2(double d_from, double xi, double zi, double Ui, double xj, double zj, double Uj) {
  const double dx = rnd<F32>(RT_DSUB(xi, xj));
  const double dz = rnd<F32>(RT_DSUB(zi, zj));
  const double d2 = rnd<F32>(RT_DADD(rnd<F32>(RT_DMUL(dx, dx)), rnd<F32>(RT_DMUL(dz, dz))));
  const double len2 = RT_DMUL(2.0, rnd<F32>(RT_DSQRT(d2)));
  const double w = rnd<F32>(RT_DDIV(len2, rnd<F32>(RT_DADD(Ui, Uj))));
  return rnd<F32>(RT_DADD(d_from, w));
}

// 3-D, weight mode 0 (default): d_from + distance3D(pI, pJ) * (1 / abs(UI + UJ)) * 2   (src/SSSP/weights.jl:20,
// src/StructuredGrid.jl:239-241); weight mode 1: d_from + distance3D(pJ, pI) / abs(UJ + UI) * 0.5, the expression of
// the legacy 3-D solvers BFM/foo! (src/Dijsktra.jl:388) and dijsktra (:44).  The two differ in rounding (reciprocal
// then multiply against a division) and by a factor 4.  `wmode` is uniform per launch.
template <bool F32>
RT_HD double exact_cand3(double d_from, double xi, double yi, double zi, double ui, double xj, double yj, double zj,
                         double uj, int wmode = 0) {
  const double dx = rnd<F32>(RT_DSUB(xi, xj)), dy = rnd<F32>(RT_DSUB(yi, yj)), dz = rnd<F32>(RT_DSUB(zi, zj));
  const double s = rnd<F32>(RT_DADD(rnd<F32>(RT_DADD(rnd<F32>(RT_DMUL(dx, dx)), rnd<F32>(RT_DMUL(dy, dy)))),
                                    rnd<F32>(RT_DMUL(dz, dz))));
  const double d = rnd<F32>(RT_DSQRT(s));
  const double us = fabs(rnd<F32>(RT_DADD(ui, uj)));
  double wgt;
  if (wmode) {
    wgt = rnd<F32>(RT_DMUL(rnd<F32>(RT_DDIV(d, us)), 0.5));
  } else {
    const double rcp = rnd<F32>(RT_DRCP(us));
    wgt = rnd<F32>(RT_DMUL(rnd<F32>(RT_DMUL(d, rcp)), 2.0));
  }
  return rnd<F32>(RT_DADD(d_from, wgt));
}
// the screens are written for w = 2*sqrt(d2)/ssum; mode 1 is 0.5*sqrt(d2)/ssum = 2*sqrt(d2)/(4*ssum): scale ssum (exact)
RT_HD double screen_ssum3(double us_abs, int wmode) { return wmode ? 4.0 * us_abs : us_abs; }
