// tor_detmath.h — deterministic float64 sin/cos/pow, bit-identical on host (gcc) and
// device (nvcc, sm_100a).
//
// Why this exists: the reference's only external arithmetic is libm — sin/cos in
// sampling.nim:51-55 (`random(UnitVector)`), pow in physics/materials.nim:60 (`schlick`)
// and primitives/canvas.nim:52-54 (`draw`, gamma).  glibc and CUDA libm both stay below
// 1-2 ulp but are not bit-identical to each other (glibc even selects FMA / non-FMA
// variants at run time), so a GPU image can never be bit-compared with a libm-based CPU
// image.  These routines use nothing but IEEE-754 correctly rounded +, -, *, /, fma and
// integer bit manipulation, so the same source gives the same bits on both sides.  The
// GPU kernels always use them; the CPU oracle can be built with them (parity gate: f64
// framebuffer bit-exact) or with glibc (parity gate: 8-bit image).
//
// Accuracy (measured against mpmath in tests/test_detmath.py):
//   sincos : < 1 ulp on [0, 2*pi] (Cody-Waite reduction by pi/2 with an exact fma +
//            degree-13/14 minimax kernels that carry the reduction tail)
//   pow    : correctly rounded in all sampled cases (double-double log / exp, ~2^-80
//            relative before the final rounding); x^5 takes a double-double product path.
//
// Must be compiled WITHOUT floating-point contraction (gcc -ffp-contract=off, nvcc
// -fmad=false): every fma below is explicit.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define TOR_HD __host__ __device__ __forceinline__
#define TOR_HD_NOINLINE __host__ __device__ __noinline__
#else
#define TOR_HD inline
#define TOR_HD_NOINLINE inline
#endif

namespace tor {
namespace detmath {

TOR_HD double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return __builtin_fma(a, b, c);
#endif
}

TOR_HD uint64_t bits_of(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}

TOR_HD double from_bits(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

TOR_HD double floor_(double x) {
#if defined(__CUDA_ARCH__)
  return floor(x);
#else
  return __builtin_floor(x);
#endif
}

TOR_HD double sqrt_(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return __builtin_sqrt(x);
#endif
}

// ---------------------------------------------------------------------------------------
// sin / cos
// ---------------------------------------------------------------------------------------

// Polynomial kernels on |x| <= ~pi/4 with a tail y (x + y is the reduced argument).
// Coefficients are the classic degree-13 (sin) / degree-14 (cos) minimax sets.
TOR_HD double kernel_sin(double x, double y) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = x * x;
  double v = z * x;
  double r = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return x - ((z * (0.5 * y - v * r) - y) - v * S1);
}

TOR_HD double kernel_cos(double x, double y) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = x * x;
  double r = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  double ax = x < 0.0 ? -x : x;
  if (ax < 0.3) return 1.0 - (0.5 * z - (z * r - x * y));
  // split 1 - z/2 as (1 - qx) - (z/2 - qx) with qx ~ |x|/4 exactly representable
  double qx;
  if (ax > 0.78125) {
    qx = 0.28125;
  } else {
    uint64_t hi = (bits_of(ax) >> 32) - 0x00200000ull;  // exponent - 2, low word cleared
    qx = from_bits(hi << 32);
  }
  double hz = 0.5 * z - qx;
  double a = 1.0 - qx;
  return a - (hz - (z * r - x * y));
}

// sin and cos of a, |a| <= 100 (the renderer only passes a in [0, 2*pi)).
TOR_HD void sincos(double a, double* s_out, double* c_out) {
  const double INV_PIO2 = 0x1.45f306dc9c883p-1;
  const double PIO2_HI = 0x1.921fb54442d18p+0;   // fl(pi/2)
  const double PIO2_LO = 0x1.1a62633145c07p-54;  // pi/2 - PIO2_HI
  double fn = floor_(a * INV_PIO2 + 0.5);
  int n = (int)fn;
  double r0 = fma_(-fn, PIO2_HI, a);  // exact for the documented range
  double rl = -(fn * PIO2_LO);
  double x = r0 + rl;
  double y = (r0 - x) + rl;
  double s = kernel_sin(x, y);
  double c = kernel_cos(x, y);
  switch (n & 3) {
    case 0: *s_out = s; *c_out = c; break;
    case 1: *s_out = c; *c_out = -s; break;
    case 2: *s_out = -s; *c_out = -c; break;
    default: *s_out = -c; *c_out = s; break;
  }
}

// ---------------------------------------------------------------------------------------
// double-double helpers
// ---------------------------------------------------------------------------------------
struct dd {
  double hi, lo;
};

TOR_HD dd two_sum(double a, double b) {
  double s = a + b;
  double bb = s - a;
  double e = (a - (s - bb)) + (b - bb);
  return dd{s, e};
}
TOR_HD dd quick_two_sum(double a, double b) {
  double s = a + b;
  double e = b - (s - a);
  return dd{s, e};
}
TOR_HD dd two_prod(double a, double b) {
  double p = a * b;
  double e = fma_(a, b, -p);
  return dd{p, e};
}
TOR_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  dd t = two_sum(a.lo, b.lo);
  s.lo += t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo += t.lo;
  return quick_two_sum(s.hi, s.lo);
}
TOR_HD dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo += a.hi * b.lo + a.lo * b.hi;
  return quick_two_sum(p.hi, p.lo);
}
TOR_HD dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo += a.lo * b;
  return quick_two_sum(p.hi, p.lo);
}
TOR_HD dd dd_div(dd a, dd b) {
  double q1 = a.hi / b.hi;
  dd r = dd_add(a, dd_mul_d(b, -q1));
  double q2 = r.hi / b.hi;
  r = dd_add(r, dd_mul_d(b, -q2));
  double q3 = r.hi / b.hi;
  dd q = quick_two_sum(q1, q2);
  return dd_add(q, dd{q3, 0.0});
}

// x^5 through double-double products (error ~2^-100 before the final rounding).
TOR_HD double pow5(double x) {
  dd x2 = two_prod(x, x);
  dd x4 = dd_mul(x2, x2);
  dd x5 = dd_mul_d(x4, x);
  return x5.hi + x5.lo;
}

// 2^n * v for integer n, v normal; saturates to inf / 0.
TOR_HD double scale2(double v, int n) {
  if (n > 1023) {
    v *= 0x1p1023;
    n -= 1023;
    if (n > 1023) n = 1023;
  } else if (n < -1022) {
    v *= 0x1p-969;  // keep v normal while moving towards the subnormal range
    n += 969;
    if (n < -1022) n = -1022;
  }
  return v * from_bits((uint64_t)(n + 1023) << 52);
}

// General pow for the cases the renderer needs: finite y != 0, x >= 0.  Other inputs
// follow C99 where cheap (x == 1, y == 0, NaN propagation, x < 0 -> NaN).
TOR_HD_NOINLINE double pow_general(double x, double y) {
  if (y == 0.0 || x == 1.0) return 1.0;
  if (x != x || y != y) return x + y;
  if (x < 0.0) return from_bits(0x7ff8000000000000ull);
  const double INF = from_bits(0x7ff0000000000000ull);
  if (x == 0.0) return y > 0.0 ? 0.0 : INF;
  if (x == INF) return y > 0.0 ? INF : 0.0;
  if (y == INF) return x > 1.0 ? INF : 0.0;
  if (y == -INF) return x > 1.0 ? 0.0 : INF;

  // x = 2^k * m, m in [sqrt(1/2), sqrt(2))
  int k = 0;
  uint64_t ux = bits_of(x);
  if ((ux >> 52) == 0) {  // subnormal
    x *= 0x1p54;
    ux = bits_of(x);
    k = -54;
  }
  k += (int)(ux >> 52) - 1023;
  double m = from_bits((ux & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
  if (m > 0x1.6a09e667f3bcdp+0) {
    m *= 0.5;
    k += 1;
  }

  // log(m) = 2 atanh(s), s = (m-1)/(m+1), |s| <= 0.1716
  const dd INV_ODD[17] = {
      {0x1.0000000000000p+0, 0x0.0p+0},
      {0x1.5555555555555p-2, 0x1.5555555555555p-56},
      {0x1.999999999999ap-3, -0x1.999999999999ap-57},
      {0x1.2492492492492p-3, 0x1.2492492492492p-57},
      {0x1.c71c71c71c71cp-4, 0x1.c71c71c71c71cp-58},
      {0x1.745d1745d1746p-4, -0x1.745d1745d1746p-59},
      {0x1.3b13b13b13b14p-4, -0x1.3b13b13b13b14p-58},
      {0x1.1111111111111p-4, 0x1.1111111111111p-60},
      {0x1.e1e1e1e1e1e1ep-5, 0x1.e1e1e1e1e1e1ep-61},
      {0x1.af286bca1af28p-5, 0x1.af286bca1af28p-59},
      {0x1.8618618618618p-5, 0x1.8618618618618p-59},
      {0x1.642c8590b2164p-5, 0x1.642c8590b2164p-60},
      {0x1.47ae147ae147bp-5, -0x1.eb851eb851eb8p-61},
      {0x1.2f684bda12f68p-5, 0x1.2f684bda12f68p-59},
      {0x1.1a7b9611a7b96p-5, 0x1.1a7b9611a7b96p-61},
      {0x1.0842108421084p-5, 0x1.0842108421084p-60},
      {0x1.f07c1f07c1f08p-6, -0x1.f07c1f07c1f08p-61},
  };
  dd num = dd{m - 1.0, 0.0};  // exact (Sterbenz)
  dd den = two_sum(m, 1.0);
  dd s = dd_div(num, den);
  dd s2 = dd_mul(s, s);
  dd acc = INV_ODD[16];
  for (int i = 15; i >= 0; --i) acc = dd_add(dd_mul(acc, s2), INV_ODD[i]);
  dd logm = dd_mul(s, acc);
  logm.hi *= 2.0;
  logm.lo *= 2.0;

  const dd LN2 = {0x1.62e42fefa39efp-1, 0x1.abc9e3b39803fp-56};
  dd L = dd_add(dd_mul_d(LN2, (double)k), logm);
  dd P = dd_mul_d(L, y);

  if (P.hi > 709.79) return INF;
  if (P.hi < -745.2) return 0.0;

  const double INV_LN2 = 0x1.71547652b82fep+0;
  double fn = floor_(P.hi * INV_LN2 + 0.5);
  dd r = dd_add(P, dd_mul_d(LN2, -fn));  // |r| <= ~0.347

  const dd INV_FACT[21] = {
      {0x1.0000000000000p+0, 0x0.0p+0},
      {0x1.0000000000000p+0, 0x0.0p+0},
      {0x1.0000000000000p-1, 0x0.0p+0},
      {0x1.5555555555555p-3, 0x1.5555555555555p-57},
      {0x1.5555555555555p-5, 0x1.5555555555555p-59},
      {0x1.1111111111111p-7, 0x1.1111111111111p-63},
      {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65},
      {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73},
      {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
      {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
      {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76},
      {0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80},
      {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
      {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
      {0x1.93974a8c07c9dp-37, 0x1.05d6f8a2efd1fp-92},
      {0x1.ae7f3e733b81fp-41, 0x1.1d8656b0ee8cbp-97},
      {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
      {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
      {0x1.6827863b97d97p-53, 0x1.eec01221a8b0bp-107},
      {0x1.2f49b46814157p-57, 0x1.2650f61dbdcb4p-112},
      {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
  };
  dd e = INV_FACT[20];
  for (int i = 19; i >= 0; --i) e = dd_add(dd_mul(e, r), INV_FACT[i]);
  return scale2(e.hi + e.lo, (int)fn);
}

TOR_HD double pow(double x, double y) {
  if (y == 5.0 && x >= 0x1p-200 && x <= 0x1p+200) return pow5(x);
  return pow_general(x, y);
}

}  // namespace detmath
}  // namespace tor
