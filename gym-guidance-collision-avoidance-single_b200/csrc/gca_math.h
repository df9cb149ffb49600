// gca_math.h - bit-reproducible f64 sincos / log shared by the CUDA kernels and the CPU oracle.
//
// Why this exists: the reference evaluates the ownship / spawn velocity with libm's cos/sin
// (PKG/SingleAircraftEnv.py:274-275,306-307).  libm and CUDA's device library differ in the
// last ulp on a few per cent of arguments, which is enough to break bit-for-bit comparison of
// a GPU run against a CPU run at full batch size.  Every operation below is a single IEEE-754
// round-to-nearest multiply / add / subtract (never contracted into an FMA), so the same input
// gives the same bits on sm_100a and on any x86-64 host.  Accuracy is < 1 ulp for |x| < 1e6
// (checked against mpmath in tests/test_math.py), i.e. the result equals libm's except where
// one of the two is not correctly rounded.
//
// Algorithms: three-step Cody-Waite reduction by pi/2 with 33-bit pieces (exact products for
// |n| < 2^20) followed by the classic degree-13 / degree-14 minimax kernels on [-pi/4, pi/4]
// with a tail correction; log via k*ln2 + log1p(f), f/(2+f) series.
#ifndef GCA_MATH_H_
#define GCA_MATH_H_

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define GCA_HD __host__ __device__ __forceinline__
#else
#define GCA_HD static inline
#endif

// ---- exactly-rounded primitive ops that the compiler may not fuse -------------------------
#if defined(__CUDA_ARCH__)
#define GCA_MUL(a, b) __dmul_rn((a), (b))
#define GCA_ADD(a, b) __dadd_rn((a), (b))
#define GCA_SUB(a, b) __dadd_rn((a), -(b))
#define GCA_DIV(a, b) __ddiv_rn((a), (b))
#define GCA_SQRT(a) __dsqrt_rn((a))
#define GCA_FMA(a, b, c) __fma_rn((a), (b), (c))
#define GCA_RINT(a) rint((a))
#define GCA_FMULF(a, b) __fmul_rn((a), (b))
#define GCA_FADDF(a, b) __fadd_rn((a), (b))
#define GCA_FSUBF(a, b) __fadd_rn((a), -(b))
#define GCA_FDIVF(a, b) __fdiv_rn((a), (b))
#define GCA_FSQRTF(a) __fsqrt_rn((a))
#define GCA_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#define GCA_RINTF(a) rintf((a))
#else
// host: translation units that include this header are built with -ffp-contract=off
#define GCA_MUL(a, b) ((double)(a) * (double)(b))
#define GCA_ADD(a, b) ((double)(a) + (double)(b))
#define GCA_SUB(a, b) ((double)(a) - (double)(b))
#define GCA_DIV(a, b) ((double)(a) / (double)(b))
#define GCA_SQRT(a) __builtin_sqrt((a))
#define GCA_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define GCA_RINT(a) __builtin_rint((a))
#define GCA_FMULF(a, b) ((float)((float)(a) * (float)(b)))
#define GCA_FADDF(a, b) ((float)((float)(a) + (float)(b)))
#define GCA_FSUBF(a, b) ((float)((float)(a) - (float)(b)))
#define GCA_FDIVF(a, b) ((float)((float)(a) / (float)(b)))
#define GCA_FSQRTF(a) __builtin_sqrtf((a))
#define GCA_FMAF(a, b, c) __builtin_fmaf((a), (b), (c))
#define GCA_RINTF(a) __builtin_rintf((a))
#endif

GCA_HD uint64_t gca_f64_bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, sizeof u);
  return u;
#endif
}

GCA_HD double gca_bits_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, sizeof x);
  return x;
#endif
}

// sin(r + t) for |r| <= pi/4 + eps, t the tail of r
GCA_HD double gca_ksin(double r, double t) {
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = GCA_MUL(r, r);
  double w = GCA_MUL(z, z);
  double p = GCA_ADD(GCA_ADD(S2, GCA_MUL(z, GCA_ADD(S3, GCA_MUL(z, S4)))),
                     GCA_MUL(GCA_MUL(z, w), GCA_ADD(S5, GCA_MUL(z, S6))));
  double v = GCA_MUL(z, r);
  // r - ((z*(t/2 - v*p) - t) - v*S1)
  double a = GCA_SUB(GCA_MUL(0.5, t), GCA_MUL(v, p));
  double b = GCA_SUB(GCA_MUL(z, a), t);
  return GCA_SUB(r, GCA_SUB(b, GCA_MUL(v, S1)));
}

// cos(r + t)
GCA_HD double gca_kcos(double r, double t) {
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = GCA_MUL(r, r);
  double w = GCA_MUL(z, z);
  double p = GCA_ADD(GCA_MUL(z, GCA_ADD(C1, GCA_MUL(z, GCA_ADD(C2, GCA_MUL(z, C3))))),
                     GCA_MUL(GCA_MUL(w, w), GCA_ADD(C4, GCA_MUL(z, GCA_ADD(C5, GCA_MUL(z, C6))))));
  double hz = GCA_MUL(0.5, z);
  double q = GCA_SUB(1.0, hz);
  // q + (((1 - q) - hz) + (z*p - r*t))
  double e = GCA_SUB(GCA_SUB(1.0, q), hz);
  double f = GCA_SUB(GCA_MUL(z, p), GCA_MUL(r, t));
  return GCA_ADD(q, GCA_ADD(e, f));
}

// *s = sin(x), *c = cos(x).  Domain of the accuracy claim: |x| < 1e6.
GCA_HD void gca_sincos(double x, double* s, double* c) {
  const double INVPIO2 = 6.36619772367581382433e-01;   // 0x1.45f306dc9c883p-1
  const double P1 = 1.57079632673412561417e+00;        // 0x1.921fb54400000p+0  (33 bits of pi/2)
  const double P2 = 6.07710050630396597660e-11;        // 0x1.0b4611a600000p-34 (next 33 bits)
  const double P3 = 2.02226624871116645580e-21;        // 0x1.3198a2e000000p-69 (next 33 bits)
  const double P3T = 8.47842766036889956997e-32;       // 0x1.b839a252049c1p-104 (the rest)
  double n = GCA_RINT(GCA_MUL(x, INVPIO2));
  double r0 = GCA_SUB(x, GCA_MUL(n, P1));              // exact
  double w1 = GCA_MUL(n, P2);                          // exact
  double sh = GCA_SUB(r0, w1);                         // two_sum(r0, -w1)
  double bb = GCA_SUB(sh, r0);
  double se = GCA_ADD(GCA_SUB(r0, GCA_SUB(sh, bb)), GCA_SUB(-w1, bb));
  double lo = GCA_SUB(GCA_SUB(se, GCA_MUL(n, P3)), GCA_MUL(n, P3T));
  double r = GCA_ADD(sh, lo);                          // fast_two_sum(sh, lo)
  double t = GCA_SUB(lo, GCA_SUB(r, sh));
  double ks = gca_ksin(r, t);
  double kc = gca_kcos(r, t);
  long long q = (long long)n;
  switch ((int)(q & 3)) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
  }
}

// natural logarithm for finite x > 0 (normal numbers; the callers pass x in (2^-53, 1])
GCA_HD double gca_log(double x) {
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double L1 = 6.666666666666735130e-01, L2 = 3.999999999940941908e-01,
               L3 = 2.857142874366239149e-01, L4 = 2.222219843214978396e-01,
               L5 = 1.818357216161805012e-01, L6 = 1.531383769920937332e-01,
               L7 = 1.479819860511658591e-01;
  uint64_t u = gca_f64_bits(x);
  int k = (int)((u >> 52) & 0x7ff) - 1023;
  uint64_t m = u & 0x000fffffffffffffULL;
  // normalise the significand into [sqrt(2)/2, sqrt(2))
  if (m >= 0x6a09e667f3bcdULL) {
    k += 1;
    u = m | 0x3fe0000000000000ULL;
  } else {
    u = m | 0x3ff0000000000000ULL;
  }
  double f = GCA_SUB(gca_bits_f64(u), 1.0);
  double dk = (double)k;
  double sq = GCA_DIV(f, GCA_ADD(2.0, f));
  double z = GCA_MUL(sq, sq);
  double w = GCA_MUL(z, z);
  double t1 = GCA_MUL(w, GCA_ADD(L2, GCA_MUL(w, GCA_ADD(L4, GCA_MUL(w, L6)))));
  double t2 = GCA_MUL(z, GCA_ADD(L1, GCA_MUL(w, GCA_ADD(L3, GCA_MUL(w, GCA_ADD(L5, GCA_MUL(w, L7)))))));
  double R = GCA_ADD(t2, t1);
  double hfsq = GCA_MUL(0.5, GCA_MUL(f, f));
  // dk*ln2_hi - ((hfsq - (s*(hfsq+R) + dk*ln2_lo)) - f)
  double a = GCA_ADD(GCA_MUL(sq, GCA_ADD(hfsq, R)), GCA_MUL(dk, LN2_LO));
  return GCA_SUB(GCA_MUL(dk, LN2_HI), GCA_SUB(GCA_SUB(hfsq, a), f));
}

// Correctly rounded f32 quotient x / d for a divisor known in advance, inv_d = RN(1 / d).
// Multiply by the reciprocal, then two exact-residual FMA corrections (Markstein): the result
// equals IEEE division for every normal quotient (checked exhaustively for the divisors the
// observation uses in tests/test_math.py); 5 FMA-pipe instructions, no MUFU / XU traffic.
GCA_HD float gca_div_const_f32(float x, float d, float inv_d) {
  float q = GCA_FMULF(x, inv_d);
  float r = GCA_FMAF(-q, d, x);
  q = GCA_FMAF(r, inv_d, q);
  r = GCA_FMAF(-q, d, x);
  return GCA_FMAF(r, inv_d, q);
}

// One correction step only.  Exact for a given divisor iff it is exact for all 2^23 significands
// of x (the computation is invariant under scaling x by powers of two away from underflow):
// gca_div1_is_exact checks precisely that, and the host selects this variant only when it holds
// (it does for 800 and for 2 * 8/3 rounded to f32, the divisors of the default Config).
GCA_HD float gca_div_const_f32_1(float x, float d, float inv_d) {
  float q = GCA_FMULF(x, inv_d);
  float r = GCA_FMAF(-q, d, x);
  return GCA_FMAF(r, inv_d, q);
}

static inline int gca_div1_is_exact(float d) {   /* host only */
  const float inv = 1.0f / d;
  for (uint32_t m = 0; m < (1u << 23); ++m) {
    const uint32_t u = (127u << 23) | m;
    float x;
    memcpy(&x, &u, sizeof x);
    if (gca_div_const_f32_1(x, d, inv) != x / d) return 0;
  }
  return 1;
}

// f64 division by a divisor whose correctly rounded reciprocal is known: the closing steps of the classic
// FMA division sequence (Markstein 1990; Cornea, Harrison, Tang 2002): q0 = RN(x * r) is within 2 ulp,
// one exact-residual correction makes it faithful, the second one yields RN(x / d) for every x whose
// quotient is normal - provided the significand of d is not all ones (gca_div_f64_divisor_ok; the
// callers fall back to a real division otherwise).  5 dependent FMA-pipe instructions instead of the
// ~28 of a software IEEE division.  A zero (of either sign) passes through with its sign.
GCA_HD double gca_div_const_f64(double x, double d, double inv_d) {
  double q = GCA_MUL(x, inv_d);
  if (q == 0.0) return q;
  double r = GCA_FMA(-q, d, x);
  q = GCA_FMA(r, inv_d, q);
  r = GCA_FMA(-q, d, x);
  return GCA_FMA(r, inv_d, q);
}

static inline int gca_div_f64_divisor_ok(double d) {   /* host only */
  uint64_t u;
  memcpy(&u, &d, sizeof u);
  const uint64_t mant = u & 0xFFFFFFFFFFFFFull, expo = (u >> 52) & 0x7FF;
  return d > 0 && expo > 64 && expo < 1983 && mant != 0xFFFFFFFFFFFFFull;   /* normal, mid-range, not 2 - ulp */
}

#endif  // GCA_MATH_H_
