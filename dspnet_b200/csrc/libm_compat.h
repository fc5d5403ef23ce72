// Bit-exact re-evaluation of glibc 2.39's expf / logf (sysdeps/ieee754/flt-32/e_expf.c, e_logf.c -- the
// ARM "optimized routines" algorithms: table lookup + low-degree polynomial, all in binary64, rounded once to
// binary32).  The reference CPU operators call std::exp / std::log on float
// (operator/multibox_target.cc:53-54,228,230; operator/multibox_detection.cc:117-118), i.e. the host libm, and
// the discrete outputs of the path (hard-negative ranking, NMS decisions) depend on those bits, so the CUDA
// kernels reproduce the same sequence of correctly rounded binary64 operations instead of calling CUDA's
// expf/logf (<= 2 ulp, different bits).
//
// glibc ships two builds of each routine and picks one at load time (ifunc): a plain SSE2 build and an
// FMA build (sysdeps/x86_64/fpu/multiarch, -mfma -mavx2) in which gcc contracted specific multiply-adds.
// The contraction pattern below was read from the disassembly of the libm.so.6 of this image
// (md5 f8e590c62ca7258e57ee58b356a68ef6, Ubuntu GLIBC 2.39-0ubuntu8.5) and is verified exhaustively
// against it by tests/test_libm_compat.py (host build of this same header).
//
// The header compiles both as CUDA device code and as plain host C++ (for that exhaustive check).
#ifndef DSPMB_LIBM_COMPAT_H_
#define DSPMB_LIBM_COMPAT_H_

#include <stdint.h>

#ifdef __CUDACC__
#define DSPMB_HD __host__ __device__ __forceinline__
#define DSPMB_HDM __host__ __device__ __forceinline__
#else
#define DSPMB_HDM inline
#include <math.h>
#include <string.h>
#define DSPMB_HD static inline
#endif

namespace dspmb {
namespace libm {

// ---- correctly rounded binary64 primitives that the compiler must not contract or reassociate ----
#ifdef __CUDA_ARCH__
DSPMB_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
DSPMB_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
DSPMB_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
DSPMB_HD double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
DSPMB_HD uint64_t dbits(double a) { return (uint64_t)__double_as_longlong(a); }
DSPMB_HD double dfrom(uint64_t a) { return __longlong_as_double((long long)a); }
DSPMB_HD uint32_t fbits(float a) { return __float_as_uint(a); }
DSPMB_HD float ffrom(uint32_t a) { return __uint_as_float(a); }
DSPMB_HD float d2f(double a) { return __double2float_rn(a); }
#else  // host build: compile with -ffp-contract=off
DSPMB_HD double dmul(double a, double b) { return a * b; }
DSPMB_HD double dadd(double a, double b) { return a + b; }
DSPMB_HD double dsub(double a, double b) { return a - b; }
DSPMB_HD double dfma(double a, double b, double c) { return fma(a, b, c); }
DSPMB_HD uint64_t dbits(double a) { uint64_t u; memcpy(&u, &a, 8); return u; }
DSPMB_HD double dfrom(uint64_t a) { double d; memcpy(&d, &a, 8); return d; }
DSPMB_HD uint32_t fbits(float a) { uint32_t u; memcpy(&u, &a, 4); return u; }
DSPMB_HD float ffrom(uint32_t a) { float f; memcpy(&f, &a, 4); return f; }
DSPMB_HD float d2f(double a) { return (float)a; }
#endif

// 2^(i/32) table: bits(2^(i/32)) - (i << 47)   (__exp2f_data.tab, EXP2F_TABLE_BITS = 5)
#define DSPMB_EXP2F_TAB \
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, \
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, \
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull, \
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, \
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull, \
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull, \
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, \
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull,
// {invc, logc} pairs of __logf_data.tab (LOGF_TABLE_BITS = 4), as binary64 bit patterns.
#define DSPMB_LOGF_TAB \
    0x3ff661ec79f8f3beull, 0xbfd57bf7808caadeull, 0x3ff571ed4aaf883dull, 0xbfd2bef0a7c06ddbull, \
    0x3ff49539f0f010b0ull, 0xbfd01eae7f513a67ull, 0x3ff3c995b0b80385ull, 0xbfcb31d8a68224e9ull, \
    0x3ff30d190c8864a5ull, 0xbfc6574f0ac07758ull, 0x3ff25e227b0b8ea0ull, 0xbfc1aa2bc79c8100ull, \
    0x3ff1bb4a4a1a343full, 0xbfba4e76ce8c0e5eull, 0x3ff12358f08ae5baull, 0xbfb1973c5a611cccull, \
    0x3ff0953f419900a7ull, 0xbfa252f438e10c1eull, 0x3ff0000000000000ull, 0x0000000000000000ull, \
    0x3fee608cfd9a47acull, 0x3faaa5aa5df25984ull, 0x3feca4b31f026aa0ull, 0x3fbc5e53aa362eb4ull, \
    0x3feb2036576afce6ull, 0x3fc526e57720db08ull, 0x3fe9c2d163a1aa2dull, 0x3fcbc2860d224770ull, \
    0x3fe886e6037841edull, 0x3fd1058bc8a07ee1ull, 0x3fe767dcf5534862ull, 0x3fd4043057b6ee09ull,

static const uint64_t kExp2fTabHost[32] = {DSPMB_EXP2F_TAB};
static const uint64_t kLogfTabHost[32] = {DSPMB_LOGF_TAB};
#ifdef __CUDACC__
static __device__ const uint64_t kExp2fTabDev[32] = {DSPMB_EXP2F_TAB};
static __device__ const uint64_t kLogfTabDev[32] = {DSPMB_LOGF_TAB};
#endif
DSPMB_HD uint64_t exp2f_tab(unsigned i) {
#ifdef __CUDA_ARCH__
  return kExp2fTabDev[i];
#else
  return kExp2fTabHost[i];
#endif
}
DSPMB_HD uint64_t logf_tab(unsigned i) {
#ifdef __CUDA_ARCH__
  return kLogfTabDev[i];
#else
  return kLogfTabHost[i];
#endif
}

// expf: x*N/ln2 = k + r, exp(x) = 2^(k/N) * (C0 r^3 + C1 r^2 + C2 r + 1), N = 32.
// kFma selects glibc's FMA build; `tab(i)` returns entry i of the 2^(i/32) table (global copy, or a shared-memory
// copy in the hot kernels).
template <bool kFma, typename Tab>
DSPMB_HD float expf_glibc_t(float x, Tab tab) {
  const double kShift = 0x1.8p+52;
  const double kInvLn2N = 0x1.71547652b82fep+5;
  const double kC0 = 0x1.c6af84b912394p-20, kC1 = 0x1.ebfce50fac4f3p-13, kC2 = 0x1.62e42ff0c52d6p-6;
  const uint32_t ix = fbits(x);
  const uint32_t abstop = (ix >> 20) & 0x7ff;
  if (abstop >= 0x42b) {            // |x| >= 88 or NaN
    if (ix == 0xff800000u) return 0.0f;
    if (abstop >= 0x7f8) return x + x;
    if (x > 0x1.62e42ep6f) return ffrom(0x7f800000u);   // overflow
    if (x < -0x1.9fe368p6f) return 0.0f;                 // underflow to zero
    if (x < -0x1.9d1d9ep6f) return ffrom(0x00000001u);   // __math_may_uflowf: 0x1.4p-75f squared
  }
  const double xd = (double)x;
  double kd, r;
  if (kFma) {
    kd = dfma(kInvLn2N, xd, kShift);
  } else {
    kd = dadd(dmul(kInvLn2N, xd), kShift);
  }
  const uint64_t ki = dbits(kd);
  kd = dsub(kd, kShift);
  if (kFma) {
    r = dfma(kInvLn2N, xd, -kd);
  } else {
    r = dsub(dmul(kInvLn2N, xd), kd);
  }
  const double s = dfrom(tab((unsigned)(ki & 31)) + (ki << 47));
  double z, r2, y;
  if (kFma) {
    z = dfma(kC0, r, kC1);
    r2 = dmul(r, r);
    y = dfma(kC2, r, 1.0);
    y = dfma(z, r2, y);
  } else {
    z = dadd(dmul(kC0, r), kC1);
    r2 = dmul(r, r);
    y = dadd(dmul(kC2, r), 1.0);
    y = dadd(dmul(z, r2), y);
  }
  return d2f(dmul(y, s));
}
struct GlobalExpTab {
  DSPMB_HDM uint64_t operator()(unsigned i) const { return exp2f_tab(i); }
};
DSPMB_HD float expf_glibc(float x, bool fma_build) {
  return fma_build ? expf_glibc_t<true>(x, GlobalExpTab()) : expf_glibc_t<false>(x, GlobalExpTab());
}

// logf: x = 2^k z, z in [OFF, 2 OFF); log(x) = log1p(z/c - 1) + log(c) + k ln2 with c from a 16-entry table.
DSPMB_HD float logf_glibc(float x, bool fma_build) {
  const double kLn2 = 0x1.62e42fefa39efp-1;
  const double kA0 = -0x1.00ea348b88334p-2, kA1 = 0x1.5575b0be00b6ap-2, kA2 = -0x1.ffffef20a4123p-2;
  uint32_t ix = fbits(x);
  if (ix == 0x3f800000u) return 0.0f;
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    if (ix * 2 == 0) return ffrom(0xff800000u);                                      // log(+-0) = -inf
    if (ix == 0x7f800000u) return x;                                                  // log(inf) = inf
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return ffrom(0x7fc00000u) ;      // x < 0 or NaN
    ix = fbits(x * 0x1p23f);                                                          // subnormal: normalise
    ix -= 23u << 23;
  }
  const uint32_t tmp = ix - 0x3f330000u;
  const int i = (tmp >> 19) & 15;
  const int k = (int32_t)tmp >> 23;
  const uint32_t iz = ix - (tmp & 0xff800000u);
  const double invc = dfrom(logf_tab(2 * i));
  const double logc = dfrom(logf_tab(2 * i + 1));
  const double z = (double)ffrom(iz);
  double r, y0, r2, y;
  if (fma_build) {
    y0 = dfma((double)k, kLn2, logc);
    r = dfma(z, invc, -1.0);
    y = dfma(kA1, r, kA2);
    r2 = dmul(r, r);
    const double y0r = dadd(r, y0);
    y = dfma(r2, kA0, y);
    return d2f(dfma(r2, y, y0r));
  }
  r = dsub(dmul(z, invc), 1.0);
  y = dadd(dmul(kA1, r), kA2);
  r2 = dmul(r, r);
  y = dadd(dmul(kA0, r2), y);
  y0 = dadd(dmul((double)k, kLn2), logc);
  return d2f(dadd(dadd(y0, r), dmul(y, r2)));
}

}  // namespace libm
}  // namespace dspmb
#endif  // DSPMB_LIBM_COMPAT_H_
