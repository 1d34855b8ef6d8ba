// exp / log restated operation-by-operation from glibc 2.39's x86_64 FMA variants
// (__exp_fma / __log_fma: sysdeps/ieee754/dbl-64/e_exp.c, e_log.c built with -mfma; the
// algorithms are ARM optimized-routines' table-driven exp/log, N = 128).
//
// Purpose: the reference computes log_sum_exp (src/lattice.rs:321-333) and the expected
// count update (src/lattice.rs:305-307) with f64::exp / f64::ln, i.e. the platform libm.
// Reproducing libm's exact sequence of roundings — which products are fused into FMAs and
// which are not was read off the disassembly — makes the CUDA forward-backward pass agree
// with the reference bit for bit, instead of to within a few ulps per operation.
// tests/test_libm_port.py compiles this header for the host and compares it with the
// system libm on millions of arguments.
//
// The header is shared by nvcc (device code) and gcc (the test shim), hence the macros.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TGX_HD __host__ __device__ __forceinline__
#else
#define TGX_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define TGX_FMA(a, b, c) __fma_rn((a), (b), (c))
#define TGX_ADD(a, b) __dadd_rn((a), (b))
#define TGX_MUL(a, b) __dmul_rn((a), (b))
#define TGX_AS_F64(u) __longlong_as_double((long long)(u))
#define TGX_AS_U64(d) ((uint64_t)__double_as_longlong(d))
#else
#include <string.h>
TGX_HD double tgx_as_f64_(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
TGX_HD uint64_t tgx_as_u64_(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
#define TGX_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define TGX_ADD(a, b) ((a) + (b))
#define TGX_MUL(a, b) ((a) * (b))
#define TGX_AS_F64(u) tgx_as_f64_(u)
#define TGX_AS_U64(d) tgx_as_u64_(d)
#endif

// H = {invln2N, shift, negln2hiN, negln2loN, C2, C3, C4, C5} (as doubles), T = exp table (bits).
TGX_HD double tgx_exp_impl(double x, const double* H, const uint64_t* T) {
  const uint64_t ix = TGX_AS_U64(x);
  uint32_t abstop = (uint32_t)(ix >> 52) & 0x7ffu;
  if (abstop - 0x3c9u > 0x3eu) {
    if ((int32_t)(abstop - 0x3c9u) < 0) return TGX_ADD(x, 1.0);  // |x| < 2^-54
    if (abstop > 0x408u) {                                        // |x| >= 1024, inf, nan
      if (ix == 0xfff0000000000000ull) return 0.0;
      if (abstop == 0x7ffu) return TGX_ADD(x, 1.0);
      if (ix >> 63) return TGX_MUL(TGX_AS_F64(0x1000000000000000ull), TGX_AS_F64(0x1000000000000000ull));  // __math_uflow
      return TGX_MUL(TGX_AS_F64(0x7000000000000000ull), TGX_AS_F64(0x7000000000000000ull));                // __math_oflow
    }
    abstop = 0;  // 512 <= |x| < 1024: result may be sub/supernormal -> specialcase below
  }
  double kd = TGX_FMA(x, H[0], H[1]);
  const uint64_t ki = TGX_AS_U64(kd);
  kd = TGX_ADD(kd, -H[1]);
  double r = TGX_FMA(kd, H[2], x);
  r = TGX_FMA(kd, H[3], r);
  const uint32_t idx = 2u * (uint32_t)(ki & 127u);
  const uint64_t top = ki << 45;
  const double tail = TGX_AS_F64(T[idx]);
  uint64_t sbits = T[idx + 1] + top;
  const double p23 = TGX_FMA(r, H[5], H[4]);
  const double t3 = TGX_ADD(r, tail);
  const double r2 = TGX_MUL(r, r);
  const double p45 = TGX_FMA(r, H[7], H[6]);
  const double t = TGX_FMA(p23, r2, t3);
  const double r4 = TGX_MUL(r2, r2);
  const double tmp = TGX_FMA(r4, p45, t);
  if (abstop == 0) {  // specialcase()
    if ((ki & 0x80000000ull) == 0) {
      sbits -= 1009ull << 52;
      const double scale = TGX_AS_F64(sbits);
      return TGX_MUL(TGX_FMA(scale, tmp, scale), TGX_AS_F64(0x7f00000000000000ull));  // * 0x1p1009
    }
    sbits += 1022ull << 52;
    const double scale = TGX_AS_F64(sbits);
    const double st = TGX_MUL(tmp, scale);  // not fused in the FMA build either
    double y = TGX_ADD(scale, st);
    if (1.0 > y) {
      const double hi = TGX_ADD(y, 1.0);
      double lo = TGX_ADD(scale, -y);
      lo = TGX_ADD(lo, st);
      double v = TGX_ADD(1.0, -hi);
      v = TGX_ADD(v, y);
      v = TGX_ADD(v, lo);
      v = TGX_ADD(v, hi);
      y = TGX_ADD(v, -1.0);
      if (y == 0.0) y = 0.0;
    }
    return TGX_MUL(y, TGX_AS_F64(0x0010000000000000ull));  // * 0x1p-1022
  }
  const double scale = TGX_AS_F64(sbits);
  return TGX_FMA(scale, tmp, scale);
}

// H = {ln2hi, ln2lo, A[0..4], B[0..10]}, T = {invc, logc} pairs.  x > 0, finite, normal is the
// only class the E-step produces (log(exp(d) + 1.0) with d <= 0); the other classes are
// handled for completeness.
TGX_HD double tgx_log_impl(double x, const double* H, const double* T) {
  uint64_t ix = TGX_AS_U64(x);
  const double* A = H + 2;
  const double* B = H + 7;
  if (ix - 0x3fee000000000000ull <= 0x308ffffffffffull) {  // 1 - 2^-4 <= x < 1 + 0x1.09p-4
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = TGX_ADD(x, -1.0);
    double q1 = TGX_FMA(r, B[2], B[1]);
    double q4 = TGX_FMA(r, B[5], B[4]);
    const double r2 = TGX_MUL(r, r);
    double q7 = TGX_FMA(r, B[8], B[7]);
    q1 = TGX_FMA(r2, B[3], q1);
    q4 = TGX_FMA(r2, B[6], q4);
    const double r3 = TGX_MUL(r, r2);
    q7 = TGX_FMA(r2, B[9], q7);
    q7 = TGX_FMA(r3, B[10], q7);
    q4 = TGX_FMA(q7, r3, q4);
    const double q = TGX_FMA(q4, r3, q1);
    const double c27 = 134217728.0;  // 0x1p27
    const double w = TGX_FMA(r, c27, r);
    const double rhi = TGX_FMA(-c27, r, w);
    const double rhi2 = TGX_MUL(rhi, rhi);
    const double rlo = TGX_ADD(r, -rhi);
    const double hi = TGX_FMA(rhi2, B[0], r);
    const double rmhi = TGX_ADD(r, -hi);
    const double rprhi = TGX_ADD(r, rhi);
    double lo = TGX_FMA(rhi2, B[0], rmhi);
    const double b0rlo = TGX_MUL(B[0], rlo);
    lo = TGX_FMA(b0rlo, rprhi, lo);
    const double y = TGX_FMA(q, r3, lo);
    return TGX_ADD(hi, y);
  }
  uint32_t top = (uint32_t)(ix >> 48);
  if (top - 0x10u > 0x7fdfu) {
    if ((ix << 1) == 0) return TGX_AS_F64(0xfff0000000000000ull);            // log(+-0) = -inf
    if (ix == 0x7ff0000000000000ull) return x;                                // log(inf) = inf
    if ((top & 0x8000u) || (top & 0x7ff0u) == 0x7ff0u) return TGX_AS_F64(0x7ff8000000000000ull);  // x < 0 or nan
    ix = TGX_AS_U64(TGX_MUL(x, 4503599627370496.0));  // subnormal: scale by 2^52
    ix -= 52ull << 52;
  }
  const uint64_t tmp = ix - 0x3fe6000000000000ull;
  const uint32_t i = (uint32_t)(tmp >> 45) & 127u;
  const int32_t k = (int32_t)((int64_t)tmp >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
  const double invc = T[2 * i], logc = T[2 * i + 1];
  const double z = TGX_AS_F64(iz);
  const double kd = (double)k;
  const double w = TGX_FMA(kd, H[0], logc);
  const double r = TGX_FMA(z, invc, -1.0);
  const double p12 = TGX_FMA(r, A[2], A[1]);
  const double hi = TGX_ADD(r, w);
  const double r2 = TGX_MUL(r, r);
  double lo = TGX_ADD(w, -hi);
  lo = TGX_ADD(lo, r);
  lo = TGX_FMA(kd, H[1], lo);
  const double r3 = TGX_MUL(r, r2);
  const double p34 = TGX_FMA(r, A[4], A[3]);
  lo = TGX_FMA(r2, A[0], lo);
  const double p = TGX_FMA(p34, r2, p12);
  const double y = TGX_FMA(r3, p, lo);
  return TGX_ADD(y, hi);
}

// log(exp(x) + 1.0) for the x = vmin - vmax of log_sum_exp (src/lattice.rs:321-333): the composition of the two
// functions above restricted to -60 <= x <= 0, where exp has no over/underflow class and its result + 1.0 lies in
// [1, 2], so the class tests of both collapse to one select (|x| < 2^-54) and one branch (the near-1 window of log).
// Same operations in the same order on that domain => the same bits; anything else (NaN included) takes the general
// functions.  He/Te, Hl/Tl as above.
TGX_HD double tgx_softplus_impl(double x, const double* He, const uint64_t* Te, const double* Hl, const double* Tl) {
  if (!(x >= -60.0 && x <= 0.0)) return tgx_log_impl(TGX_ADD(tgx_exp_impl(x, He, Te), 1.0), Hl, Tl);
  // ---- exp(x), main path
  double kd = TGX_FMA(x, He[0], He[1]);
  const uint64_t ki = TGX_AS_U64(kd);
  kd = TGX_ADD(kd, -He[1]);
  double r = TGX_FMA(kd, He[2], x);
  r = TGX_FMA(kd, He[3], r);
  const uint32_t idx = 2u * (uint32_t)(ki & 127u);
  const uint64_t top = ki << 45;
  const double tail = TGX_AS_F64(Te[idx]);
  const uint64_t sbits = Te[idx + 1] + top;
  const double p23 = TGX_FMA(r, He[5], He[4]);
  const double t3 = TGX_ADD(r, tail);
  const double r2 = TGX_MUL(r, r);
  const double p45 = TGX_FMA(r, He[7], He[6]);
  const double t = TGX_FMA(p23, r2, t3);
  const double r4 = TGX_MUL(r2, r2);
  const double tmp = TGX_FMA(r4, p45, t);
  const double scale = TGX_AS_F64(sbits);
  const uint32_t abstop = (uint32_t)(TGX_AS_U64(x) >> 52) & 0x7ffu;
  const double e = (abstop < 0x3c9u) ? TGX_ADD(x, 1.0) : TGX_FMA(scale, tmp, scale);  // |x| < 2^-54: 1 + x
  // ---- log(s), s = e + 1.0 in [1, 2]
  const double s = TGX_ADD(e, 1.0);
  const uint64_t ix = TGX_AS_U64(s);
  const double* A = Hl + 2;
  const double* B = Hl + 7;
  if (ix - 0x3fee000000000000ull <= 0x308ffffffffffull) {  // s < 1 + 0x1.09p-4
    const double q0 = TGX_ADD(s, -1.0);
    double q1 = TGX_FMA(q0, B[2], B[1]);
    double q4 = TGX_FMA(q0, B[5], B[4]);
    const double q02 = TGX_MUL(q0, q0);
    double q7 = TGX_FMA(q0, B[8], B[7]);
    q1 = TGX_FMA(q02, B[3], q1);
    q4 = TGX_FMA(q02, B[6], q4);
    const double q03 = TGX_MUL(q0, q02);
    q7 = TGX_FMA(q02, B[9], q7);
    q7 = TGX_FMA(q03, B[10], q7);
    q4 = TGX_FMA(q7, q03, q4);
    const double q = TGX_FMA(q4, q03, q1);
    const double c27 = 134217728.0;  // 0x1p27
    const double w = TGX_FMA(q0, c27, q0);
    const double rhi = TGX_FMA(-c27, q0, w);
    const double rhi2 = TGX_MUL(rhi, rhi);
    const double rlo = TGX_ADD(q0, -rhi);
    const double hi = TGX_FMA(rhi2, B[0], q0);
    const double rmhi = TGX_ADD(q0, -hi);
    const double rprhi = TGX_ADD(q0, rhi);
    double lo = TGX_FMA(rhi2, B[0], rmhi);
    const double b0rlo = TGX_MUL(B[0], rlo);
    lo = TGX_FMA(b0rlo, rprhi, lo);
    const double y = TGX_FMA(q, q03, lo);
    return (ix == 0x3ff0000000000000ull) ? 0.0 : TGX_ADD(hi, y);
  }
  const uint64_t tm = ix - 0x3fe6000000000000ull;
  const uint32_t i = (uint32_t)(tm >> 45) & 127u;
  const int32_t k = (int32_t)((int64_t)tm >> 52);  // 0, or 1 for s == 2
  const uint64_t iz = ix - (tm & 0xfff0000000000000ull);
  const double invc = Tl[2 * i], logc = Tl[2 * i + 1];
  const double z = TGX_AS_F64(iz);
  const double kf = (double)k;
  const double w = TGX_FMA(kf, Hl[0], logc);
  const double rr = TGX_FMA(z, invc, -1.0);
  const double p12 = TGX_FMA(rr, A[2], A[1]);
  const double hi = TGX_ADD(rr, w);
  const double rr2 = TGX_MUL(rr, rr);
  double lo = TGX_ADD(w, -hi);
  lo = TGX_ADD(lo, rr);
  lo = TGX_FMA(kf, Hl[1], lo);
  const double rr3 = TGX_MUL(rr, rr2);
  const double p34 = TGX_FMA(rr, A[4], A[3]);
  lo = TGX_FMA(rr2, A[0], lo);
  const double pp = TGX_FMA(p34, rr2, p12);
  const double y = TGX_FMA(rr3, pp, lo);
  return TGX_ADD(y, hi);
}
