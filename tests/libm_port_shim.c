/* Host build of tokengeex_b200/csrc/tgx_libm.h for tests/test_libm_port.py (test infrastructure). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../tokengeex_b200/csrc/tgx_libm.h"
#include "../tokengeex_b200/csrc/tgx_libm_tables.h"

static double H_exp[8], H_log[18], T_log[256];
static int inited = 0;
static void init(void) {
  if (inited) return;
  memcpy(H_exp, TGX_EXP_HDR, sizeof H_exp);
  memcpy(H_log, TGX_LOG_HDR, sizeof H_log);
  memcpy(T_log, TGX_LOG_TAB, sizeof T_log);
  inited = 1;
}
double shim_exp(double x) { init(); return tgx_exp_impl(x, H_exp, TGX_EXP_TAB); }
double shim_log(double x) { init(); return tgx_log_impl(x, H_log, T_log); }
double shim_softplus(double x) { init(); return tgx_softplus_impl(x, H_exp, TGX_EXP_TAB, H_log, T_log); }

/* returns the number of arguments where the port differs from libm (bitwise; NaNs compare equal) */
uint64_t shim_compare(const double* xs, uint64_t n, int which, double* first_bad) {
  init();
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; i++) {
    double a = which == 2 ? shim_softplus(xs[i]) : which ? shim_log(xs[i]) : shim_exp(xs[i]);
    double b = which == 2 ? log(exp(xs[i]) + 1.0) : which ? log(xs[i]) : exp(xs[i]);
    uint64_t ua, ub;
    memcpy(&ua, &a, 8);
    memcpy(&ub, &b, 8);
    if (ua != ub && !(a != a && b != b)) {
      if (!bad && first_bad) *first_bad = xs[i];
      bad++;
    }
  }
  return bad;
}
