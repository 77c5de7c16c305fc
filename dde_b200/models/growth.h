/* Per-component exponential decay and constant growth, any n, parms = r[n]
 * (the reference's event-test fixtures,
 * /root/reference/tests/testthat/growth.c:4-19):
 *   exponential: y_i' = -y_i r_i          linear: y_i' = r_i
 * and a delayed variant used for the analytic DDE checks
 * (tests/testthat/test-dde.R:162-211): y_i' = r_i y_i(t - 1). */
#include <dde/dde.h>

DDE_MODEL_FN void exponential(size_t n, double t, const double *y,
                              double *dydt, const void *data) {
  const double *r = (const double *) data;
  for (size_t i = 0; i < n; ++i) {
    dydt[i] = -y[i] * r[i];
  }
}

DDE_MODEL_FN void linear(size_t n, double t, const double *y, double *dydt,
                         const void *data) {
  const double *r = (const double *) data;
  for (size_t i = 0; i < n; ++i) {
    dydt[i] = r[i];
  }
}

DDE_MODEL_FN void delayed_growth(size_t n, double t, const double *y,
                                 double *dydt, const void *data) {
  const double *r = (const double *) data;
  for (size_t i = 0; i < n; ++i) {
    dydt[i] = r[i] * ylag_1(t - 1.0, i);
  }
}
