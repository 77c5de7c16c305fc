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

/* as delayed_growth, each equation driven by the lagged value of its mirror
   image, fetched with ylag_vec (size_t indices, src/dopri.c:802-814); n <= 16 */
DDE_MODEL_FN void delayed_growth_vec(size_t n, double t, const double *y,
                                     double *dydt, const void *data) {
  const double *r = (const double *) data;
  size_t idx[16];
  double lagged[16];
  for (size_t i = 0; i < n; ++i) {
    idx[i] = n - 1 - i;
  }
  ylag_vec(t - 1.0, idx, n, lagged);
  for (size_t i = 0; i < n; ++i) {
    dydt[i] = r[i] * lagged[i];
  }
}

/* ... the whole lagged state fetched with ylag_all (src/dopri.c:790-800) by a
   thread-per-trajectory model; n <= 16 */
DDE_MODEL_FN void delayed_growth_all(size_t n, double t, const double *y,
                                     double *dydt, const void *data) {
  const double *r = (const double *) data;
  double lagged[16];
  ylag_all(t - 1.0, lagged);
  for (size_t i = 0; i < n; ++i) {
    dydt[i] = r[i] * lagged[n - 1 - i];
  }
}

/* The reference's event fixtures (/root/reference/tests/testthat/growth.c:20-49):
 * applied to the state -- or to the parameters -- at critical times. */
DDE_MODEL_FN void identity(size_t n, double t, double *y, void *data) {
}

DDE_MODEL_FN void double_variables(size_t n, double t, double *y, void *data) {
  for (size_t i = 0; i < n; ++i) {
    y[i] *= 2.0;
  }
}

DDE_MODEL_FN void halve_variables(size_t n, double t, double *y, void *data) {
  for (size_t i = 0; i < n; ++i) {
    y[i] /= 2.0;
  }
}

DDE_MODEL_FN void double_parameters(size_t n, double t, double *y, void *data) {
  double *r = (double *) data;
  for (size_t i = 0; i < n; ++i) {
    r[i] *= 2;
  }
}

DDE_MODEL_FN void halve_parameters(size_t n, double t, double *y, void *data) {
  double *r = (double *) data;
  for (size_t i = 0; i < n; ++i) {
    r[i] /= 2;
  }
}

/* The same two models written for wide state vectors with the cooperative
 * macros of <dde/dde.h> (any n; the lagged state is fetched whole with
 * ylag_all). */
DDE_MODEL_FN void exponential_wide(size_t n, double t, const double *y,
                                   double *dydt, const void *data) {
  const double *r = (const double *) data;
  DDE_FOR(i, 0, n) {
    dydt[i] = -y[i] * r[i];
  }
}

#define DDE_GROWTH_WIDE_MAX 128
DDE_MODEL_FN void delayed_growth_wide(size_t n, double t, const double *y,
                                      double *dydt, const void *data) {
  const double *r = (const double *) data;
  DDE_SCRATCH_BEGIN();
  DDE_SCRATCH(double, ylag, DDE_GROWTH_WIDE_MAX);
  ylag_all(t - 1.0, ylag);
  DDE_FOR(i, 0, n) {
    dydt[i] = r[i] * ylag[i];
  }
}
