/* SEIR with a latent-period delay, n = 4 (S, E, I, R), hard-coded
 * parameters.  Same equations and evaluation order as the reference's
 * example (/root/reference/inst/examples/seir.c:5-30), in two flavours that
 * the reference tests require to agree (tests/testthat/test-dde.R:272-283):
 *   seir      -- two ylag_1 look-ups
 *   seir_int  -- one ylag_vec_int look-up (tests/testthat/seir_int.c:5-33)
 * Output function: the total population, n_out = 1. */
#include <dde/dde.h>

DDE_MODEL_FN void seir_rhs(const double *y, double S_lag, double I_lag,
                           double *dydt) {
  double b = 0.1, N = 1e7, beta = 10.0, sigma = 1.0 / 3.0, delta = 1.0 / 21.0,
         t_latent = 14.0;
  double births = N * b, surv = exp(-b * t_latent);
  const double S = y[0], E = y[1], I = y[2], R = y[3];
  const double new_inf = beta * S * I / N;
  const double lag_inf = beta * S_lag * I_lag * surv / N;
  dydt[0] = births - b * S - new_inf + delta * R;
  dydt[1] = new_inf - lag_inf - b * E;
  dydt[2] = lag_inf - (b + sigma) * I;
  dydt[3] = sigma * I - b * R - delta * R;
}

DDE_MODEL_FN void seir(size_t n, double t, const double *y, double *dydt,
                       const void *data) {
  const double tau = t - 14.0;
  const double S_lag = ylag_1(tau, 0), I_lag = ylag_1(tau, 2);
  seir_rhs(y, S_lag, I_lag, dydt);
}

DDE_MODEL_FN void seir_int(size_t n, double t, const double *y, double *dydt,
                           const void *data) {
  const double tau = t - 14.0;
  const int idx[2] = {0, 2};
  double lagged[2];
  ylag_vec_int(tau, idx, 2, lagged);
  seir_rhs(y, lagged[0], lagged[1], dydt);
}

DDE_MODEL_FN void seir_output(size_t n, double t, const double *y,
                              size_t n_out, double *out, const void *data) {
  out[0] = y[0] + y[1] + y[2] + y[3];
}
