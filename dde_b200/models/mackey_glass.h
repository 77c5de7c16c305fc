/* Mackey-Glass delay equation, n = 1, parms = (beta, gamma, n), tau = 17
 * (BASELINE.json configs[1]):
 *   y'(t) = beta y(t - tau) / (1 + y(t - tau)^n) - gamma y(t)
 * Pre-history is the constant y0 (src/dopri.c:777-778). */
#include <dde/dde.h>

DDE_MODEL_FN void mackey_glass(size_t n, double t, const double *y,
                               double *dydt, const void *data) {
  const double *p = (const double *) data;
  const double beta = p[0], gamma = p[1], expo = p[2];
  const double ylag = ylag_1(t - 17.0, 0);
  dydt[0] = beta * ylag / (1.0 + pow(ylag, expo)) - gamma * y[0];
}

/* The one time at which mackey_glass(t, ...) reads the history (optional: lets the
 * solver answer a step attempt's look-ups ahead of its evaluations, dde_model_def.h). */
DDE_MODEL_FN void mackey_glass_lags(double t, const void *data, double *tlag) {
  tlag[0] = t - 17.0;
}
