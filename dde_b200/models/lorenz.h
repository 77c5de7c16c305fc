/* Lorenz attractor, n = 3, parms = (sigma, rho, beta).
 * Same equations and evaluation order as the reference's test fixture
 * (/root/reference/tests/testthat/lorenz.c:4-18), so the reference's
 * known-answer runs apply to it.  Output function: the
 * running (min, max) over the three states, n_out = 2. */
#include <dde/dde.h>

DDE_MODEL_FN void lorenz(size_t n, double t, const double *y, double *dydt,
                         const void *data) {
  const double *p = (const double *) data;
  const double sigma = p[0], rho = p[1], beta = p[2];
  dydt[0] = sigma * (y[1] - y[0]);
  dydt[1] = rho * y[0] - y[1] - y[0] * y[2];
  dydt[2] = -beta * y[2] + y[0] * y[1];
}

DDE_MODEL_FN void lorenz_output(size_t n, double t, const double *y,
                                size_t n_out, double *out, const void *data) {
  out[0] = fmin(fmin(y[0], y[1]), y[2]);
  out[1] = fmax(fmax(y[0], y[1]), y[2]);
}
