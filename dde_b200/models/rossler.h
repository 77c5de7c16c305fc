/* Roessler system, n = 3, parms = (a, b, c)  (SURVEY.md section 8d, C3):
 *   x' = -y - z,  y' = x + a y,  z' = b + z (x - c) */
#include <dde/dde.h>

DDE_MODEL_FN void rossler(size_t n, double t, const double *y, double *dydt,
                          const void *data) {
  const double *p = (const double *) data;
  const double a = p[0], b = p[1], c = p[2];
  dydt[0] = -y[1] - y[2];
  dydt[1] = y[0] + a * y[1];
  dydt[2] = b + y[2] * (y[0] - c);
}
