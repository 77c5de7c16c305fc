/* Delayed SIR map, a difference equation (model ABI src/difeq.h:19-21), n = 3,
 * parms = (beta, gamma), lag 14 steps through yprev_1
 * (BASELINE.json configs[3]). */
#include <dde/dde.h>

DDE_MODEL_FN void sir_delay(size_t n, size_t step, const double *y,
                            double *y_next, size_t n_out, double *output,
                            const void *data) {
  const double *p = (const double *) data;
  const double beta = p[0], gamma = p[1];
  const double I_lag = yprev_1((int) step - 14, 1);
  const double S = y[0], I = y[1], R = y[2];
  const double inf = beta * S * I_lag;
  const double rec = gamma * I;
  y_next[0] = S - inf;
  y_next[1] = I + inf - rec;
  y_next[2] = R + rec;
}
