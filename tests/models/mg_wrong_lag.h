/* A delay equation whose lag DECLARATION is wrong (it says t - 16, the right-hand side
 * asks for t - 17): the kernel that answers declared look-ups ahead must end every
 * trajectory with DDE_ERR_LAG_UNDECLARED (-11), never integrate something else.  Test
 * fixture only. */
#include <dde/dde.h>

DDE_MODEL_FN void mg_wrong_lag(size_t n, double t, const double *y, double *dydt,
                               const void *data) {
  const double *p = (const double *) data;
  const double ylag = ylag_1(t - 17.0, 0);
  dydt[0] = p[0] * ylag / (1.0 + pow(ylag, p[2])) - p[1] * y[0];
}

DDE_MODEL_FN void mg_wrong_lag_lags(double t, const void *data, double *tlag) {
  tlag[0] = t - 16.0;
}
