/* Difference-equation test fixtures (model ABI src/difeq.h:19-21): compiled into
 * tests/models/libdde_test_models.so, not into the product library.
 *
 *   logistic   -- y' = r y (1 - y), any n, parms = r[n]
 *                 (/root/reference/tests/testthat/logistic.c:4-11)
 *   additive   -- y' = y + p, any n, parms = p[n]; exact in FP64 for integer
 *                 data (tests/testthat/test-difeq.R:13-20)
 *   fibonacci  -- y' = y + y(step - 1) through yprev_vec_int, n = 5
 *                 (tests/testthat/growth_int.c:3-13; with y0 = 1 this is
 *                 1, 2, 3, 5, 8, ... , tests/testthat/test-difeq-delay.R:3-24)
 *   fib_out    -- as fibonacci, n = 1 via yprev_1, with n_out = 1 output = the
 *                 lagged value (exercises the output-row pairing,
 *                 src/difeq.c:204-213,249-255)
 *   fib_all    -- y' = y + y(step - 1) through yprev_all, n <= 16
 *   fib_vec    -- y'_i = y_i + y_{n-1-i}(step - 1) through yprev_vec (size_t
 *                 indices), n <= 16
 */
#include <dde/dde.h>

DDE_MODEL_FN void logistic(size_t n, size_t step, const double *y,
                           double *y_next, size_t n_out, double *output,
                           const void *data) {
  const double *r = (const double *) data;
  for (size_t i = 0; i < n; ++i) {
    y_next[i] = r[i] * y[i] * (1 - y[i]);
  }
}

DDE_MODEL_FN void additive(size_t n, size_t step, const double *y,
                           double *y_next, size_t n_out, double *output,
                           const void *data) {
  const double *p = (const double *) data;
  for (size_t i = 0; i < n; ++i) {
    y_next[i] = y[i] + p[i];
  }
  for (size_t i = 0; i < n_out; ++i) {
    output[i] = y[i % n] * 2.0 + (double) step;
  }
}

DDE_MODEL_FN void fibonacci(size_t n, size_t step, const double *y,
                            double *y_next, size_t n_out, double *output,
                            const void *data) {
  const int idx[5] = {0, 1, 2, 3, 4};
  double prev[5];
  yprev_vec_int((int) step - 1, idx, 5, prev);
  for (size_t i = 0; i < n; ++i) {
    y_next[i] = y[i] + prev[i];
  }
}

DDE_MODEL_FN void fib_out(size_t n, size_t step, const double *y,
                          double *y_next, size_t n_out, double *output,
                          const void *data) {
  const double prev = yprev_1((int) step - 1, 0);
  y_next[0] = y[0] + prev;
  output[0] = prev;
}

DDE_MODEL_FN void fib_all(size_t n, size_t step, const double *y,
                          double *y_next, size_t n_out, double *output,
                          const void *data) {
  double prev[16];
  yprev_all((int) step - 1, prev);
  for (size_t i = 0; i < n; ++i) {
    y_next[i] = y[i] + prev[i];
  }
}

DDE_MODEL_FN void fib_vec(size_t n, size_t step, const double *y,
                          double *y_next, size_t n_out, double *output,
                          const void *data) {
  size_t idx[16];
  double prev[16];
  for (size_t i = 0; i < n; ++i) {
    idx[i] = n - 1 - i;
  }
  yprev_vec((int) step - 1, idx, n, prev);
  for (size_t i = 0; i < n; ++i) {
    y_next[i] = y[i] + prev[i];
  }
}
