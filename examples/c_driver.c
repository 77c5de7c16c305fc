/* c_driver.c -- the C-ABI used the way the reference's own glue would use it
 * (src/r_dopri.c:148-180 calls dopri_data_alloc / dopri_integrate on one
 * trajectory; here one call integrates a batch): plain C, no Python, no torch.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_driver.c -Ldde_b200 -ldde_b200 \
 *       -Wl,-rpath,$PWD/dde_b200 -lm -o build/c_driver
 *   build/c_driver [n_traj]
 *
 * Integrates n_traj Mackey-Glass trajectories (the bench workload's generator,
 * SURVEY.md section 8d), then continues a device-resident batch handle the way
 * Cdopri_continue does, and prints sums the Python binding is checked against
 * (tests/test_c_driver_gpu.py). */
#include <inttypes.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "dde_b200.h"

static double splitmix_uniform(uint64_t k) { /* draw k (0-based) of the stream */
  uint64_t z = 0x9E3779B97F4A7C15ULL + (k + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double) (z >> 11) * 0x1p-53;
}

int main(int argc, char **argv) {
  const size_t n_traj = argc > 1 ? (size_t) strtoull(argv[1], NULL, 10) : 1024;
  const size_t n = 1, n_parms = 3, n_times = 11;
  double times[11], *y0, *parms, *y, *t_last, *step;
  int32_t *stats, *code;
  dde_dopri_control ctl;
  size_t j, i;

  if (dde_device_count() < 1) {
    fprintf(stderr, "no CUDA device: dde_b200 has no CPU fallback\n");
    return 2;
  }
  y0 = malloc(sizeof(double) * n_traj);
  parms = malloc(sizeof(double) * n_parms * n_traj);
  y = malloc(sizeof(double) * n * n_times * n_traj);
  t_last = malloc(sizeof(double) * n_traj);
  step = malloc(sizeof(double) * n_traj);
  stats = malloc(sizeof(int32_t) * 4 * n_traj);
  code = malloc(sizeof(int32_t) * n_traj);
  for (j = 0; j < n_traj; ++j) {
    parms[3 * j + 0] = 0.15 + 0.10 * splitmix_uniform(4 * j + 0);
    parms[3 * j + 1] = 0.08 + 0.04 * splitmix_uniform(4 * j + 1);
    parms[3 * j + 2] = 8.0 + 4.0 * splitmix_uniform(4 * j + 2);
    y0[j] = 0.5 + splitmix_uniform(4 * j + 3);
  }
  for (i = 0; i < n_times; ++i) {
    times[i] = 50.0 * (double) i;
  }
  dde_dopri_control_default(&ctl, DDE_DOPRI_853);
  ctl.rtol = ctl.atol = 1e-6;
  ctl.n_history = 64;

  /* one shot: host arrays in, host arrays out */
  if (dde_dopri_batch("mackey_glass", &ctl, n, 0, n_traj, y0, parms, n_parms, n_parms, times,
                      n_times, NULL, 0, y, NULL, stats, code, t_last, step) != 0) {
    fprintf(stderr, "dde_dopri_batch: %s\n", dde_last_error());
    return 1;
  }
  {
    double sum = 0.0;
    long long evals = 0, accepts = 0, failed = 0;
    for (j = 0; j < n_traj; ++j) {
      evals += stats[4 * j + DDE_STAT_N_EVAL];
      accepts += stats[4 * j + DDE_STAT_N_ACCEPT];
      if (code[j] == DDE_OK_COMPLETE) {
        sum += y[j * n_times + (n_times - 1)];
      } else {
        char msg[256];
        ++failed;
        if (failed == 1) {
          dde_dopri_strerror(code[j], t_last[j], ctl.n_history, msg, sizeof msg);
          printf("first failure: trajectory %zu: %s\n", j, msg);
        }
      }
    }
    printf("batch n_traj %zu evals %lld accepts %lld failed %lld sum_y_end %.17g\n", n_traj, evals,
           accepts, failed, sum);
  }

  /* device-resident handle: run to t = 250, continue to t = 500 (Cdopri_continue) */
  {
    dde_dopri_batch_t *b = dde_dopri_batch_alloc("mackey_glass", &ctl, n, 0, n_parms, n_traj);
    const double first[2] = {0.0, 250.0}, second[2] = {250.0, 500.0};
    double sum = 0.0;
    long long restarted = 0;
    if (b == NULL || dde_dopri_batch_upload(b, y0, parms, n_parms) != 0 ||
        dde_dopri_batch_set_times(b, first, 2, NULL, 0) != 0 || dde_dopri_batch_run(b) != 0 ||
        dde_dopri_batch_continue(b, NULL, second, 2, NULL, 0) != 0 ||
        dde_dopri_batch_download(b, y, NULL, stats, code, t_last, step) != 0) {
      fprintf(stderr, "handle API: %s\n", dde_last_error());
      return 1;
    }
    for (j = 0; j < n_traj; ++j) {
      if (code[j] == DDE_OK_COMPLETE) {
        sum += y[j * 2 + 1];
        ++restarted;
      }
    }
    printf("continue completed %lld sum_y_end %.17g last_ms %.3f\n", restarted, sum,
           dde_dopri_batch_last_ms(b));
    dde_dopri_batch_free(b);
  }
  free(y0); free(parms); free(y); free(t_last); free(step); free(stats); free(code);
  return 0;
}
