/* emu_driver.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h).
 *
 * The thread-per-trajectory dopri kernels of include/dde_b200 compiled by g++ and
 * run one lane at a time over host arrays, behind the batch signature of the
 * checkers (oracle/pyoracle.py): tests/test_emu_cpu.py compares the result
 * with the oracle bit for bit.  This checks the device code's control flow and
 * arithmetic where there is no GPU; the GPU parity tests stay the gate.
 *
 * Build: tests/emu/build.sh (g++ -O2 -mfma -ffp-contract=off, the flags the
 * oracle's restatement is built with).
 */
#include "cuda_emu.h"

#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "dde_b200.h"
#include <dde/dde.h>

#include "dopri_kernels.cuh"
#include "dopri_dde1_kernel.cuh"
#include "dde_model_def.h"

#define pow(x, y) dde_pow((x), (y))
#define exp(x) dde_exp((x))

#include "../../dde_b200/models/lorenz.h"
#include "../../dde_b200/models/mackey_glass.h"
#include "../../dde_b200/models/rossler.h"
#include "../models/growth.h"
#include "../models/seir.h"

#define DDE_SCOPED_MODEL mackey_glass
#define DDE_SCOPED_SOURCE "mackey_glass.h"
#include "dde_model_scoped.inc"

DDE_DEFINE_DOPRI_MODEL(lorenz, lorenz, 3, 3, lorenz_output, 2, DDE_NO_EVENTS, 0)
DDE_DEFINE_DOPRI_MODEL(rossler, rossler, 3, 3, DDE_NO_OUTPUT, 0, DDE_NO_EVENTS, 0)
DDE_DEFINE_DOPRI_MODEL_LAGS(mackey_glass, mackey_glass, 1, 3, DDE_NO_OUTPUT, 0, DDE_NO_EVENTS, 0,
                            mackey_glass_lags, 1)
DDE_DEFINE_DOPRI_MODEL(seir, seir, 4, -1, seir_output, 1, DDE_NO_EVENTS, 0)
DDE_DEFINE_DOPRI_MODEL(delayed_growth_vec, delayed_growth_vec, 0, -1, DDE_NO_OUTPUT, 0, DDE_NO_EVENTS, 0)
DDE_DEFINE_DOPRI_MODEL(delayed_growth_all, delayed_growth_all, 0, -1, DDE_NO_OUTPUT, 0, DDE_NO_EVENTS, 0)

namespace {

std::string g_message;

/* run `kernel` as a grid of blocks of DDE_TPB threads, one thread after the other */
template <class K>
void run_grid(K kernel, const dde::DopriArgs &a, unsigned grid, unsigned block_threads) {
  blockDim.x = block_threads;
  gridDim.x = grid;
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    for (unsigned t = 0; t < block_threads; ++t) {
      threadIdx.x = t;
      kernel(a);
    }
  }
  threadIdx.x = 0;
  blockIdx.x = 0;
}

template <class M, int METHOD, int MODE>
void run_model(dde::DopriArgs a) {
  constexpr size_t order = METHOD == 0 ? 5 : 8;
  a.win_staged = 0;
  if (a.n_history > 0 &&
      sizeof(double) * DDE_WIN * order * (size_t) a.n * DDE_TPB <= DDE_WIN_COEF_BUDGET) {
    a.win_staged = 1;
  }
  const unsigned init_grid = (unsigned) ((a.n_traj + DDE_TPB - 1) / DDE_TPB);
  run_grid(dde::dopri_init_kernel<M, METHOD>, a, init_grid, DDE_TPB);
  /* the stepping kernel: persistent threads -- one lane works through the batch */
  run_grid(dde::dopri_tpt_kernel<M, METHOD, MODE>, a, 1, 1);
}

} /* namespace */

/* the library is built with hidden visibility and -Bsymbolic so that its instantiations of
   the kernel templates cannot be interposed by the product library's launch stubs of the
   same names when both are loaded into one process */
#pragma GCC visibility push(default)
extern "C" {

const char *emu_last_message(void) { return g_message.c_str(); }
double emu_last_seconds(void) { return 0.0; }

#ifdef DDE_LAG_STATS
void emu_lag_stats(unsigned long long *out) {
  memcpy(out, dde::dde_lag_stats, sizeof dde::dde_lag_stats);
  memset(dde::dde_lag_stats, 0, sizeof dde::dde_lag_stats);
}
#endif

/* mode: the stepping kernel's MODE template argument (dopri_kernels.cuh) */
static int g_mode = -1;
void emu_set_mode(int mode) { g_mode = mode; }

int emu_dopri_batch(const char *model, const dde_dopri_control *ctl, size_t n, size_t n_out,
                    size_t n_traj, const double *y0, const double *parms, size_t n_parms,
                    size_t parms_stride, const double *times, size_t n_times,
                    const double *tcrit, size_t n_tcrit, double *y_out, double *out,
                    int32_t *stats, int32_t *code, double *t_last, double *step_size) {
  const std::string name(model);
  const int order = ctl->method == 0 ? 5 : 8;
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  const size_t hl = (size_t) order * n + 2;
  memset(y_out, 0, sizeof(double) * n * nt * n_traj);
  if (out != nullptr) {
    memset(out, 0, sizeof(double) * n_out * nt * n_traj);
  }
  std::vector<double> ring((size_t) ctl->n_history * hl * n_traj + 1, 0.0);
  std::vector<unsigned> ring_count(n_traj, 0u);
  std::vector<double> init_k1(n * n_traj, 0.0), init_h(n_traj, 0.0), y_final(n * n_traj, 0.0);
  unsigned long long next = 0;

  dde::DopriArgs a;
  memset(&a, 0, sizeof a);
  a.n_traj = (long long) n_traj;
  a.n = (int) n;
  a.n_out = (int) n_out;
  a.n_parms = (int) n_parms;
  a.parms_stride = (long long) parms_stride;
  a.y0 = y0;
  a.parms = parms;
  a.times = times;
  a.tcrit = tcrit;
  a.n_times = (int) n_times;
  a.n_tcrit = (int) n_tcrit;
  a.sign = times[n_times - 1] >= times[0] ? 1.0 : -1.0;
  /* dopri_data_reset, src/dopri.c:193-211: critical times at or before the start are skipped */
  a.tcrit_idx0 = 0;
  while (a.tcrit_idx0 < (int) n_tcrit && a.sign * tcrit[a.tcrit_idx0] <= a.sign * times[0]) {
    ++a.tcrit_idx0;
  }
  a.atol = ctl->atol;
  a.rtol = ctl->rtol;
  a.step_size_min = ctl->step_size_min;
  a.step_size_max = ctl->step_size_max;
  a.step_size_initial = ctl->step_size_initial;
  a.step_max_n = ctl->step_max_n;
  a.stiff_check = ctl->stiff_check;
  a.n_history = (int) ctl->n_history;
  a.step_size_min_allow = ctl->step_size_min_allow;
  a.return_initial = ctl->return_initial;
  a.y_out = y_out;
  a.out = out;
  a.stats = stats;
  a.code = code;
  a.t_last = t_last;
  a.step_size = step_size;
  a.ring = ring.data();
  a.ring_count = ring_count.data();
  a.init_k1 = init_k1.data();
  a.init_h = init_h.data();
  a.y_final = y_final.data();
  a.next_traj = &next;
  for (size_t j = 0; j < n_traj; ++j) {
    code[j] = 0;
  }

  const int method = ctl->method;
  const bool hist = ctl->n_history > 0;
#define EMU_RUN(M, MODE)                   \
  do {                                     \
    if (method == 0) {                     \
      run_model<M, 0, MODE>(a);            \
    } else {                               \
      run_model<M, 1, MODE>(a);            \
    }                                      \
  } while (0)
  if (name == "mackey_glass") {
    /* what the product launches (dde_model.cuh): dopri_dde1_kernel when it can take the run */
    const bool dde1 = method == 1 && hist && a.sign > 0 && n_tcrit == 0 && n_out == 0 &&
                      ctl->stiff_check == 0;
    const int mode = g_mode >= 0 ? g_mode : (dde1 ? 4 : (hist ? 2 : 1));
    if (mode == 0) {
      EMU_RUN(dde_dopri_model_mackey_glass, 0);
    } else if (mode == 1) {
      EMU_RUN(dde_dopri_model_mackey_glass, 1);
    } else if (mode == 2) {
      EMU_RUN(dde_dopri_model_mackey_glass, 2);
    } else { /* 4: dopri_dde1_kernel */
      if (!dde1) {
        g_message = "emu: dopri_dde1_kernel does not take this run";
        return 1;
      }
      a.win_staged = 1;
      const unsigned init_grid = (unsigned) ((a.n_traj + DDE_TPB - 1) / DDE_TPB);
      run_grid(dde::dopri_init_kernel<dde_dopri_model_mackey_glass, 1>, a, init_grid, DDE_TPB);
      run_grid(dde::dopri_dde1_kernel<dde_dopri_model_mackey_glass>, a, 1, 1);
    }
  } else if (name == "lorenz") {
    const int mode = g_mode >= 0 ? g_mode : (hist ? 0 : 1);
    if (mode == 1) {
      EMU_RUN(dde_dopri_model_lorenz, 1);
    } else {
      EMU_RUN(dde_dopri_model_lorenz, 0);
    }
  } else if (name == "rossler") {
    const int mode = g_mode >= 0 ? g_mode : (hist ? 0 : 1);
    if (mode == 1) {
      EMU_RUN(dde_dopri_model_rossler, 1);
    } else {
      EMU_RUN(dde_dopri_model_rossler, 0);
    }
  } else if (name == "seir") {
    EMU_RUN(dde_dopri_model_seir, 0);
  } else if (name == "delayed_growth_vec") {
    EMU_RUN(dde_dopri_model_delayed_growth_vec, 0);
  } else if (name == "delayed_growth_all") {
    EMU_RUN(dde_dopri_model_delayed_growth_all, 0);
  } else {
    g_message = "emu: unknown model " + name;
    return 1;
  }
  return 0;
}

} /* extern "C" */
#pragma GCC visibility pop
