/* models_builtin.cu -- the models compiled into libdde_b200.so.
 *
 * The same sources (dde_b200/models/) are compiled as plain C against
 * the reference solver for the oracle, so host and device run
 * character-identical model code.  These are the reference's own fixture
 * models restated (see each file's header) plus the BASELINE.json workloads.
 */
#include "dde_model.cuh"

#include "../models/difeq_models.h"
#include "../models/growth.h"
#include "../models/lorenz.h"
#include "../models/mackey_glass.h"
#include "../models/rossler.h"
#include "../models/seir.h"

/* the Mackey-Glass model declares its lag: compiled once more inside a struct, for the
   kernel that answers the look-up itself (dopri_dde1_kernel.cuh) */
#define DDE_SCOPED_MODEL mackey_glass
#define DDE_SCOPED_SOURCE "../models/mackey_glass.h"
#include "dde_model_scoped.inc"

/*                         name            deriv           n  n_parms  output        n_out */
DDE_REGISTER_DOPRI_OUTPUT(lorenz,         lorenz,         3, 3,       lorenz_output, 2)
DDE_REGISTER_DOPRI(       rossler,        rossler,        3, 3)
DDE_REGISTER_DOPRI_LAGS(  mackey_glass,   mackey_glass,   1, 3,       mackey_glass_lags, 1)
DDE_REGISTER_DOPRI_OUTPUT(seir,           seir,           4, -1,      seir_output,   1)
DDE_REGISTER_DOPRI_OUTPUT(seir_int,       seir_int,       4, -1,      seir_output,   1)
/* the reference's growth fixtures come with five event functions */
DDE_EVENT_TABLE(growth_events, identity, double_variables, halve_variables, double_parameters,
                halve_parameters)
DDE_REGISTER_DOPRI_EVENTS(exponential,    exponential,    0, -1,      DDE_NO_OUTPUT, 0, growth_events, 5)
DDE_REGISTER_DOPRI_EVENTS(linear,         linear,         0, -1,      DDE_NO_OUTPUT, 0, growth_events, 5)
DDE_REGISTER_DOPRI(       delayed_growth, delayed_growth, 0, -1)
DDE_REGISTER_DOPRI(       delayed_growth_vec, delayed_growth_vec, 0, -1)

/*                  name       target     n  n_parms n_out */
DDE_REGISTER_DIFEQ(logistic,  logistic,  0, -1,     0)
DDE_REGISTER_DIFEQ(additive,  additive,  0, -1,     -1)
DDE_REGISTER_DIFEQ(fibonacci, fibonacci, 5, -1,     0)
DDE_REGISTER_DIFEQ(fib_out,   fib_out,   1, -1,     1)
DDE_REGISTER_DIFEQ(fib_all,   fib_all,   0, -1,     0)
DDE_REGISTER_DIFEQ(fib_vec,   fib_vec,   0, -1,     0)
DDE_REGISTER_DIFEQ(sir_delay, sir_delay, 3, 2,      0)

#ifdef DDE_LAG_STATS
/* debug build only: read and clear the lag-window counters of this module */
extern "C" void dde_debug_lag_stats(unsigned long long *out) {
  unsigned long long zero[16] = {0};
  cudaMemcpyFromSymbol(out, dde::dde_lag_stats, sizeof zero);
  cudaMemcpyToSymbol(dde::dde_lag_stats, zero, sizeof zero);
}
#endif
