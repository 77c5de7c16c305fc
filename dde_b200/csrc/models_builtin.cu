/* models_builtin.cu -- the models compiled into libdde_b200.so: the BASELINE.json
 * workloads, nothing else.  (The reference's own fixture models -- seir, the growth
 * family, logistic, fibonacci ... -- are a model library of their own,
 * tests/models/libdde_test_models.so, built out of tree against include/ the way a
 * user's model library is: INTEGRATION.md.)
 *
 * The same sources (dde_b200/models/) are compiled as plain C against the reference
 * solver for the oracle, so host and device run character-identical model code.
 */
#include <dde_b200/dde_model.cuh>

#include "../models/lorenz.h"
#include "../models/mackey_glass.h"
#include "../models/rossler.h"
#include "../models/sir_delay.h"

/* the Mackey-Glass model declares its lag: compiled once more inside a struct, for the
   kernel that answers the look-up itself (dopri_dde1_kernel.cuh) */
#define DDE_SCOPED_MODEL mackey_glass
#define DDE_SCOPED_SOURCE "mackey_glass.h" /* found through -Idde_b200/models */
#include <dde_b200/dde_model_scoped.inc>

/*                         name            deriv           n  n_parms  output        n_out */
DDE_REGISTER_DOPRI_OUTPUT(lorenz,         lorenz,         3, 3,       lorenz_output, 2)
DDE_REGISTER_DOPRI(       rossler,        rossler,        3, 3)
DDE_REGISTER_DOPRI_LAGS(  mackey_glass,   mackey_glass,   1, 3,       mackey_glass_lags, 1)

/* the delayed SIR map compiled once more in scope too: its look-back reads the history
   state from registers (difeq_kernels.cuh: DifeqScope) */
#define DDE_SCOPED_MODEL sir_delay
#define DDE_SCOPED_SOURCE "sir_delay.h"
#define DDE_SCOPED_BASE ::dde::DifeqScope
#include <dde_b200/dde_model_scoped.inc>

/*                         name       target     n  n_parms n_out */
DDE_REGISTER_DIFEQ_SCOPED(sir_delay, sir_delay, 3, 2,      0)

#ifdef DDE_LAG_STATS
/* debug build only: read and clear the lag-window counters of this module */
extern "C" void dde_debug_lag_stats(unsigned long long *out) {
  unsigned long long zero[16] = {0};
  cudaMemcpyFromSymbol(out, dde::dde_lag_stats, sizeof zero);
  cudaMemcpyToSymbol(dde::dde_lag_stats, zero, sizeof zero);
}
#endif
