/* test_models.cu -- TEST INFRASTRUCTURE: the reference's own fixture models as an
 * out-of-tree dde_b200 model library (tests/models/libdde_test_models.so), built by
 * tests/models/build.sh against include/ only -- the way a user's model library is built
 * (the reference loads a model .so at run time, R/common.R:89-98,
 * inst/include/dde/dde.c:10-78).  Loading the library registers the models by name
 * (dde_register_model); the GPU parity tests then run them through the C-ABI of the
 * product library, which contains none of them.
 */
#include <dde_b200/dde_model.cuh>

#include "difeq_fixtures.h"
#include "growth.h"
#include "seir.h"
#include "mg_wrong_lag.h"

/* a model with a (wrong) lag declaration: compiled once more inside a struct, like any
   model registered with DDE_REGISTER_DOPRI_LAGS (dde_model_scoped.inc) */
#define DDE_SCOPED_MODEL mg_wrong_lag
#define DDE_SCOPED_SOURCE "mg_wrong_lag.h"
#include <dde_b200/dde_model_scoped.inc>

/*                         name            deriv           n  n_parms  output        n_out */
DDE_REGISTER_DOPRI_OUTPUT(seir,           seir,           4, -1,      seir_output,   1)
DDE_REGISTER_DOPRI_OUTPUT(seir_int,       seir_int,       4, -1,      seir_output,   1)
/* the reference's growth fixtures come with five event functions */
DDE_EVENT_TABLE(growth_events, identity, double_variables, halve_variables, double_parameters,
                halve_parameters)
DDE_REGISTER_DOPRI_EVENTS(exponential,    exponential,    0, -1,      DDE_NO_OUTPUT, 0, growth_events, 5)
DDE_REGISTER_DOPRI_EVENTS(linear,         linear,         0, -1,      DDE_NO_OUTPUT, 0, growth_events, 5)
/* the same right-hand side with n fixed at compile time: small models without history run
   the unrolled kernel, whose rows leave by sector (1, 2 and 4 doubles per row) */
DDE_REGISTER_DOPRI(       exponential1,   exponential,    1, 1)
DDE_REGISTER_DOPRI(       exponential2,   exponential,    2, 2)
DDE_REGISTER_DOPRI(       exponential4,   exponential,    4, 4)
DDE_REGISTER_DOPRI(       delayed_growth, delayed_growth, 0, -1)
DDE_REGISTER_DOPRI(       delayed_growth_vec, delayed_growth_vec, 0, -1)
DDE_REGISTER_DOPRI(       delayed_growth_all, delayed_growth_all, 0, -1)
DDE_REGISTER_DOPRI_LAGS(  mg_wrong_lag,   mg_wrong_lag,   1, 3,       mg_wrong_lag_lags, 1)

/* three of the look-back fixtures compiled in scope as well (yprev_1 / yprev_all / yprev_vec
   then read the history state from registers, difeq_kernels.cuh: DifeqScope); the others
   stay on the global functions, so that the parity tests cover both */
#define DDE_SCOPED_MODEL fibonacci
#define DDE_SCOPED_SOURCE "difeq_fixtures.h"
#define DDE_SCOPED_BASE ::dde::DifeqScope
#include <dde_b200/dde_model_scoped.inc>
#define DDE_SCOPED_MODEL fib_all
#define DDE_SCOPED_SOURCE "difeq_fixtures.h"
#define DDE_SCOPED_BASE ::dde::DifeqScope
#include <dde_b200/dde_model_scoped.inc>
#define DDE_SCOPED_MODEL fib_vec
#define DDE_SCOPED_SOURCE "difeq_fixtures.h"
#define DDE_SCOPED_BASE ::dde::DifeqScope
#include <dde_b200/dde_model_scoped.inc>

/*                         name       target     n  n_parms n_out */
DDE_REGISTER_DIFEQ(       logistic,  logistic,  0, -1,     0)
DDE_REGISTER_DIFEQ(       additive,  additive,  0, -1,     -1)
DDE_REGISTER_DIFEQ_SCOPED(fibonacci, fibonacci, 5, -1,     0)
DDE_REGISTER_DIFEQ(       fib_out,   fib_out,   1, -1,     1)
DDE_REGISTER_DIFEQ_SCOPED(fib_all,   fib_all,   0, -1,     0)
DDE_REGISTER_DIFEQ_SCOPED(fib_vec,   fib_vec,   0, -1,     0)
