/* test_models_coop.cu -- TEST INFRASTRUCTURE: the block-per-trajectory fixtures of
 * tests/models/libdde_test_models.so (see test_models.cu). */
#include <dde_b200/dde_model_coop.cuh>

#include "growth.h"

/* one warp per trajectory, n <= 128, parameters r[n] */
/*                       name                 deriv                threads components n_parms */
DDE_REGISTER_DOPRI_COOP(exponential_wide,    exponential_wide,    32, 4, -1)
DDE_REGISTER_DOPRI_COOP(delayed_growth_wide, delayed_growth_wide, 32, 4, -1)
