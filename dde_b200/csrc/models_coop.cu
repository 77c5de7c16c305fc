/* models_coop.cu -- the large models compiled into libdde_b200.so: one thread
 * block per trajectory (dopri_coop_kernel.cuh).  Same sources as the oracle
 * compiles as sequential C. */
#include "dde_model_coop.cuh"

#include "../models/growth.h"
#include "../models/seir_age.h"

/*                       name      deriv     threads components n_parms */
DDE_REGISTER_DOPRI_COOP(seir_age, seir_age, 256,    2,         3)
/* one warp per trajectory, n <= 128, parameters r[n] */
DDE_REGISTER_DOPRI_COOP(exponential_wide,    exponential_wide,    32, 4, -1)
DDE_REGISTER_DOPRI_COOP(delayed_growth_wide, delayed_growth_wide, 32, 4, -1)
