/* models_coop.cu -- the large model compiled into libdde_b200.so: one thread block per
 * trajectory (dopri_coop_kernel.cuh).  Same source as the oracle compiles as sequential C. */
#include <dde_b200/dde_model_coop.cuh>

#include "../models/seir_age.h"

/*                       name      deriv     threads components n_parms */
DDE_REGISTER_DOPRI_COOP(seir_age, seir_age, 256,    2,         3)
