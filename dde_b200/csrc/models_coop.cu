/* models_coop.cu -- the large model compiled into libdde_b200.so: one thread block per
 * trajectory (dopri_coop_kernel.cuh).  Same source as the oracle compiles as sequential C. */
#include <dde_b200/dde_model_coop.cuh>

#include "../models/seir_age.h"

#ifndef DDE_SEIR_AGE_THREADS
#define DDE_SEIR_AGE_THREADS 256
#endif
/*                       name      deriv     threads               components                   n_parms */
DDE_REGISTER_DOPRI_COOP(seir_age, seir_age, DDE_SEIR_AGE_THREADS, (512 / DDE_SEIR_AGE_THREADS), 3)
