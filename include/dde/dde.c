/* <dde/dde.c> -- where the reference asks a model library to include this file exactly
 * once (/root/reference/inst/include/dde/dde.c:1-78: the R_GetCCallable trampolines of
 * ylag_* / yprev_*), a dde_b200 model translation unit has nothing to add: the history
 * functions are __device__ functions of <dde_b200/dde_model.cuh> and the model registers
 * itself with a DDE_REGISTER_* line.  Kept so that a reference model source compiles
 * unchanged apart from the DDE_MODEL_FN decoration (tools/decorate_model.py). */
