/* dde_b200 -- C-ABI of the B200-native batched integrator.
 *
 * This is the drop-in boundary for the hot path of mrc-ide/dde: everything
 * below is plain C (pointers, sizes, PODs; no torch / CUDA types) and is what
 * the reference's R<->C glue (src/r_dopri.c, src/r_difeq.c) would bind in
 * place of its serial calls.  Every entry point names the reference
 * interface it replaces.  Per-trajectory semantics equal one call of the
 * reference function on that trajectory's inputs.
 *
 * Array layouts are the reference's (column-major R arrays):
 *   y0      [n      x n_traj]   n fastest            (src/r_difeq.c:9-17)
 *   parms   [n_parms x n_traj]  or one shared vector (parms_stride == 0;
 *                               src/r_difeq.c:40 passes one `data` pointer)
 *   y_out   [n      x nt x n_traj]                   (src/r_difeq.c:77,
 *   out     [n_out  x nt x n_traj]                    src/r_dopri.c:168-176)
 *   nt = n_times (return_initial) or n_times - 1     (src/r_dopri.c:124)
 *
 * Batch-level failures (bad arguments, CUDA errors, unknown model) return a
 * non-zero int and set dde_last_error().  Per-trajectory solver failures do
 * NOT abort the batch: they are reported in code[] with the reference's
 * return_code values (src/dopri.h:24-35) -- the reference raises Rf_error at
 * that point (src/r_dopri.c:264-297); dde_dopri_strerror() reproduces those
 * messages.
 */
#ifndef DDE_B200_H
#define DDE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- enums ------------------------------------------------------------ */

/* dopri_method, src/dopri.h:37-40 */
#define DDE_DOPRI_5 0
#define DDE_DOPRI_853 1

/* return_code, src/dopri.h:24-35 (same values) */
#define DDE_ERR_ZERO_TIME_DIFFERENCE (-1)
#define DDE_ERR_INCONSISTENT_TIME (-2)
#define DDE_ERR_TOO_MANY_STEPS (-3)
#define DDE_ERR_STEP_SIZE_TOO_SMALL (-4)
#define DDE_ERR_STEP_SIZE_VANISHED (-5)
#define DDE_ERR_YLAG_FAIL (-6)
#define DDE_ERR_STIFF (-7)
#define DDE_NOT_SET 0
#define DDE_OK_COMPLETE 1
#define DDE_OK_INTERRUPTED 2
/* Paths on which the reference calls Rf_error directly, long-jumping out of
 * the solver; a batch needs a per-trajectory code instead. */
#define DDE_ERR_NONFINITE_DERIV (-8) /* src/dopri.c:331-336 */
#define DDE_ERR_DIFEQ_HISTORY (-9)   /* src/difeq.c:267-270 */
/* dopri_continue from a time the trajectory did not stop at (or from a failed
 * solve): "Incorrect initial time on integration restart", src/r_dopri.c:215-217 */
#define DDE_ERR_RESTART (-10)
/* an extension: a model registered with its lags declared (DDE_REGISTER_DOPRI_LAGS,
 * include/dde_b200/dde_model_def.h) made a history look-up it did not declare */
#define DDE_ERR_LAG_UNDECLARED (-11)

/* index into stats[4 x n_traj]; order of the `statistics` attribute,
 * src/r_dopri.c:415-422 */
#define DDE_STAT_N_EVAL 0
#define DDE_STAT_N_STEP 1
#define DDE_STAT_N_ACCEPT 2
#define DDE_STAT_N_REJECT 3

/* ---- controls --------------------------------------------------------- */

/* The control fields of dopri_data (src/dopri.h:133-146,158) plus the
 * allocation-time arguments of dopri_data_alloc (src/dopri.h:163-168). */
typedef struct dde_dopri_control {
  double atol;              /* default 1e-6, src/dopri.c:64 */
  double rtol;              /* default 1e-6, src/dopri.c:65 */
  double step_size_min;     /* clamped to >= DBL_EPSILON, src/r_dopri.c:160 */
  double step_size_max;     /* clamped to <= DBL_MAX,     src/r_dopri.c:161 */
  double step_size_initial; /* 0 = choose (dopri_h_init), src/dopri.c:527-530 */
  uint64_t step_max_n;      /* default 100000, src/dopri.c:83 */
  uint64_t stiff_check;     /* 0 = never, src/dopri.c:213-215 */
  uint64_t n_history;       /* ring capacity in accepted steps, src/dopri.c:53 */
  int32_t method;           /* DDE_DOPRI_5 | DDE_DOPRI_853 */
  int32_t step_size_min_allow; /* src/dopri.c:351-354 */
  int32_t grow_history;     /* grow the ring instead of overwriting (src/dopri.c:92-95): the run
                               is repeated with larger rings until none has wrapped */
  int32_t return_initial;   /* src/dopri.c:320-327 */
} dde_dopri_control;

/* Fill *ctl with the reference defaults for `method`
 * (src/dopri.c:64-87, R/dopri.R:293-310). */
void dde_dopri_control_default(dde_dopri_control *ctl, int32_t method);

/* ---- library / models ------------------------------------------------- */

const char *dde_last_error(void);
const char *dde_version(void);

/* Compiled models known to the library (the analogue of the reference's
 * dyn.load'ed model .so + getNativeSymbolInfo, R/common.R:89-98).  A model
 * is one translation unit holding `deriv`/`output`/`difeq_target` functions
 * with the reference's model ABI (src/dopri.h:49-53, src/difeq.h:19-21)
 * compiled as __device__ code together with the kernels. */
size_t dde_model_count(void);
const char *dde_model_name(size_t i);
/* returns 0 and fills the outputs if the model exists; kind bit 0 = dopri
 * (deriv), bit 1 = difeq (target); n == 0 means "any n" */
int dde_model_info(const char *model, int32_t *kind, size_t *n,
                   size_t *n_parms, size_t *n_out);

/* The message r_dopri_error() raises for `code` (src/r_dopri.c:264-297).
 * `t` is the trajectory's time at failure, `n_history` its ring capacity
 * (ERR_YLAG_FAIL has two messages, src/r_dopri.c:284-289). */
int dde_dopri_strerror(int32_t code, double t, uint64_t n_history, char *buf,
                       size_t len);

/* ---- dopri: one-shot batched solve, HOST buffers ----------------------- */

/* Replaces, for every trajectory j in [0, n_traj):
 *   dopri_data_alloc(model.deriv, n, model.output, n_out, parms_j, method,
 *                    n_history, false, VERBOSE_QUIET, R_NilValue)
 *   + control assignment                       (src/r_dopri.c:148-166)
 *   dopri_integrate(obj, y0_j, times, n_times, tcrit, n_tcrit, is_event=all
 *                   false, NULL, y_out_j, out_j, return_initial)
 *                                              (src/dopri.c:291-514)
 *   + statistics / step_size extraction        (src/r_dopri.c:412-428)
 *   dopri_data_free(obj)
 * Copies inputs host->device, integrates all trajectories concurrently on
 * the current CUDA device, copies results device->host.  Any output pointer
 * may be NULL (not returned).  y_out/out rows that the reference would not
 * have written (failed trajectories) are zero, like the zero-initialised R
 * matrix (src/r_dopri.c:169). */
int dde_dopri_batch(const char *model, const dde_dopri_control *ctl, size_t n,
                    size_t n_out, size_t n_traj, const double *y0,
                    const double *parms, size_t n_parms, size_t parms_stride,
                    const double *times, size_t n_times, const double *tcrit,
                    size_t n_tcrit, double *y_out, double *out,
                    int32_t *stats,    /* [4 x n_traj] */
                    int32_t *code,     /* [n_traj] obj->code */
                    double *t_last,    /* [n_traj] obj->t on return */
                    double *step_size  /* [n_traj] obj->step_size_initial on
                                          return, src/dopri.c:463 */);

/* ---- dopri: device-resident batch handle ------------------------------- */

/* Opaque; owns all device state of n_traj concurrent solver objects: the
 * analogue of n_traj dopri_data structs (src/dopri.h:55-161). */
typedef struct dde_dopri_batch_s dde_dopri_batch_t;

/* dopri_data_alloc (src/dopri.c:8-90) for n_traj objects on the current
 * device.  NULL on failure. */
dde_dopri_batch_t *dde_dopri_batch_alloc(const char *model,
                                         const dde_dopri_control *ctl,
                                         size_t n, size_t n_out,
                                         size_t n_parms, size_t n_traj);
/* dopri_data_free (src/dopri.c:221-235) */
void dde_dopri_batch_free(dde_dopri_batch_t *b);

/* Host -> device copy of initial state and parameters (layouts above).
 * Either pointer may be NULL to keep what is resident. */
int dde_dopri_batch_upload(dde_dopri_batch_t *b, const double *y0,
                           const double *parms, size_t parms_stride);
/* Output times and critical times (shared by all trajectories, like
 * `steps` in difeq_replicate); validated as dopri_data_reset does
 * (src/dopri.c:168-183). */
int dde_dopri_batch_set_times(dde_dopri_batch_t *b, const double *times,
                              size_t n_times, const double *tcrit,
                              size_t n_tcrit);
/* Events (is_event / events of dopri_integrate, src/dopri.h:170-171; R/events.R):
 * event[i] >= 0 applies the model's event function number event[i] -- in the
 * order of its DDE_EVENT_TABLE -- to the state when critical time tcrit[i] is
 * reached (src/dopri.c:476-482); -1 makes tcrit[i] a plain stop.  Call after
 * dde_dopri_batch_set_times, which clears the events.  An event function may
 * change the trajectory's parameters; the change lasts for the run. */
int dde_dopri_batch_set_events(dde_dopri_batch_t *b, const int32_t *event,
                               size_t n_tcrit);
/* The order in which the device's persistent threads take trajectories (default: index
 * order; no counterpart in the reference, whose loop over replicates is serial,
 * src/r_difeq.c:94-103).  Results do not depend on it; the time of a batch does -- it ends
 * when its last trajectory does.  order: a permutation of 0 .. n_traj - 1, or NULL for index
 * order.  Thread-per-trajectory models. */
int dde_dopri_batch_set_order(dde_dopri_batch_t *b, const uint32_t *order);
/* ... longest first by the step attempts of this handle's LAST run: for a handle run again
 * on the same or slowly changing inputs (a fitting loop).  Built on the device,
 * asynchronous.  In force until dde_dopri_batch_set_order(b, NULL). */
int dde_dopri_batch_order_by_last_run(dde_dopri_batch_t *b);
/* dopri_integrate for every trajectory, from the resident y0/parms.
 * Asynchronous on the handle's stream (blocking with ctl.grow_history). */
int dde_dopri_batch_run(dde_dopri_batch_t *b);
/* dopri_continue (R/dopri.R:469-511, src/r_dopri.c:193-262): integrate on from
 * where the last run of this handle ended -- the checkpoint/resume of the
 * reference, which keeps the live solver object (`restartable = TRUE`).  The
 * rings, the final states and the step sizes the last run would have taken
 * next are resident; times[0] must equal the time each trajectory stopped at.
 * y == NULL keeps the final states; a new y [n x n_traj] also becomes the
 * pre-history (src/dopri.c:157).  A trajectory that did not complete, or
 * stopped at another time, gets DDE_ERR_RESTART and keeps its state.
 * Asynchronous like dde_dopri_batch_run; results through _download.  With
 * ctl.grow_history the call blocks: it keeps a copy of the state it starts
 * from and, if a ring wraps, repeats the continuation from that copy in
 * larger rings. */
int dde_dopri_batch_continue(dde_dopri_batch_t *b, const double *y,
                             const double *times, size_t n_times,
                             const double *tcrit, size_t n_tcrit);
/* dopri_data_copy (src/dopri.c:92-134; Cdopri_copy): an independent handle
 * with the same device state, rings included.  NULL on failure. */
dde_dopri_batch_t *dde_dopri_batch_copy(dde_dopri_batch_t *b);
/* Wait for the handle's stream. */
int dde_dopri_batch_sync(dde_dopri_batch_t *b);
/* Device -> host copy of results (any pointer may be NULL). Synchronises. */
int dde_dopri_batch_download(dde_dopri_batch_t *b, double *y_out, double *out,
                             int32_t *stats, int32_t *code, double *t_last,
                             double *step_size);
/* The same for trajectories first .. first + count - 1 only; the host arrays
 * hold `count` trajectories. */
int dde_dopri_batch_download_range(dde_dopri_batch_t *b, size_t first, size_t count,
                                   double *y_out, double *out, int32_t *stats,
                                   int32_t *code, double *t_last, double *step_size);
/* dde_dopri_batch_run + dde_dopri_batch_download in one blocking call, the
 * two overlapped: trajectories are integrated in index order, so the results
 * of each block of 32768 are copied to the host (on a second stream) as soon
 * as the kernel reports the block complete, while later ones are still being
 * integrated.  Same results as run + download; this is what the one-shot
 * dde_dopri_batch uses.  Pinned host buffers make the copies asynchronous. */
int dde_dopri_batch_run_download(dde_dopri_batch_t *b, double *y_out, double *out,
                                 int32_t *stats, int32_t *code, double *t_last,
                                 double *step_size);
/* Ring capacity in entries: ctl.n_history, or what grow_history made of it. */
uint64_t dde_dopri_batch_n_history(dde_dopri_batch_t *b);
/* Milliseconds the last dde_dopri_batch_run spent on the device (CUDA
 * events on the handle's stream); < 0 if none has completed. */
double dde_dopri_batch_last_ms(dde_dopri_batch_t *b);
/* Number of kernels the last dde_dopri_batch_run launched. */
int dde_dopri_batch_last_launches(dde_dopri_batch_t *b);
/* The `history` attribute for trajectory j: ring entries oldest first, each
 * (order*n + 2) doubles (src/r_dopri.c:401-410, src/dopri.h:94-126).  Returns
 * the number of entries, or < 0 on error; at most max_entries are copied. */
long dde_dopri_batch_history(dde_dopri_batch_t *b, size_t j, double *history,
                             size_t max_entries);
/* Raw device pointers for zero-copy consumers (torch tensors etc.):
 * which = 0 y_out, 1 out, 2 stats, 3 code, 4 t_last, 5 step_size, 6 y0,
 * 7 parms. */
void *dde_dopri_batch_device_ptr(dde_dopri_batch_t *b, int which);

/* ---- dopri_interpolate -------------------------------------------------- */

/* Batched equivalent of R/dopri.R:577-642: evaluate the dense-output
 * polynomial of a stored history at arbitrary times.
 *   history [history_len x n_entries x n_traj], history_len = order*n + 2
 *   t       [n_t]  (shared)
 *   y       [n_t x n x n_traj]  (R returns a length(t) x n matrix)
 * Returns non-zero (and touches nothing) if any t lies outside
 * [t_first, t_last + h_last] of any trajectory (R/dopri.R:596-600). */
int dde_dopri_interpolate_batch(const double *history, size_t history_len,
                                size_t n_entries, size_t n, size_t n_traj,
                                const double *t, size_t n_t, double *y);
/* The same for the trajectories of a batch handle, from the history rings where the
 * solver left them in HBM: no history travels to the host and back, the range check
 * runs on the device.  t [n_t] and y [n_t x n x n_traj] are host arrays.  A time
 * outside a trajectory's surviving history (entries the ring has overwritten are
 * gone, src/dopri.h:94-126) gives NaN for that trajectory and time, and the call
 * returns non-zero naming the first such trajectory's range (R/dopri.R:596-600). */
int dde_dopri_batch_interpolate(dde_dopri_batch_t *b, const double *t, size_t n_t,
                                double *y);

/* ---- difeq -------------------------------------------------------------- */

/* Replaces the replicate loop of r_difeq (src/r_difeq.c:94-103) and, per
 * replicate j, difeq_data_alloc + difeq_run (src/difeq.c:9-47,127-258):
 *   target(n, step, y, y_next, n_out, out, parms_j) iterated from steps[0]
 *   to steps[n_steps-1], rows stored at the requested steps.
 * Every replicate gets FRESH history state: the documented contract
 * (R/difeq_replicate.R:1-4: each replicate == an independent difeq() run),
 * not the reference's shared-ring artefact (SURVEY.md finding 5).
 *   steps [n_steps]  strictly increasing (src/difeq.c:89-93)
 *   y_out [n x nt x n_rep], out [n_out x nt x n_rep]
 *   code  [n_rep]: DDE_OK_COMPLETE or DDE_ERR_DIFEQ_HISTORY
 *   y_final [n x n_rep]: obj->y1 (src/difeq.c:244), may be NULL */
int dde_difeq_batch(const char *model, size_t n, size_t n_out, size_t n_rep,
                    const double *y0, const double *parms, size_t n_parms,
                    size_t parms_stride, const uint64_t *steps, size_t n_steps,
                    size_t n_history, int32_t return_initial, double *y_out,
                    double *out, int32_t *code, double *y_final);

typedef struct dde_difeq_batch_s dde_difeq_batch_t;
dde_difeq_batch_t *dde_difeq_batch_alloc(const char *model, size_t n,
                                         size_t n_out, size_t n_parms,
                                         size_t n_rep, size_t n_history);
void dde_difeq_batch_free(dde_difeq_batch_t *b);
int dde_difeq_batch_upload(dde_difeq_batch_t *b, const double *y0,
                           const double *parms, size_t parms_stride);
int dde_difeq_batch_set_steps(dde_difeq_batch_t *b, const uint64_t *steps,
                              size_t n_steps, int32_t return_initial);
int dde_difeq_batch_run(dde_difeq_batch_t *b);
/* The reference's `restartable = TRUE` (R/difeq.R, src/r_difeq.c:268-271): make
 * the following runs leave their history in HBM so that they can be continued
 * (an on-chip history is otherwise dropped when the kernel ends).  Off by
 * default; costs one write of the history per run. */
int dde_difeq_batch_restartable(dde_difeq_batch_t *b, int32_t on);
/* difeq_continue (src/r_difeq.c:114-165): iterate on from the step the last
 * run of this handle ended at; steps[0] must be that step ("Incorrect initial
 * step on simulation restart").  y == NULL continues from the final states
 * (obj->y1); the history is resident.  As in the reference the restart state is
 * not committed to the history a second time.  A replicate whose last run
 * failed gets DDE_ERR_RESTART. */
int dde_difeq_batch_continue(dde_difeq_batch_t *b, const double *y,
                             const uint64_t *steps, size_t n_steps,
                             int32_t return_initial);
/* difeq_data_copy (src/difeq.c:49-72; Cdifeq_copy) */
dde_difeq_batch_t *dde_difeq_batch_copy(dde_difeq_batch_t *b);
int dde_difeq_batch_sync(dde_difeq_batch_t *b);
int dde_difeq_batch_download(dde_difeq_batch_t *b, double *y_out, double *out,
                             int32_t *code, double *y_final);
/* The `history` attribute for replicate j (src/r_difeq.c:178-186): entries
 * {step, y[n], out[n_out]} oldest first, [max_entries x (1 + n + n_out)];
 * returns the number of entries the ring holds, < 0 on error.  Needs a
 * restartable run (dde_difeq_batch_restartable), which keeps the history in
 * HBM.  The out columns hold what the reference's do (src/difeq.c:203-213,
 * 249-255: the stored output row that was written last while the lagging
 * out_next pointer rested on the entry's slot, zeros elsewhere, NA in the
 * first entry) for a first run; after dde_difeq_batch_continue they are
 * NA_real_. */
long dde_difeq_batch_history(dde_difeq_batch_t *b, size_t j, double *history,
                             size_t max_entries);
double dde_difeq_batch_last_ms(dde_difeq_batch_t *b);

/* ---- multi-GPU ---------------------------------------------------------------- */

/* Trajectories are independent, so a batch shards over the GPUs of one box
 * with no communication between them: contiguous blocks of trajectory
 * indices, one host thread and one stream per device, every device copying
 * its results straight into its slice of the caller's arrays (layouts above,
 * so concatenation is free).  devices == NULL: all visible devices.  Same
 * arguments and semantics as dde_dopri_batch / dde_difeq_batch otherwise; the
 * serial loop they replace is src/r_difeq.c:94-103. */
int dde_device_count(void);
int dde_dopri_batch_sharded(const char *model, const dde_dopri_control *ctl,
                            size_t n, size_t n_out, size_t n_traj,
                            const double *y0, const double *parms,
                            size_t n_parms, size_t parms_stride,
                            const double *times, size_t n_times,
                            const double *tcrit, size_t n_tcrit, double *y_out,
                            double *out, int32_t *stats, int32_t *code,
                            double *t_last, double *step_size,
                            const int32_t *devices, size_t n_devices);
/* ... with the model's events at the critical times (dde_dopri_batch_set_events;
 * event == NULL: plain stops), src/dopri.c:476-491 */
int dde_dopri_batch_sharded_events(const char *model, const dde_dopri_control *ctl,
                                   size_t n, size_t n_out, size_t n_traj,
                                   const double *y0, const double *parms,
                                   size_t n_parms, size_t parms_stride,
                                   const double *times, size_t n_times,
                                   const double *tcrit, size_t n_tcrit,
                                   const int32_t *event, double *y_out,
                                   double *out, int32_t *stats, int32_t *code,
                                   double *t_last, double *step_size,
                                   const int32_t *devices, size_t n_devices);
int dde_difeq_batch_sharded(const char *model, size_t n, size_t n_out,
                            size_t n_rep, const double *y0,
                            const double *parms, size_t n_parms,
                            size_t parms_stride, const uint64_t *steps,
                            size_t n_steps, size_t n_history,
                            int32_t return_initial, double *y_out, double *out,
                            int32_t *code, double *y_final,
                            const int32_t *devices, size_t n_devices);
/* The same batch as a handle: one dde_dopri_batch_t per device, allocated once
 * (rings included), so that repeated solves of one batch shape -- a parameter
 * scan, the strong-scaling measurement of bench.py -- pay for allocation once.
 * upload / set_times / set_events address every shard in turn (asynchronous
 * copies on each shard's stream); run_download integrates all shards
 * concurrently, one host thread per device, each streaming its results into its
 * slice of the caller's arrays, and returns when all have finished.  last_ms:
 * the slowest shard's device time of the last run. */
typedef struct dde_dopri_shards_s dde_dopri_shards_t;
dde_dopri_shards_t *dde_dopri_shards_alloc(const char *model,
                                           const dde_dopri_control *ctl, size_t n,
                                           size_t n_out, size_t n_parms,
                                           size_t n_traj, const int32_t *devices,
                                           size_t n_devices);
void dde_dopri_shards_free(dde_dopri_shards_t *h);
size_t dde_dopri_shards_count(dde_dopri_shards_t *h);
int dde_dopri_shards_upload(dde_dopri_shards_t *h, const double *y0,
                            const double *parms, size_t parms_stride);
int dde_dopri_shards_set_times(dde_dopri_shards_t *h, const double *times,
                               size_t n_times, const double *tcrit,
                               size_t n_tcrit);
int dde_dopri_shards_set_events(dde_dopri_shards_t *h, const int32_t *event,
                                size_t n_tcrit);
/* every shard's own order from its own last run; enable == 0: back to index order */
int dde_dopri_shards_order_by_last_run(dde_dopri_shards_t *h, int32_t enable);
int dde_dopri_shards_run_download(dde_dopri_shards_t *h, double *y_out,
                                  double *out, int32_t *stats, int32_t *code,
                                  double *t_last, double *step_size);
double dde_dopri_shards_last_ms(dde_dopri_shards_t *h);

/* The block of a batch of n_total that shard `rank` of `world` owns:
 * [*first, *first + *count).  Blocks differ in size by at most one. */
void dde_shard_range(size_t n_total, size_t world, size_t rank, size_t *first,
                     size_t *count);

/* ---- host memory ------------------------------------------------------------ */

/* Page-locked host buffers (cudaHostAlloc) so the upload/download calls above
 * are true asynchronous DMA.  Ordinary malloc'ed / R-owned memory works too,
 * through a staging copy. */
void *dde_host_alloc(size_t bytes);
void dde_host_free(void *p);

/* ---- math ---------------------------------------------------------------- */

/* The library's pow/exp, bit-identical restatements of glibc 2.39's
 * x86-64 FMA variants (the libm the reference links to on this platform;
 * call sites src/dopri.c:500,517,521,569).  Host build of the same source
 * that runs on the device; exported so tests can pin it against libm. */
double dde_pow_host(double x, double y);
double dde_exp_host(double x);
/* out[0] = x^y1, out[1] = x^y2 with the logarithm of x computed once (the dopri5 step-size
 * controller's two powers): each the bits dde_pow_host gives alone */
void dde_pow2_host(double x, double y1, double y2, double *out);

/* Measured FP64 issue ceiling of the current device in TFLOP/s: fma != 0 a
 * chain of DFMA (2 flop/instr), fma == 0 alternating DMUL/DADD (1 flop/instr,
 * what FMA-free code can reach).  Roofline denominator for bench.py. */
double dde_fp64_peak(int fma);

#ifdef __cplusplus
}
#endif

#endif /* DDE_B200_H */
