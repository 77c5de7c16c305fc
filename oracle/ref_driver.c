/* TEST INFRASTRUCTURE ONLY (oracle/_ref): batch driver around the UNMODIFIED
 * reference solver.
 *
 * oracle/Makefile compiles /root/reference/src/{dopri,dopri_5,dopri_853,
 * difeq}.c from where they lie, against oracle/shim (R.h, Rinternals.h,
 * ring/), together with this file and the shared model sources
 * (dde_b200/models/), into oracle/_ref/libdde_ref.so.  Nothing here is
 * on the product path: tests and bench.py's CPU-baseline legs load it as
 * the checker / the thing timed beside the GPU.
 *
 * Each trajectory gets a FRESH solver object (dopri_data_reset does not
 * clear the ring, src/dopri.c:185-191), the controls are
 * assigned exactly as the R glue does (src/r_dopri.c:158-166), and
 * Rf_error (a longjmp in R) is caught per trajectory and turned into the
 * DDE_ERR_* codes of include/dde_b200.h.  The entry points take the same
 * arguments as dde_dopri_batch / dde_difeq_batch so tests can drive both
 * with one argument list.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <float.h>
#include <setjmp.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "dopri.h"
#include "difeq.h"

#include "dde_b200.h"

/* ---- R shim runtime ---------------------------------------------------- */

SEXP R_NilValue = NULL;

static jmp_buf shim_jmp;
static int shim_jmp_armed = 0;
static char shim_msg[512];

typedef struct transient {
  struct transient *next;
} transient;
static transient *shim_transient = NULL;

void *dde_shim_calloc(size_t n, size_t size) {
  void *p = calloc(n ? n : 1, size ? size : 1);
  if (!p) {
    fprintf(stderr, "oracle shim: out of memory\n");
    abort();
  }
  return p;
}

void dde_shim_free(void *p) { free(p); }

char *R_alloc(size_t n, int size) {
  transient *t = (transient *) calloc(1, sizeof(transient) + n * (size_t) size + 8);
  t->next = shim_transient;
  shim_transient = t;
  return (char *) (t + 1);
}

static void shim_release_transient(void) {
  while (shim_transient) {
    transient *t = shim_transient;
    shim_transient = t->next;
    free(t);
  }
}

double dde_shim_na_real(void) {
  /* R's NA_real_: quiet NaN whose low word is 1954 */
  const uint64_t bits = 0x7FF00000000007A2ULL;
  double x;
  memcpy(&x, &bits, sizeof x);
  return x;
}

void Rf_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(shim_msg, sizeof shim_msg, fmt, ap);
  va_end(ap);
  if (shim_jmp_armed) {
    longjmp(shim_jmp, 1);
  }
  fprintf(stderr, "oracle shim: uncaught Rf_error: %s\n", shim_msg);
  abort();
}

void Rprintf(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vfprintf(stdout, fmt, ap);
  va_end(ap);
}

static void shim_unreachable(const char *what) {
  fprintf(stderr, "oracle shim: %s called (verbose callbacks are not supported)\n", what);
  abort();
}
SEXP VECTOR_ELT(SEXP x, long i) { shim_unreachable("VECTOR_ELT"); return NULL; }
SEXP Rf_ScalarReal(double x) { shim_unreachable("Rf_ScalarReal"); return NULL; }
SEXP Rf_allocVector(unsigned int type, long n) { shim_unreachable("Rf_allocVector"); return NULL; }
double *REAL(SEXP x) { shim_unreachable("REAL"); return NULL; }
SEXP Rf_lang4(SEXP a, SEXP b, SEXP c, SEXP d) { shim_unreachable("Rf_lang4"); return NULL; }
SEXP Rf_eval(SEXP call, SEXP env) { shim_unreachable("Rf_eval"); return NULL; }

const char *ref_last_message(void) { return shim_msg; }

/* The reference's own fixtures end with `#include <dde/dde.c>`, which
 * resolves the lag functions through R_GetCCallable("dde", name)
 * (inst/include/dde/dde.c:10-16, registered at src/zzz.c:27-35). */
typedef void *(*DL_FUNC)(void);
DL_FUNC R_GetCCallable(const char *package, const char *name) {
  void *self = dlopen(NULL, RTLD_NOW);
  void *sym = dlsym(RTLD_DEFAULT, name);
  (void) package;
  (void) self;
  if (!sym) {
    Rf_error("R_GetCCallable: %s not found", name);
  }
  return (DL_FUNC) sym;
}

/* ---- models ------------------------------------------------------------ */

#include "../dde_b200/models/lorenz.h"
#include "../dde_b200/models/rossler.h"
#include "../dde_b200/models/mackey_glass.h"
#include "../tests/models/seir.h"
#include "../tests/models/growth.h"
#include "../tests/models/difeq_fixtures.h"
#include "../dde_b200/models/sir_delay.h"
#include "../dde_b200/models/seir_age.h"

typedef struct {
  const char *name;
  deriv_func *deriv;
  output_func *output;
  difeq_target *target;
  event_func *const *events; /* in the order of the model's DDE_EVENT_TABLE */
  int n_events;
} model_entry;

static event_func *const growth_event_table[5] = {identity, double_variables, halve_variables,
                                                  double_parameters, halve_parameters};

/* Events for the dopri calls that follow (cleared by n = 0): event_index[i] >= 0
 * applies the model's event number event_index[i] at tcrit[i], exactly the
 * is_event / events arguments of dopri_integrate (src/r_dopri.c:77-103). */
static int32_t event_index[64];
static size_t n_event_index = 0;
void ref_set_events(const int32_t *index, size_t n) {
  n_event_index = n < 64 ? n : 64;
  for (size_t i = 0; i < n_event_index; ++i) {
    event_index[i] = index[i];
  }
}

static const model_entry models[] = {
    {"lorenz", lorenz, lorenz_output, NULL},
    {"rossler", rossler, NULL, NULL},
    {"mackey_glass", mackey_glass, NULL, NULL},
    {"seir", seir, seir_output, NULL},
    {"seir_int", seir_int, seir_output, NULL},
    {"exponential", exponential, NULL, NULL, growth_event_table, 5},
    {"linear", linear, NULL, NULL, growth_event_table, 5},
    {"delayed_growth", delayed_growth, NULL, NULL},
    {"delayed_growth_vec", delayed_growth_vec, NULL, NULL},
    {"delayed_growth_all", delayed_growth_all, NULL, NULL},
    {"exponential_wide", exponential_wide, NULL, NULL},
    {"delayed_growth_wide", delayed_growth_wide, NULL, NULL},
    {"seir_age", seir_age, NULL, NULL},
    {"logistic", NULL, NULL, logistic},
    {"additive", NULL, NULL, additive},
    {"fibonacci", NULL, NULL, fibonacci},
    {"fib_out", NULL, NULL, fib_out},
    {"fib_all", NULL, NULL, fib_all},
    {"fib_vec", NULL, NULL, fib_vec},
    {"sir_delay", NULL, NULL, sir_delay},
};

/* "name" looks up the table above; "path.so:deriv[:output]" loads one of the
 * reference's own fixture files compiled by oracle/Makefile. */
static int find_model(const char *name, model_entry *out) {
  const char *colon = strchr(name, ':');
  if (colon) {
    char path[1024], sym[256], sym2[256] = "";
    const size_t len = (size_t) (colon - name);
    if (len >= sizeof path) {
      return -1;
    }
    memcpy(path, name, len);
    path[len] = 0;
    const char *rest = colon + 1;
    const char *colon2 = strchr(rest, ':');
    if (colon2) {
      snprintf(sym, sizeof sym, "%.*s", (int) (colon2 - rest), rest);
      snprintf(sym2, sizeof sym2, "%s", colon2 + 1);
    } else {
      snprintf(sym, sizeof sym, "%s", rest);
    }
    void *h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
      snprintf(shim_msg, sizeof shim_msg, "dlopen failed: %s", dlerror());
      return -1;
    }
    void *f = dlsym(h, sym);
    if (!f) {
      snprintf(shim_msg, sizeof shim_msg, "symbol %s not found", sym);
      return -1;
    }
    out->name = name;
    out->deriv = (deriv_func *) f;
    out->target = (difeq_target *) f;
    out->output = sym2[0] ? (output_func *) dlsym(h, sym2) : NULL;
    out->events = NULL;
    out->n_events = 0;
    return 0;
  }
  for (size_t i = 0; i < sizeof models / sizeof models[0]; ++i) {
    if (strcmp(models[i].name, name) == 0) {
      *out = models[i];
      return 0;
    }
  }
  snprintf(shim_msg, sizeof shim_msg, "unknown model '%s'", name);
  return -1;
}

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static double last_seconds = -1.0;
double ref_last_seconds(void) { return last_seconds; }

/* ---- dopri --------------------------------------------------------------- */

/* One trajectory through the reference; `history` (may be NULL) receives the
 * ring contents oldest-first as r_dopri_cleanup does (src/r_dopri.c:401-404). */
static void run_one_dopri(const model_entry *m, const dde_dopri_control *ctl,
                          size_t n, size_t n_out, const double *y0,
                          const double *parms, const double *times,
                          size_t n_times, const double *tcrit, size_t n_tcrit,
                          double *y_out, double *out, int32_t *stats,
                          int32_t *code, double *t_last, double *step_size,
                          double *history, size_t max_entries,
                          long *n_entries) {
  dopri_data *obj = dopri_data_alloc(
      m->deriv, n, n_out > 0 ? m->output : NULL, n_out, (void *) parms,
      ctl->method == DDE_DOPRI_853 ? DOPRI_853 : DOPRI_5, (size_t) ctl->n_history,
      ctl->grow_history != 0, VERBOSE_QUIET, R_NilValue);
  /* src/r_dopri.c:158-166 */
  obj->rtol = ctl->rtol;
  obj->atol = ctl->atol;
  obj->step_size_min = fmax(fabs(ctl->step_size_min), DBL_EPSILON);
  obj->step_size_max = fmin(fabs(ctl->step_size_max), DBL_MAX);
  obj->step_size_initial = ctl->step_size_initial;
  obj->step_max_n = (size_t) ctl->step_max_n;
  obj->step_size_min_allow = ctl->step_size_min_allow != 0;
  obj->stiff_check = (size_t) ctl->stiff_check;

  bool *is_event = (bool *) calloc(n_tcrit ? n_tcrit : 1, sizeof(bool));
  event_func **events = (event_func **) calloc(n_tcrit ? n_tcrit : 1, sizeof(event_func *));
  if (n_event_index == n_tcrit && m->events) {
    for (size_t i = 0; i < n_tcrit; ++i) {
      if (event_index[i] >= 0 && event_index[i] < m->n_events) {
        is_event[i] = true;
        events[i] = m->events[event_index[i]];
      }
    }
  }
  int32_t rc;
  shim_jmp_armed = 1;
  if (setjmp(shim_jmp) == 0) {
    dopri_integrate(obj, y0, times, n_times, tcrit, n_tcrit, is_event, events,
                    y_out, out, ctl->return_initial != 0);
    rc = (int32_t) obj->code;
  } else {
    /* the only Rf_error inside dopri_integrate, src/dopri.c:331-336 */
    rc = DDE_ERR_NONFINITE_DERIV;
  }
  shim_jmp_armed = 0;
  free(is_event);
  free(events);
  shim_release_transient();

  if (stats) {
    stats[DDE_STAT_N_EVAL] = (int32_t) obj->n_eval;
    stats[DDE_STAT_N_STEP] = (int32_t) obj->n_step;
    stats[DDE_STAT_N_ACCEPT] = (int32_t) obj->n_accept;
    stats[DDE_STAT_N_REJECT] = (int32_t) obj->n_reject;
  }
  if (code) {
    *code = rc;
  }
  if (t_last) {
    *t_last = obj->t;
  }
  if (step_size) {
    *step_size = obj->step_size_initial;
  }
  if (n_entries) {
    size_t used = ring_buffer_used(obj->history, false);
    *n_entries = (long) used;
    if (history) {
      if (used > max_entries) {
        used = max_entries;
      }
      ring_buffer_read(obj->history, history, used);
    }
  }
  dopri_data_free(obj);
}

static void run_dopri_range(const model_entry *m, const dde_dopri_control *ctl,
                            size_t n, size_t n_out, size_t j0, size_t j1,
                            const double *y0, const double *parms,
                            size_t parms_stride, const double *times,
                            size_t n_times, const double *tcrit, size_t n_tcrit,
                            double *y_out, double *out, int32_t *stats,
                            int32_t *code, double *t_last, double *step_size) {
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  for (size_t j = j0; j < j1; ++j) {
    run_one_dopri(m, ctl, n, n_out, y0 + j * n, parms + j * parms_stride, times,
                  n_times, tcrit, n_tcrit, y_out ? y_out + j * n * nt : NULL,
                  out ? out + j * n_out * nt : NULL, stats ? stats + 4 * j : NULL,
                  code ? code + j : NULL, t_last ? t_last + j : NULL,
                  step_size ? step_size + j : NULL, NULL, 0, NULL);
  }
}

int ref_dopri_batch(const char *model, const dde_dopri_control *ctl, size_t n,
                    size_t n_out, size_t n_traj, const double *y0,
                    const double *parms, size_t n_parms, size_t parms_stride,
                    const double *times, size_t n_times, const double *tcrit,
                    size_t n_tcrit, double *y_out, double *out, int32_t *stats,
                    int32_t *code, double *t_last, double *step_size) {
  model_entry m;
  (void) n_parms;
  if (find_model(model, &m) != 0 || !m.deriv) {
    return 1;
  }
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  /* the R glue hands the solver a zero-filled matrix, src/r_dopri.c:169 */
  if (y_out) {
    memset(y_out, 0, sizeof(double) * n * nt * n_traj);
  }
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt * n_traj);
  }
  /* y_out must exist for the solver to write into */
  double *scratch = NULL;
  if (!y_out) {
    scratch = (double *) calloc(n * nt + n_out * nt + 1, sizeof(double));
  }
  const double t0 = now_seconds();
  if (y_out) {
    run_dopri_range(&m, ctl, n, n_out, 0, n_traj, y0, parms, parms_stride, times,
                    n_times, tcrit, n_tcrit, y_out, out, stats, code, t_last,
                    step_size);
  } else {
    for (size_t j = 0; j < n_traj; ++j) {
      run_one_dopri(&m, ctl, n, n_out, y0 + j * n, parms + j * parms_stride,
                    times, n_times, tcrit, n_tcrit, scratch, scratch + n * nt,
                    stats ? stats + 4 * j : NULL, code ? code + j : NULL,
                    t_last ? t_last + j : NULL, step_size ? step_size + j : NULL,
                    NULL, 0, NULL);
    }
  }
  last_seconds = now_seconds() - t0;
  free(scratch);
  return 0;
}

/* Same, spread over n_proc forked processes: the reference keeps its active
 * solver in a static global (src/dopri.c:245), so threads are unsafe.
 * Results come back through shared anonymous mappings. */
int ref_dopri_batch_mp(const char *model, const dde_dopri_control *ctl,
                       size_t n, size_t n_out, size_t n_traj, const double *y0,
                       const double *parms, size_t n_parms, size_t parms_stride,
                       const double *times, size_t n_times, const double *tcrit,
                       size_t n_tcrit, double *y_out, double *out,
                       int32_t *stats, int32_t *code, double *t_last,
                       double *step_size, int n_proc) {
  model_entry m;
  (void) n_parms;
  if (find_model(model, &m) != 0 || !m.deriv) {
    return 1;
  }
  if (n_proc < 1) {
    n_proc = 1;
  }
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  const size_t sz_y = sizeof(double) * n * nt * n_traj;
  const size_t sz_o = sizeof(double) * n_out * nt * n_traj;
  const size_t sz_s = sizeof(int32_t) * 4 * n_traj;
  const size_t sz_c = sizeof(int32_t) * n_traj;
  const size_t sz_d = sizeof(double) * n_traj;
  const size_t total = sz_y + sz_o + sz_s + sz_c + 2 * sz_d + 64;
  char *shm = (char *) mmap(NULL, total, PROT_READ | PROT_WRITE,
                            MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (shm == MAP_FAILED) {
    snprintf(shim_msg, sizeof shim_msg, "mmap failed");
    return 1;
  }
  double *s_y = (double *) shm;
  double *s_o = (double *) (shm + sz_y);
  double *s_t = (double *) (shm + sz_y + sz_o);
  double *s_h = (double *) (shm + sz_y + sz_o + sz_d);
  int32_t *s_s = (int32_t *) (shm + sz_y + sz_o + 2 * sz_d);
  int32_t *s_c = (int32_t *) (shm + sz_y + sz_o + 2 * sz_d + sz_s);

  const double t0 = now_seconds();
  pid_t *pids = (pid_t *) calloc((size_t) n_proc, sizeof(pid_t));
  int failed = 0;
  for (int p = 0; p < n_proc; ++p) {
    const size_t j0 = n_traj * (size_t) p / (size_t) n_proc;
    const size_t j1 = n_traj * (size_t) (p + 1) / (size_t) n_proc;
    pid_t pid = fork();
    if (pid == 0) {
      run_dopri_range(&m, ctl, n, n_out, j0, j1, y0, parms, parms_stride, times,
                      n_times, tcrit, n_tcrit, s_y, n_out ? s_o : NULL, s_s, s_c,
                      s_t, s_h);
      _exit(0);
    }
    if (pid < 0) {
      failed = 1;
    }
    pids[p] = pid;
  }
  for (int p = 0; p < n_proc; ++p) {
    int status = 0;
    if (pids[p] > 0) {
      waitpid(pids[p], &status, 0);
      if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) {
        failed = 1;
      }
    }
  }
  last_seconds = now_seconds() - t0;
  free(pids);
  if (!failed) {
    if (y_out) memcpy(y_out, s_y, sz_y);
    if (out) memcpy(out, s_o, sz_o);
    if (stats) memcpy(stats, s_s, sz_s);
    if (code) memcpy(code, s_c, sz_c);
    if (t_last) memcpy(t_last, s_t, sz_d);
    if (step_size) memcpy(step_size, s_h, sz_d);
  } else {
    snprintf(shim_msg, sizeof shim_msg, "a worker process failed");
  }
  munmap(shm, total);
  return failed;
}

/* One trajectory, also returning the `history` attribute. */
long ref_dopri_history(const char *model, const dde_dopri_control *ctl, size_t n,
                       size_t n_out, const double *y0, const double *parms,
                       const double *times, size_t n_times, const double *tcrit,
                       size_t n_tcrit, double *y_out, double *out,
                       int32_t *stats, int32_t *code, double *t_last,
                       double *step_size, double *history, size_t max_entries) {
  model_entry m;
  if (find_model(model, &m) != 0 || !m.deriv) {
    return -1;
  }
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  memset(y_out, 0, sizeof(double) * n * nt);
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt);
  }
  long n_entries = 0;
  run_one_dopri(&m, ctl, n, n_out, y0, parms, times, n_times, tcrit, n_tcrit,
                y_out, out, stats, code, t_last, step_size, history, max_entries,
                &n_entries);
  return n_entries;
}

/* dopri_continue: every trajectory is integrated over times1, then -- the
 * solver object kept alive, as `restartable = TRUE` does (src/r_dopri.c:432-436)
 * -- over times2 exactly as r_dopri_continue does it (src/r_dopri.c:193-262):
 * y_new == NULL continues from obj->y.  The outputs are those of the SECOND
 * segment; a trajectory whose first segment failed, or that stopped at another
 * time than times2[0], gets DDE_ERR_RESTART. */
int ref_dopri_batch_continue(const char *model, const dde_dopri_control *ctl, size_t n,
                             size_t n_out, size_t n_traj, const double *y0,
                             const double *parms, size_t n_parms, size_t parms_stride,
                             const double *times1, size_t n_times1, const double *times2,
                             size_t n_times2, const double *y_new, const double *tcrit2,
                             size_t n_tcrit2, double *y_out, double *out, int32_t *stats,
                             int32_t *code, double *t_last, double *step_size) {
  model_entry m;
  (void) n_parms;
  if (find_model(model, &m) != 0 || !m.deriv) {
    return 1;
  }
  const size_t nt1 = ctl->return_initial ? n_times1 : n_times1 - 1;
  const size_t nt2 = ctl->return_initial ? n_times2 : n_times2 - 1;
  memset(y_out, 0, sizeof(double) * n * nt2 * n_traj);
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt2 * n_traj);
  }
  double *scratch = (double *) calloc(n * nt1 + n_out * nt1 + 1, sizeof(double));
  bool *is_event = (bool *) calloc(n_tcrit2 ? n_tcrit2 : 1, sizeof(bool));
  for (size_t j = 0; j < n_traj; ++j) {
    dopri_data *obj = dopri_data_alloc(
        m.deriv, n, n_out > 0 ? m.output : NULL, n_out, (void *) (parms + j * parms_stride),
        ctl->method == DDE_DOPRI_853 ? DOPRI_853 : DOPRI_5, (size_t) ctl->n_history,
        ctl->grow_history != 0, VERBOSE_QUIET, R_NilValue);
    obj->rtol = ctl->rtol;
    obj->atol = ctl->atol;
    obj->step_size_min = fmax(fabs(ctl->step_size_min), DBL_EPSILON);
    obj->step_size_max = fmin(fabs(ctl->step_size_max), DBL_MAX);
    obj->step_size_initial = ctl->step_size_initial;
    obj->step_max_n = (size_t) ctl->step_max_n;
    obj->step_size_min_allow = ctl->step_size_min_allow != 0;
    obj->stiff_check = (size_t) ctl->stiff_check;
    int32_t rc = DDE_ERR_RESTART;
    volatile int segment = 1;
    shim_jmp_armed = 1;
    if (setjmp(shim_jmp) == 0) {
      dopri_integrate(obj, y0 + j * n, times1, n_times1, NULL, 0, is_event, NULL, scratch,
                      scratch + n * nt1, ctl->return_initial != 0);
      if (obj->code == OK_COMPLETE && times2[0] == obj->t) {
        segment = 2;
        dopri_integrate(obj, y_new ? y_new + j * n : obj->y, times2, n_times2, tcrit2, n_tcrit2,
                        is_event, NULL, y_out + j * n * nt2, out ? out + j * n_out * nt2 : NULL,
                        ctl->return_initial != 0);
        rc = (int32_t) obj->code;
      } else {
        obj->n_eval = obj->n_step = obj->n_accept = obj->n_reject = 0;
      }
    } else if (segment == 2) {
      rc = DDE_ERR_NONFINITE_DERIV;
    } else {
      obj->n_eval = obj->n_step = obj->n_accept = obj->n_reject = 0;
    }
    shim_jmp_armed = 0;
    shim_release_transient();
    stats[4 * j + DDE_STAT_N_EVAL] = (int32_t) obj->n_eval;
    stats[4 * j + DDE_STAT_N_STEP] = (int32_t) obj->n_step;
    stats[4 * j + DDE_STAT_N_ACCEPT] = (int32_t) obj->n_accept;
    stats[4 * j + DDE_STAT_N_REJECT] = (int32_t) obj->n_reject;
    code[j] = rc;
    t_last[j] = obj->t;
    step_size[j] = obj->step_size_initial;
    dopri_data_free(obj);
  }
  free(is_event);
  free(scratch);
  return 0;
}

/* ---- difeq --------------------------------------------------------------- */

static int32_t run_one_difeq(difeq_data *obj, size_t n, const double *y0,
                             const size_t *steps, size_t n_steps, double *y_out,
                             double *out, int32_t return_initial,
                             double *y_final) {
  int32_t rc;
  shim_jmp_armed = 1;
  if (setjmp(shim_jmp) == 0) {
    difeq_run(obj, y0, steps, n_steps, y_out, out, return_initial != 0);
    rc = DDE_OK_COMPLETE;
    if (y_final) {
      memcpy(y_final, obj->y1, n * sizeof(double));
    }
  } else {
    /* src/difeq.c:86,91 are argument errors checked by the caller;
       what remains is the history miss, src/difeq.c:267-270 */
    rc = DDE_ERR_DIFEQ_HISTORY;
  }
  shim_jmp_armed = 0;
  shim_release_transient();
  return rc;
}

/* shared_ring = 0: fresh difeq_data per replicate (the documented contract,
 * R/difeq_replicate.R:1-4).  shared_ring = 1: ONE difeq_data for all
 * replicates, exactly what src/r_difeq.c:65,95-102 does (SURVEY.md finding
 * 5) -- only used to demonstrate the difference. */
int ref_difeq_batch_ex(const char *model, size_t n, size_t n_out, size_t n_rep,
                       const double *y0, const double *parms, size_t n_parms,
                       size_t parms_stride, const uint64_t *steps,
                       size_t n_steps, size_t n_history, int32_t return_initial,
                       double *y_out, double *out, int32_t *code,
                       double *y_final, int shared_ring) {
  model_entry m;
  (void) n_parms;
  if (find_model(model, &m) != 0 || !m.target) {
    return 1;
  }
  if (n_steps < 2 || steps[n_steps - 1] == steps[0]) {
    snprintf(shim_msg, sizeof shim_msg,
             "Initialisation failure: Beginning and end steps are the same");
    return 2;
  }
  for (size_t i = 0; i + 1 < n_steps; ++i) {
    if (steps[i + 1] <= steps[i]) {
      snprintf(shim_msg, sizeof shim_msg,
               "Initialisation failure: Steps not strictly increasing");
      return 2;
    }
  }
  const size_t nt = return_initial ? n_steps : n_steps - 1;
  size_t *st = (size_t *) calloc(n_steps, sizeof(size_t));
  for (size_t i = 0; i < n_steps; ++i) {
    st[i] = (size_t) steps[i];
  }
  double *scratch_out = NULL;
  if (n_out > 0 && !out) {
    scratch_out = (double *) calloc(n_out * nt + 1, sizeof(double));
  }
  const double t0 = now_seconds();
  difeq_data *shared = NULL;
  if (shared_ring) {
    shared = difeq_data_alloc(m.target, n, n_out, parms, n_history, false);
  }
  for (size_t j = 0; j < n_rep; ++j) {
    difeq_data *obj = shared;
    if (!shared_ring) {
      obj = difeq_data_alloc(m.target, n, n_out, parms + j * parms_stride,
                             n_history, false);
    } else {
      obj->data = parms + j * parms_stride;
    }
    const int32_t rc = run_one_difeq(
        obj, n, y0 + j * n, st, n_steps, y_out + j * n * nt,
        n_out > 0 ? (out ? out + j * n_out * nt : scratch_out) : NULL,
        return_initial, y_final ? y_final + j * n : NULL);
    if (code) {
      code[j] = rc;
    }
    if (!shared_ring) {
      difeq_data_free(obj);
    }
  }
  if (shared) {
    difeq_data_free(shared);
  }
  last_seconds = now_seconds() - t0;
  free(st);
  free(scratch_out);
  return 0;
}

int ref_difeq_batch(const char *model, size_t n, size_t n_out, size_t n_rep,
                    const double *y0, const double *parms, size_t n_parms,
                    size_t parms_stride, const uint64_t *steps, size_t n_steps,
                    size_t n_history, int32_t return_initial, double *y_out,
                    double *out, int32_t *code, double *y_final) {
  return ref_difeq_batch_ex(model, n, n_out, n_rep, y0, parms, n_parms,
                            parms_stride, steps, n_steps, n_history,
                            return_initial, y_out, out, code, y_final, 0);
}

/* One replicate, also returning the `history` attribute as r_difeq does
 * (ring_buffer_read from the tail, src/r_difeq.c:178-186). */
long ref_difeq_history(const char *model, size_t n, size_t n_out, const double *y0,
                       const double *parms, const uint64_t *steps, size_t n_steps,
                       size_t n_history, int32_t grow_history, int32_t return_initial,
                       double *y_out, double *out, int32_t *code, double *history,
                       size_t max_entries) {
  model_entry m;
  if (find_model(model, &m) != 0 || !m.target) {
    return -1;
  }
  const size_t nt = return_initial ? n_steps : n_steps - 1;
  size_t *st = (size_t *) calloc(n_steps, sizeof(size_t));
  for (size_t i = 0; i < n_steps; ++i) {
    st[i] = (size_t) steps[i];
  }
  double *scratch_out = (double *) calloc(n_out * nt + 1, sizeof(double));
  difeq_data *obj = difeq_data_alloc(m.target, n, n_out, parms, n_history, grow_history != 0);
  *code = run_one_difeq(obj, n, y0, st, n_steps, y_out,
                        n_out > 0 ? (out ? out : scratch_out) : NULL, return_initial, NULL);
  long used = 0;
  if (n_history > 0) {
    used = (long) ring_buffer_used(obj->history, false);
    if (history) {
      const size_t take = (size_t) used < max_entries ? (size_t) used : max_entries;
      ring_buffer_read(obj->history, history, take);
    }
  }
  difeq_data_free(obj);
  free(st);
  free(scratch_out);
  return used;
}

/* difeq_continue (src/r_difeq.c:114-165): steps1, then -- same difeq_data
 * object -- steps2 from y_new, or from obj->y1 when y_new is NULL.  Results of
 * the second segment; a replicate whose first segment failed gets
 * DDE_ERR_RESTART. */
int ref_difeq_batch_continue(const char *model, size_t n, size_t n_out, size_t n_rep,
                             const double *y0, const double *parms, size_t n_parms,
                             size_t parms_stride, const uint64_t *steps1, size_t n_steps1,
                             const uint64_t *steps2, size_t n_steps2, const double *y_new,
                             size_t n_history, int32_t return_initial, double *y_out,
                             double *out, int32_t *code, double *y_final) {
  model_entry m;
  (void) n_parms;
  if (find_model(model, &m) != 0 || !m.target) {
    return 1;
  }
  const size_t nt1 = return_initial ? n_steps1 : n_steps1 - 1;
  const size_t nt2 = return_initial ? n_steps2 : n_steps2 - 1;
  size_t *st1 = (size_t *) calloc(n_steps1, sizeof(size_t));
  size_t *st2 = (size_t *) calloc(n_steps2, sizeof(size_t));
  for (size_t i = 0; i < n_steps1; ++i) st1[i] = (size_t) steps1[i];
  for (size_t i = 0; i < n_steps2; ++i) st2[i] = (size_t) steps2[i];
  double *scratch_y = (double *) calloc(n * nt1 + 1, sizeof(double));
  double *scratch_o = (double *) calloc(n_out * (nt1 > nt2 ? nt1 : nt2) + 1, sizeof(double));
  memset(y_out, 0, sizeof(double) * n * nt2 * n_rep);
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt2 * n_rep);
  }
  for (size_t j = 0; j < n_rep; ++j) {
    difeq_data *obj = difeq_data_alloc(m.target, n, n_out, parms + j * parms_stride, n_history,
                                       false);
    int32_t rc = run_one_difeq(obj, n, y0 + j * n, st1, n_steps1, scratch_y,
                               n_out > 0 ? scratch_o : NULL, return_initial, NULL);
    if (rc == DDE_OK_COMPLETE && st2[0] == obj->step) {
      rc = run_one_difeq(obj, n, y_new ? y_new + j * n : obj->y1, st2, n_steps2,
                         y_out + j * n * nt2,
                         n_out > 0 ? (out ? out + j * n_out * nt2 : scratch_o) : NULL,
                         return_initial, y_final ? y_final + j * n : NULL);
    } else {
      rc = DDE_ERR_RESTART;
    }
    code[j] = rc;
    difeq_data_free(obj);
  }
  free(st1);
  free(st2);
  free(scratch_y);
  free(scratch_o);
  return 0;
}
