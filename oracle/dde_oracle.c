/* TEST INFRASTRUCTURE ONLY -- not shipped, never on the product path.
 *
 * dde_oracle.c: plain-C restatement of the hot path of mrc-ide/dde, the
 * checker the CUDA kernels are compared with when the compiled reference
 * (oracle/_ref) is not at hand.  Only tests/, __graft_entry__.smoke() and
 * bench.py's CPU-baseline legs load it (through oracle/pyoracle.py).
 *
 * PARITY PINNED: tests/test_oracle_cpu.py checks this file bit for bit
 * against the fixtures in tests/golden/, which tools/gen_golden.py wrote from the
 * UNMODIFIED reference sources (oracle/_ref), and -- where oracle/_ref is
 * present -- against the reference itself on fresh random inputs.
 *
 * What it restates (file:line under /root/reference/):
 *   adaptive driver       src/dopri.c:291-514       -> integrate()
 *   reset / time checks   src/dopri.c:138-219       -> integrate() prologue
 *   first step size       src/dopri.c:527-572       -> first_step()
 *   controller            src/dopri.c:516-525       -> integrate()
 *   dopri5 stages/error   src/dopri_5.c:42-104      -> tableau DP5, attempt()
 *   dop853 stages/error   src/dopri_853.c:159-283   -> tableau DP8, attempt()
 *   dense coefficients    src/dopri_5.c:84-89,106-118; src/dopri_853.c:285-367
 *                                                   -> dense5(), dense853()
 *   dense evaluation      src/dopri.c:582-666, src/dopri_5.c:120-127,
 *                         src/dopri_853.c:369-380   -> dense_eval()
 *   history search, ylag  src/dopri.c:742-828       -> locate(), ylag_*()
 *   stiffness detector    src/dopri.c:668-693, src/dopri_5.c:129-140,
 *                         src/dopri_853.c:382-394   -> stiff_test()
 *   difeq loop, yprev     src/difeq.c:74-105,127-324 -> iterate(), yprev_*()
 *   ring buffer           third-party mrc-ide/ring (>= 1.0.6, DESCRIPTION:20),
 *                         not under /root/reference: semantics restated from
 *                         the call sites -> hist_*
 *   pow                   third-party glibc 2.39 libm (src/dopri.c:500,517,
 *                         521,569) -> dde_pow (include/dde_b200/dde_math.h), itself
 *                         pinned against libm by tests/test_math_cpu.py
 *
 * The structure is deliberately unlike the reference's hand-unrolled stage
 * code and unlike the CUDA kernel's switch: one table-driven explicit
 * Runge-Kutta stepper walks a list of (stage, coefficient) terms.  Rounding
 * is unchanged because each stage sum is formed left to right over the same
 * terms in the same order, and the file is compiled with -ffp-contract=off.
 *
 * Batch entry points take the argument lists of dde_dopri_batch /
 * dde_difeq_batch (include/dde_b200.h).
 */
#define _GNU_SOURCE
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "dde_b200.h"
#include "dde_math.h"
#include "tableau.h"

/* ---- tableaux ---------------------------------------------------------- */

typedef struct {
  int k;    /* which stored derivative (1-based name of the reference's k array) */
  double a; /* its weight */
} term;

typedef struct {
  double c;     /* node: the stage is evaluated at t + c h */
  int dst;      /* k slot that receives f(t + c h, y_stage) */
  int n_terms;
  term terms[9];
} stage;

/* dopri5_step, src/dopri_5.c:54-80.  Slot 7 is the reference's `ysti`. */
static const stage DP5_STAGES[6] = {
    {DP5_C2, 2, 1, {{1, DP5_A21}}},
    {DP5_C3, 3, 2, {{1, DP5_A31}, {2, DP5_A32}}},
    {DP5_C4, 4, 3, {{1, DP5_A41}, {2, DP5_A42}, {3, DP5_A43}}},
    {DP5_C5, 5, 4, {{1, DP5_A51}, {2, DP5_A52}, {3, DP5_A53}, {4, DP5_A54}}},
    {1.0, 6, 5, {{1, DP5_A61}, {2, DP5_A62}, {3, DP5_A63}, {4, DP5_A64}, {5, DP5_A65}}},
    {1.0, 2, 5, {{1, DP5_A71}, {3, DP5_A73}, {4, DP5_A74}, {5, DP5_A75}, {6, DP5_A76}}},
};
static const term DP5_ERR[6] = {{1, DP5_E1}, {3, DP5_E3}, {4, DP5_E4},
                                {5, DP5_E5}, {6, DP5_E6}, {2, DP5_E7}};
static const term DP5_DENSE[6] = {{1, DP5_D1}, {3, DP5_D3}, {4, DP5_D4},
                                  {5, DP5_D5}, {6, DP5_D6}, {2, DP5_D7}};

/* dopri853_step, src/dopri_853.c:178-240: stages 11 and 12 reuse slots 2, 3 */
static const stage DP8_STAGES[11] = {
    {DP8_C2, 2, 1, {{1, DP8_A21}}},
    {DP8_C3, 3, 2, {{1, DP8_A31}, {2, DP8_A32}}},
    {DP8_C4, 4, 2, {{1, DP8_A41}, {3, DP8_A43}}},
    {DP8_C5, 5, 3, {{1, DP8_A51}, {3, DP8_A53}, {4, DP8_A54}}},
    {DP8_C6, 6, 3, {{1, DP8_A61}, {4, DP8_A64}, {5, DP8_A65}}},
    {DP8_C7, 7, 4, {{1, DP8_A71}, {4, DP8_A74}, {5, DP8_A75}, {6, DP8_A76}}},
    {DP8_C8, 8, 5, {{1, DP8_A81}, {4, DP8_A84}, {5, DP8_A85}, {6, DP8_A86}, {7, DP8_A87}}},
    {DP8_C9, 9, 6,
     {{1, DP8_A91}, {4, DP8_A94}, {5, DP8_A95}, {6, DP8_A96}, {7, DP8_A97}, {8, DP8_A98}}},
    {DP8_C10, 10, 7,
     {{1, DP8_A101}, {4, DP8_A104}, {5, DP8_A105}, {6, DP8_A106}, {7, DP8_A107}, {8, DP8_A108},
      {9, DP8_A109}}},
    {DP8_C11, 2, 8,
     {{1, DP8_A111}, {4, DP8_A114}, {5, DP8_A115}, {6, DP8_A116}, {7, DP8_A117}, {8, DP8_A118},
      {9, DP8_A119}, {10, DP8_A1110}}},
    {1.0, 3, 9,
     {{1, DP8_A121}, {4, DP8_A124}, {5, DP8_A125}, {6, DP8_A126}, {7, DP8_A127}, {8, DP8_A128},
      {9, DP8_A129}, {10, DP8_A1210}, {2, DP8_A1211}}},
};
static const term DP8_B[8] = {{1, DP8_B1}, {6, DP8_B6},   {7, DP8_B7},  {8, DP8_B8},
                              {9, DP8_B9}, {10, DP8_B10}, {2, DP8_B11}, {3, DP8_B12}};
static const term DP8_ER[8] = {{1, DP8_ER1}, {6, DP8_ER6},   {7, DP8_ER7},  {8, DP8_ER8},
                               {9, DP8_ER9}, {10, DP8_ER10}, {2, DP8_ER11}, {3, DP8_ER12}};
/* the three extra stages of the dense output, src/dopri_853.c:330-347 */
static const stage DP8_EXTRA[3] = {
    {DP8_C14, 10, 8,
     {{1, DP8_A141}, {7, DP8_A147}, {8, DP8_A148}, {9, DP8_A149}, {10, DP8_A1410}, {2, DP8_A1411},
      {3, DP8_A1412}, {4, DP8_A1413}}},
    {DP8_C15, 2, 8,
     {{1, DP8_A151}, {6, DP8_A156}, {7, DP8_A157}, {8, DP8_A158}, {2, DP8_A1511}, {3, DP8_A1512},
      {4, DP8_A1513}, {10, DP8_A1514}}},
    {DP8_C16, 3, 8,
     {{1, DP8_A161}, {6, DP8_A166}, {7, DP8_A167}, {8, DP8_A168}, {9, DP8_A169}, {4, DP8_A1613},
      {10, DP8_A1614}, {2, DP8_A1615}}},
};
/* dense-output weights, first pass (src/dopri_853.c:315-326) and second pass
   (:350-363), rows = coefficients 4..7 */
static const term DP8_D1[4][8] = {
    {{1, DP8_D41}, {6, DP8_D46}, {7, DP8_D47}, {8, DP8_D48}, {9, DP8_D49}, {10, DP8_D410},
     {2, DP8_D411}, {3, DP8_D412}},
    {{1, DP8_D51}, {6, DP8_D56}, {7, DP8_D57}, {8, DP8_D58}, {9, DP8_D59}, {10, DP8_D510},
     {2, DP8_D511}, {3, DP8_D512}},
    {{1, DP8_D61}, {6, DP8_D66}, {7, DP8_D67}, {8, DP8_D68}, {9, DP8_D69}, {10, DP8_D610},
     {2, DP8_D611}, {3, DP8_D612}},
    {{1, DP8_D71}, {6, DP8_D76}, {7, DP8_D77}, {8, DP8_D78}, {9, DP8_D79}, {10, DP8_D710},
     {2, DP8_D711}, {3, DP8_D712}},
};
static const term DP8_D2[4][4] = {
    {{4, DP8_D413}, {10, DP8_D414}, {2, DP8_D415}, {3, DP8_D416}},
    {{4, DP8_D513}, {10, DP8_D514}, {2, DP8_D515}, {3, DP8_D516}},
    {{4, DP8_D613}, {10, DP8_D614}, {2, DP8_D615}, {3, DP8_D616}},
    {{4, DP8_D713}, {10, DP8_D714}, {2, DP8_D715}, {3, DP8_D716}},
};

/* ---- model ABI (src/dopri.h:49-53, src/difeq.h:19-21) ----------------------- */

typedef void deriv_fn(size_t n, double t, const double *y, double *dydt, const void *data);
typedef void output_fn(size_t n, double t, const double *y, size_t n_out, double *out,
                       const void *data);
typedef void target_fn(size_t n, size_t step, const double *y, double *y_next, size_t n_out,
                       double *out, const void *data);

/* the shared model sources call libm by name; give them the restated pow/exp
   exactly as the device build does (include/dde_b200/dde_model.cuh) */
#define pow(x, y) dde_pow((x), (y))
#define exp(x) dde_exp((x))
#include "../tests/models/difeq_fixtures.h"
#include "../dde_b200/models/sir_delay.h"
#include "../tests/models/growth.h"
#include "../dde_b200/models/lorenz.h"
#include "../dde_b200/models/mackey_glass.h"
#include "../dde_b200/models/rossler.h"
#include "../tests/models/seir.h"
#include "../dde_b200/models/seir_age.h"
#undef pow
#undef exp

typedef void event_fn(size_t n, double t, double *y, void *data); /* src/dopri.h:53 */

typedef struct {
  const char *name;
  deriv_fn *deriv;
  output_fn *output;
  target_fn *target;
  event_fn *const *events; /* in the order of the model's DDE_EVENT_TABLE */
  int n_events;
} model_entry;

static event_fn *const growth_event_table[5] = {identity, double_variables, halve_variables,
                                                double_parameters, halve_parameters};

/* events for the dopri calls that follow (see ref_driver.c) */
static int32_t event_index[64];
static size_t n_event_index = 0;
void oracle_set_events(const int32_t *index, size_t n) {
  n_event_index = n < 64 ? n : 64;
  for (size_t i = 0; i < n_event_index; ++i) {
    event_index[i] = index[i];
  }
}

static const model_entry MODELS[] = {
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

static char last_message[512];
static double last_seconds = -1.0;
const char *oracle_last_message(void) { return last_message; }
double oracle_last_seconds(void) { return last_seconds; }

static const model_entry *find_model(const char *name) {
  for (size_t i = 0; i < sizeof MODELS / sizeof MODELS[0]; ++i) {
    if (strcmp(MODELS[i].name, name) == 0) {
      return &MODELS[i];
    }
  }
  snprintf(last_message, sizeof last_message, "unknown model '%s'", name);
  return NULL;
}

static double now_seconds(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
}

static double na_real(void) {
  const uint64_t bits = 0x7FF00000000007A2ULL; /* R's NA_real_ */
  double x;
  memcpy(&x, &bits, sizeof x);
  return x;
}

/* ---- history: the ring subset dde relies on ---------- */

typedef struct {
  size_t cap;      /* entries that can be committed before the oldest is lost */
  size_t len;      /* doubles per entry */
  size_t first;    /* slot of the oldest committed entry */
  size_t used;     /* committed entries */
  int grow;        /* enlarge instead of losing the oldest entry (OVERFLOW_GROW) */
  double *slots;   /* cap + 1 slots: one spare so the head is always writable */
} hist;

static void hist_init(hist *h, size_t cap, size_t len) {
  h->cap = cap;
  h->len = len;
  h->first = 0;
  h->used = 0;
  h->grow = 0;
  h->slots = (double *) calloc((cap + 1) * len, sizeof(double));
}

static double *hist_slot(const hist *h, size_t k) { /* k-th oldest committed; k == used: head */
  return h->slots + ((h->first + k) % (h->cap + 1)) * h->len;
}

static double *hist_head(const hist *h) { return hist_slot(h, h->used); }

static void hist_commit(hist *h) { /* ring_buffer_head_advance */
  if (h->used == h->cap && h->grow && h->cap > 0) {
    /* grow policy: nothing was ever dropped, so first == 0 and the entries are
       contiguous; double the storage and carry on */
    const size_t bigger = 2 * h->cap;
    h->slots = (double *) realloc(h->slots, (bigger + 1) * h->len * sizeof(double));
    memset(h->slots + (h->cap + 1) * h->len, 0, (bigger - h->cap) * h->len * sizeof(double));
    h->cap = bigger;
  }
  if (h->used == h->cap) {
    if (h->cap > 0) {
      h->first = (h->first + 1) % (h->cap + 1);
    }
    /* cap == 0: the single spare slot is rewritten every step */
  } else {
    h->used++;
  }
}

/* ---- dopri ----------------------------------------------------------------- */

typedef struct {
  /* problem */
  size_t n, n_out;
  deriv_fn *deriv;
  output_fn *output;
  event_fn *const *events;
  int n_events;
  const void *data;
  int order; /* 5 or 8 */
  /* controls, src/dopri.h:133-146 */
  double atol, rtol, step_size_min, step_size_max, step_size_initial;
  size_t step_max_n, stiff_check;
  int step_size_min_allow;
  /* state */
  double sign, t0, t;
  double *y0, *y, *ynew, *ystage, *ysti;
  double *k[11]; /* k[1] .. k[10] */
  hist history;
  int error, code;
  size_t n_eval, n_step, n_accept, n_reject;
  size_t stiff_n_stiff, stiff_n_nonstiff;
} solver;

static solver *active_solver = NULL; /* the reference's dopri_global_obj, src/dopri.c:245 */

static void eval(solver *s, double t, const double *y, double *dydt) { /* src/dopri.c:831-835 */
  s->deriv(s->n, t, y, dydt, s->data);
  s->n_eval++;
}

static double weighted(const solver *s, const term *terms, int n_terms, size_t i) {
  double acc = terms[0].a * s->k[terms[0].k][i];
  for (int j = 1; j < n_terms; ++j) {
    acc = acc + terms[j].a * s->k[terms[j].k][i];
  }
  return acc;
}

static void run_stage(solver *s, const stage *st, int is_first, double h, double *ystage) {
  for (size_t i = 0; i < s->n; ++i) {
    if (is_first) {
      /* the first stage is written h * A21 * k1, i.e. (h A21) k1 */
      ystage[i] = s->y[i] + h * st->terms[0].a * s->k[1][i];
    } else {
      ystage[i] = s->y[i] + h * weighted(s, st->terms, st->n_terms, i);
    }
  }
  eval(s, st->c == 1.0 ? s->t + h : s->t + st->c * h, ystage, s->k[st->dst]);
}

/* One attempt at a step of size h; returns the error norm.  On return ynew
 * holds the proposed state, and (dop853) k[4] the weighted slope. */
static double attempt(solver *s, double h) {
  const size_t n = s->n;
  double *head = hist_head(&s->history);
  if (s->order == 5) {
    for (int j = 0; j < 6; ++j) {
      /* stage 5 is formed in ysti, kept for the stiffness test; stage 6 is y1 */
      run_stage(s, &DP5_STAGES[j], j == 0, h, j == 4 ? s->ysti : s->ynew);
      if (s->error) {
        /* the reference finishes the step before it looks at the flag
           (src/dopri.c:386-389): keep evaluating so n_eval agrees */
      }
    }
    double err = 0.0;
    for (size_t i = 0; i < n; ++i) {
      /* fifth dense coefficient, written on every attempt (src/dopri_5.c:84-89) */
      head[4 * n + i] = h * weighted(s, DP5_DENSE, 6, i);
    }
    for (size_t i = 0; i < n; ++i) {
      s->k[4][i] = h * weighted(s, DP5_ERR, 6, i); /* src/dopri_5.c:91-94 */
    }
    for (size_t i = 0; i < n; ++i) { /* dopri5_error, src/dopri_5.c:97-104 */
      const double sk = s->atol + s->rtol * fmax(fabs(s->y[i]), fabs(s->ynew[i]));
      const double q = s->k[4][i] / sk;
      err += q * q;
    }
    return sqrt(err / n);
  }
  for (int j = 0; j < 11; ++j) {
    run_stage(s, &DP8_STAGES[j], j == 0, h, s->ystage);
  }
  for (size_t i = 0; i < n; ++i) { /* src/dopri_853.c:242-246 */
    s->k[4][i] = weighted(s, DP8_B, 8, i);
    s->k[5][i] = s->y[i] + h * s->k[4][i];
  }
  /* dopri853_error, src/dopri_853.c:249-283: scaled by sign, guard deno < 0 */
  double err = 0.0, err2 = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double sk = s->atol + s->rtol * fmax(fabs(s->y[i]), fabs(s->k[5][i]));
    double e = s->k[4][i] - DP8_BHH1 * s->k[1][i] - DP8_BHH2 * s->k[9][i] - DP8_BHH3 * s->k[3][i];
    double q = e / sk;
    err2 += q * q;
    e = weighted(s, DP8_ER, 8, i);
    q = e / sk;
    err += q * q;
  }
  double deno = err + 0.01 * err2;
  if (deno < 0) {
    deno = 1.0;
  }
  return s->sign * err * sqrt(1.0 / (n * deno));
}

/* dopri5_save_history, src/dopri_5.c:106-118 */
static void dense5(solver *s, double h) {
  const size_t n = s->n;
  double *e = hist_head(&s->history);
  for (size_t i = 0; i < n; ++i) {
    const double ydiff = s->ynew[i] - s->y[i];
    const double bspl = h * s->k[1][i] - ydiff;
    e[i] = s->y[i];
    e[n + i] = ydiff;
    e[2 * n + i] = bspl;
    e[3 * n + i] = -h * s->k[2][i] + ydiff - bspl;
  }
  e[5 * n] = s->t;
  e[5 * n + 1] = h;
}

/* dopri853_save_history, src/dopri_853.c:285-367: four more evaluations */
static void dense853(solver *s, double h) {
  const size_t n = s->n;
  double *e = hist_head(&s->history);
  eval(s, s->t + h, s->k[5], s->k[4]);
  for (size_t i = 0; i < n; ++i) {
    const double ydiff = s->k[5][i] - s->y[i];
    const double bspl = h * s->k[1][i] - ydiff;
    e[i] = s->y[i];
    e[n + i] = ydiff;
    e[2 * n + i] = bspl;
    e[3 * n + i] = ydiff - h * s->k[4][i] - bspl;
    for (int r = 0; r < 4; ++r) {
      e[(4 + r) * n + i] = weighted(s, DP8_D1[r], 8, i);
    }
  }
  for (int j = 0; j < 3; ++j) {
    run_stage(s, &DP8_EXTRA[j], 0, h, s->ystage);
  }
  for (size_t i = 0; i < n; ++i) {
    for (int r = 0; r < 4; ++r) {
      double acc = e[(4 + r) * n + i];
      for (int j = 0; j < 4; ++j) {
        acc = acc + DP8_D2[r][j].a * s->k[DP8_D2[r][j].k][i];
      }
      e[(4 + r) * n + i] = h * acc;
    }
  }
  e[8 * n] = s->t;
  e[8 * n + 1] = h;
}

/* src/dopri_5.c:120-127, src/dopri_853.c:369-380; c = first coefficient of
   the component, coefficients n apart */
static double dense_poly(int order, size_t n, double theta, double theta1, const double *c) {
  if (order == 5) {
    return c[0] +
           theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * c[4 * n])));
  }
  const double tail = c[4 * n] + theta * (c[5 * n] + theta1 * (c[6 * n] + theta * c[7 * n]));
  return c[0] + theta * (c[n] + theta1 * (c[2 * n] + theta * (c[3 * n] + theta1 * tail)));
}

static double dense_eval(const solver *s, const double *entry, double t, size_t i) {
  const size_t it = (size_t) s->order * s->n;
  const double theta = (t - entry[it]) / entry[it + 1];
  const double theta1 = 1 - theta;
  return dense_poly(s->order, s->n, theta, theta1, entry + i);
}

/* The last committed entry whose start time is not after t (in the direction
 * of integration); NULL + ERR_YLAG_FAIL if there is none, src/dopri.c:742-769.
 * Only the result of the reference's search is defined;
 * this one walks back from the newest entry. */
static const double *locate(solver *s, double t) {
  const hist *h = &s->history;
  const size_t it = (size_t) s->order * s->n;
  for (size_t k = h->used; k-- > 0;) {
    const double *e = hist_slot(h, k);
    if (s->sign > 0 ? e[it] <= t : e[it] >= t) {
      return e;
    }
  }
  s->error = 1;
  s->code = DDE_ERR_YLAG_FAIL;
  return NULL;
}

double ylag_1(double t, size_t i) { /* src/dopri.c:776-788 */
  solver *s = active_solver;
  if (s->sign * t <= s->t0) {
    return s->y0[i];
  }
  const double *e = locate(s, t);
  return e ? dense_eval(s, e, t, i) : na_real();
}

void ylag_all(double t, double *y) { /* src/dopri.c:790-800 */
  solver *s = active_solver;
  if (s->sign * t <= s->t0) {
    memcpy(y, s->y0, s->n * sizeof(double));
    return;
  }
  const double *e = locate(s, t);
  if (e) {
    for (size_t i = 0; i < s->n; ++i) {
      y[i] = dense_eval(s, e, t, i);
    }
  }
}

void ylag_vec(double t, const size_t *idx, size_t nidx, double *y) { /* src/dopri.c:802-814 */
  solver *s = active_solver;
  if (s->sign * t <= s->t0) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = s->y0[idx[i]];
    }
    return;
  }
  const double *e = locate(s, t);
  if (e) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = dense_eval(s, e, t, idx[i]);
    }
  }
}

void ylag_vec_int(double t, const int *idx, size_t nidx, double *y) { /* src/dopri.c:816-828 */
  solver *s = active_solver;
  if (s->sign * t <= s->t0) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = s->y0[idx[i]];
    }
    return;
  }
  const double *e = locate(s, t);
  if (e) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = dense_eval(s, e, t, (size_t) idx[i]);
    }
  }
}

/* dopri_h_init, src/dopri.c:527-572 */
static double first_step(solver *s) {
  if (s->step_size_initial > 0.0) {
    return s->step_size_initial;
  }
  const size_t n = s->n;
  const double *f0 = s->k[1];
  double norm_f = 0.0, norm_y = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double sk = s->atol + s->rtol * fabs(s->y[i]);
    const double a = f0[i] / sk, b = s->y[i] / sk;
    norm_f += a * a;
    norm_y += b * b;
  }
  double h = (norm_f <= 1e-10 || norm_y <= 1e-10) ? 1e-6 : sqrt(norm_y / norm_f) * 0.01;
  h = copysign(fmin(h, s->step_size_max), s->sign);
  double *f1 = s->k[2], *y1 = s->k[3];
  for (size_t i = 0; i < n; ++i) {
    y1[i] = s->y[i] + h * f0[i];
  }
  eval(s, s->t + h, y1, f1);
  double der2 = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double sk = s->atol + s->rtol * fabs(s->y[i]);
    const double d = (f1[i] - f0[i]) / sk;
    der2 += d * d;
  }
  der2 = sqrt(der2) / h;
  const double der12 = fmax(fabs(der2), sqrt(norm_f));
  const double h1 = (der12 <= 1e-15) ? fmax(1e-6, fabs(h) * 1e-3)
                                     : dde_pow(0.01 / der12, 1.0 / s->order);
  h = fmin(fmin(100 * fabs(h), h1), s->step_size_max);
  return copysign(h, s->sign);
}

/* dopri_test_stiff, src/dopri.c:668-693 with the method parts of
   src/dopri_5.c:129-140 and src/dopri_853.c:382-394 */
static int stiff_test(solver *s, double h) {
  if (s->stiff_check == 0) {
    return 0; /* the reference turns 0 into SIZE_MAX, src/dopri.c:213-215 */
  }
  if (!(s->stiff_n_stiff > 0 || s->n_accept % s->stiff_check == 0)) {
    return 0;
  }
  double num = 0, den = 0;
  for (size_t i = 0; i < s->n; ++i) {
    double a, b;
    if (s->order == 5) {
      a = s->k[2][i] - s->k[6][i];
      b = s->ynew[i] - s->ysti[i];
    } else {
      a = s->k[4][i] - s->k[3][i];
      b = s->k[5][i] - s->ystage[i];
    }
    num += a * a;
    den += b * b;
  }
  const int stiff = den > 0 && fabs(h) * sqrt(num / den) > (s->order == 5 ? 3.25 : 6.1);
  int give_up = 0;
  if (stiff) {
    s->stiff_n_nonstiff = 0;
    if (s->stiff_n_stiff++ >= 15) {
      give_up = 1;
    }
  } else if (s->stiff_n_stiff > 0) {
    if (s->stiff_n_nonstiff++ >= 6) {
      s->stiff_n_stiff = 0;
    }
  }
  return give_up;
}

/* dopri_integrate (+ dopri_data_reset), src/dopri.c:138-219,291-514.
 * Returns the step size to restart with (obj->step_size_initial on exit). */
static double integrate(solver *s, const double *y_init, const double *times, size_t n_times,
                        const double *tcrit, size_t n_tcrit, double *y_out, double *out,
                        int return_initial) {
  const size_t n = s->n;
  const int m5 = s->order == 5;
  /* method constants, src/dopri.c:67-78,142-153 */
  const double fac_min = m5 ? 0.2 : 0.333, fac_max = m5 ? 10.0 : 6.0;
  const double beta = m5 ? 0.04 : 0.0, safe = 0.9;
  const double expo = m5 ? 0.2 - beta * 0.75 : 0.125 - beta * 0.2;
  double exit_step = s->step_size_initial;

  s->error = 0;
  s->code = DDE_NOT_SET;
  memcpy(s->y0, y_init, n * sizeof(double));
  if (s->y != y_init) { /* on a restart the caller may pass the solver's own state */
    memcpy(s->y, y_init, n * sizeof(double));
  }
  if (times[n_times - 1] == times[0]) {
    s->error = 1;
    s->code = DDE_ERR_ZERO_TIME_DIFFERENCE;
    return exit_step;
  }
  s->sign = copysign(1.0, times[n_times - 1] - times[0]);
  for (size_t i = 0; i + 1 < n_times; ++i) {
    if ((s->sign > 0 && times[i + 1] < times[i]) || (s->sign < 0 && times[i + 1] > times[i])) {
      s->error = 1;
      s->code = DDE_ERR_INCONSISTENT_TIME;
      return exit_step;
    }
  }
  /* an empty ring: a fresh start; otherwise a restart, whose pre-history ends
     at the oldest surviving entry (src/dopri.c:185-191) */
  s->t0 = s->history.used == 0
              ? s->sign * times[0]
              : s->sign * hist_slot(&s->history, 0)[(size_t) s->order * n];
  s->t = times[0];
  size_t i_time = 1, i_crit = 0;
  while (i_crit < n_tcrit && s->sign * tcrit[i_crit] <= s->sign * times[0]) {
    i_crit++;
  }
  s->n_eval = s->n_step = s->n_accept = s->n_reject = 0;
  s->stiff_n_stiff = s->stiff_n_nonstiff = 0;

  const double t_end = times[n_times - 1];
  double t_stop = t_end;
  if (i_crit < n_tcrit && s->sign * tcrit[i_crit] < s->sign * t_end) {
    t_stop = tcrit[i_crit];
  }
  active_solver = s;
  if (return_initial) {
    memcpy(y_out, y_init, n * sizeof(double));
    y_out += n;
    if (s->n_out > 0) {
      s->output(n, times[0], y_init, s->n_out, out, s->data);
      out += s->n_out;
    }
  }
  eval(s, s->t, s->y, s->k[1]);
  for (size_t i = 0; i < n; ++i) {
    if (!isfinite(s->k[1][i])) { /* Rf_error in the reference, src/dopri.c:331-336 */
      s->error = 1;
      s->code = DDE_ERR_NONFINITE_DERIV;
      active_solver = NULL;
      return exit_step;
    }
  }
  double h = first_step(s);
  double h_save = 0.0, fac_old = 1e-4;
  int stop = 0, last = 0, rejected = 0;

  for (;;) {
    int forced = 0;
    if (s->n_step > s->step_max_n) {
      s->error = 1;
      s->code = DDE_ERR_TOO_MANY_STEPS;
      break;
    }
    if (fabs(h) <= s->step_size_min) {
      if (!s->step_size_min_allow) {
        s->error = 1;
        s->code = DDE_ERR_STEP_SIZE_TOO_SMALL;
        break;
      }
      h = copysign(s->step_size_min, s->sign);
      forced = 1;
    } else if (fabs(h) <= fabs(s->t) * DBL_EPSILON) {
      s->error = 1;
      s->code = DDE_ERR_STEP_SIZE_VANISHED;
      break;
    }
    if ((s->t + 1.01 * h - t_stop) * s->sign > 0.0) {
      h = t_stop - s->t;
      stop = 1;
    }
    if ((s->t + 1.01 * h - t_end) * s->sign > 0.0) {
      h_save = h;
      h = t_end - s->t;
      last = 1;
    }
    s->n_step++;
    const double err = attempt(s, h);
    if (s->error) {
      break; /* a lag look-up failed during the step, src/dopri.c:387-389 */
    }
    /* dopri_h_new, src/dopri.c:516-525 */
    const double fac11 = dde_pow(err, expo);
    double fac = fac11 / dde_pow(fac_old, beta);
    fac = fmax(1.0 / fac_max, fmin(1.0 / fac_min, fac / safe));
    double h_new = h / fac;

    if (err <= 1 || forced) {
      fac_old = fmax(err, 1e-4);
      s->n_accept++;
      if (!m5) { /* the redundant evaluation at the old time, src/dopri.c:400-403 */
        eval(s, s->t, s->k[5], s->k[4]);
      }
      if (stiff_test(s, h)) {
        s->error = 1;
        s->code = DDE_ERR_STIFF;
        break;
      }
      if (m5) {
        dense5(s, h);
        memcpy(s->k[1], s->k[2], n * sizeof(double));
        memcpy(s->y, s->ynew, n * sizeof(double));
      } else {
        dense853(s, h);
        memcpy(s->k[1], s->k[4], n * sizeof(double));
        memcpy(s->y, s->k[5], n * sizeof(double));
      }
      s->t += h;
      /* requested times the step has passed -- all of them on the last step,
         src/dopri.c:437-456 -- from the entry still at the head */
      const double *e = hist_head(&s->history);
      while (i_time < n_times && (last || s->sign * times[i_time] <= s->sign * s->t)) {
        for (size_t i = 0; i < n; ++i) {
          y_out[i] = dense_eval(s, e, times[i_time], i);
        }
        if (s->n_out > 0) {
          s->output(n, times[i_time], y_out, s->n_out, out, s->data);
          out += s->n_out;
        }
        y_out += n;
        i_time++;
      }
      hist_commit(&s->history);
      if (last) {
        exit_step = h_save;
        s->code = DDE_OK_COMPLETE;
        break;
      }
      if (fabs(h_new) >= s->step_size_max) {
        h_new = copysign(s->step_size_max, s->sign);
      }
      if (rejected) {
        h_new = copysign(fmin(fabs(h_new), fabs(h)), s->sign);
        rejected = 0;
      }
      if (stop) { /* a critical time was reached: keep the clipped h, src/dopri.c:476-494 */
        while (i_crit < n_tcrit && s->sign * tcrit[i_crit] <= s->sign * s->t) {
          /* an event acts on the state (and maybe the parameters); k1 is left
             as it is, src/dopri.c:478-481 */
          if (n_event_index == n_tcrit && s->events && event_index[i_crit] >= 0 &&
              event_index[i_crit] < s->n_events) {
            s->events[event_index[i_crit]](n, s->t, s->y, (void *) s->data);
          }
          i_crit++;
        }
        t_stop = (i_crit < n_tcrit && s->sign * tcrit[i_crit] < s->sign * t_end) ? tcrit[i_crit]
                                                                                 : t_end;
        stop = 0;
      } else {
        h = h_new;
      }
    } else { /* src/dopri.c:495-509 */
      h = h / fmin(1 / fac_min, fac11 / safe);
      rejected = 1;
      if (s->n_accept >= 1) {
        s->n_reject++;
      }
      last = 0;
      stop = 0;
    }
  }
  active_solver = NULL;
  return exit_step;
}

static void solver_setup(solver *s, const model_entry *m, const dde_dopri_control *ctl, size_t n,
                         size_t n_out, const void *data) {
  memset(s, 0, sizeof *s);
  s->n = n;
  s->n_out = n_out;
  s->deriv = m->deriv;
  s->output = n_out > 0 ? m->output : NULL;
  s->events = m->events;
  s->n_events = m->n_events;
  s->data = data;
  s->order = ctl->method == DDE_DOPRI_853 ? 8 : 5;
  s->atol = ctl->atol;
  s->rtol = ctl->rtol;
  s->step_size_min = fmax(fabs(ctl->step_size_min), DBL_EPSILON); /* src/r_dopri.c:160-161 */
  s->step_size_max = fmin(fabs(ctl->step_size_max), DBL_MAX);
  s->step_size_initial = ctl->step_size_initial;
  s->step_max_n = (size_t) ctl->step_max_n;
  s->stiff_check = (size_t) ctl->stiff_check;
  s->step_size_min_allow = ctl->step_size_min_allow != 0;
  double *block = (double *) calloc((5 + 10) * n + 1, sizeof(double));
  s->y0 = block;
  s->y = block + n;
  s->ynew = block + 2 * n;
  s->ystage = block + 3 * n;
  s->ysti = block + 4 * n;
  for (int j = 1; j <= 10; ++j) {
    s->k[j] = block + (4 + j) * n;
  }
  hist_init(&s->history, (size_t) ctl->n_history, (size_t) s->order * n + 2);
  s->history.grow = ctl->grow_history != 0;
}

static void solver_release(solver *s) {
  free(s->y0);
  free(s->history.slots);
}

static int dopri_one(const model_entry *m, const dde_dopri_control *ctl, size_t n, size_t n_out,
                     const double *y0, const double *parms, const double *times, size_t n_times,
                     const double *tcrit, size_t n_tcrit, double *y_out, double *out,
                     int32_t *stats, int32_t *code, double *t_last, double *step_size,
                     double *history, size_t max_entries, long *n_entries) {
  solver s;
  solver_setup(&s, m, ctl, n, n_out, parms);
  const double exit_step = integrate(&s, y0, times, n_times, tcrit, n_tcrit, y_out, out,
                                     ctl->return_initial != 0);
  if (stats) {
    stats[DDE_STAT_N_EVAL] = (int32_t) s.n_eval;
    stats[DDE_STAT_N_STEP] = (int32_t) s.n_step;
    stats[DDE_STAT_N_ACCEPT] = (int32_t) s.n_accept;
    stats[DDE_STAT_N_REJECT] = (int32_t) s.n_reject;
  }
  if (code) {
    *code = s.code;
  }
  if (t_last) {
    *t_last = s.t;
  }
  if (step_size) {
    *step_size = exit_step;
  }
  if (n_entries) {
    *n_entries = (long) s.history.used;
    if (history) { /* oldest first, src/r_dopri.c:401-404 */
      for (size_t k = 0; k < s.history.used && k < max_entries; ++k) {
        memcpy(history + k * s.history.len, hist_slot(&s.history, k),
               s.history.len * sizeof(double));
      }
    }
  }
  solver_release(&s);
  return 0;
}

int oracle_dopri_batch(const char *model, const dde_dopri_control *ctl, size_t n, size_t n_out,
                       size_t n_traj, const double *y0, const double *parms, size_t n_parms,
                       size_t parms_stride, const double *times, size_t n_times,
                       const double *tcrit, size_t n_tcrit, double *y_out, double *out,
                       int32_t *stats, int32_t *code, double *t_last, double *step_size) {
  const model_entry *m = find_model(model);
  (void) n_parms;
  if (m == NULL || m->deriv == NULL) {
    return 1;
  }
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  if (y_out) { /* the R glue hands over zero-filled arrays, src/r_dopri.c:169 */
    memset(y_out, 0, sizeof(double) * n * nt * n_traj);
  }
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt * n_traj);
  }
  double *scratch = y_out ? NULL : (double *) calloc(n * nt + n_out * nt + 1, sizeof(double));
  const double t_start = now_seconds();
  for (size_t j = 0; j < n_traj; ++j) {
    dopri_one(m, ctl, n, n_out, y0 + j * n, parms + j * parms_stride, times, n_times, tcrit,
              n_tcrit, y_out ? y_out + j * n * nt : scratch,
              out ? out + j * n_out * nt : (scratch ? scratch + n * nt : NULL),
              stats ? stats + 4 * j : NULL, code ? code + j : NULL, t_last ? t_last + j : NULL,
              step_size ? step_size + j : NULL, NULL, 0, NULL);
  }
  last_seconds = now_seconds() - t_start;
  free(scratch);
  return 0;
}

/* dopri_continue, src/r_dopri.c:193-262: times1, then -- same solver object --
 * times2; the outputs are those of the second segment (see ref_driver.c). */
int oracle_dopri_batch_continue(const char *model, const dde_dopri_control *ctl, size_t n,
                                size_t n_out, size_t n_traj, const double *y0,
                                const double *parms, size_t n_parms, size_t parms_stride,
                                const double *times1, size_t n_times1, const double *times2,
                                size_t n_times2, const double *y_new, const double *tcrit2,
                                size_t n_tcrit2, double *y_out, double *out, int32_t *stats,
                                int32_t *code, double *t_last, double *step_size) {
  const model_entry *m = find_model(model);
  (void) n_parms;
  if (m == NULL || m->deriv == NULL) {
    return 1;
  }
  const size_t nt1 = ctl->return_initial ? n_times1 : n_times1 - 1;
  const size_t nt2 = ctl->return_initial ? n_times2 : n_times2 - 1;
  memset(y_out, 0, sizeof(double) * n * nt2 * n_traj);
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt2 * n_traj);
  }
  double *scratch = (double *) calloc(n * nt1 + n_out * nt1 + 1, sizeof(double));
  for (size_t j = 0; j < n_traj; ++j) {
    solver s;
    solver_setup(&s, m, ctl, n, n_out, parms + j * parms_stride);
    s.step_size_initial = integrate(&s, y0 + j * n, times1, n_times1, NULL, 0, scratch,
                                    scratch + n * nt1, ctl->return_initial != 0);
    if (s.code == DDE_OK_COMPLETE && times2[0] == s.t) {
      s.step_size_initial =
          integrate(&s, y_new ? y_new + j * n : s.y, times2, n_times2, tcrit2, n_tcrit2,
                    y_out + j * n * nt2, out ? out + j * n_out * nt2 : NULL,
                    ctl->return_initial != 0);
    } else {
      s.code = DDE_ERR_RESTART;
      s.n_eval = s.n_step = s.n_accept = s.n_reject = 0;
    }
    stats[4 * j + DDE_STAT_N_EVAL] = (int32_t) s.n_eval;
    stats[4 * j + DDE_STAT_N_STEP] = (int32_t) s.n_step;
    stats[4 * j + DDE_STAT_N_ACCEPT] = (int32_t) s.n_accept;
    stats[4 * j + DDE_STAT_N_REJECT] = (int32_t) s.n_reject;
    code[j] = s.code;
    t_last[j] = s.t;
    step_size[j] = s.step_size_initial;
    solver_release(&s);
  }
  free(scratch);
  return 0;
}

long oracle_dopri_history(const char *model, const dde_dopri_control *ctl, size_t n, size_t n_out,
                          const double *y0, const double *parms, const double *times,
                          size_t n_times, const double *tcrit, size_t n_tcrit, double *y_out,
                          double *out, int32_t *stats, int32_t *code, double *t_last,
                          double *step_size, double *history, size_t max_entries) {
  const model_entry *m = find_model(model);
  if (m == NULL || m->deriv == NULL) {
    return -1;
  }
  const size_t nt = ctl->return_initial ? n_times : n_times - 1;
  memset(y_out, 0, sizeof(double) * n * nt);
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt);
  }
  long used = 0;
  dopri_one(m, ctl, n, n_out, y0, parms, times, n_times, tcrit, n_tcrit, y_out, out, stats, code,
            t_last, step_size, history, max_entries, &used);
  return used;
}

/* ---- difeq ------------------------------------------------------------------ */

typedef struct {
  size_t n, n_out;
  const double *y0;
  long long step, step0;
  hist history; /* entries {step, y[n], out[n_out]}, src/difeq.c:31-37 */
  int have_history;
  int error;
} iterator;

static iterator *active_iterator = NULL; /* difeq_global_obj, src/difeq.c:120 */

/* src/difeq.c:261-272: the y block of the entry `step`, offset 0 = newest */
static const double *find_step(iterator *it, int step) {
  const long long offset = it->step - step;
  if (it->have_history && offset >= 0 && (size_t) offset < it->history.used) {
    return hist_slot(&it->history, it->history.used - 1 - (size_t) offset) + 1;
  }
  it->error = 1; /* Rf_error in the reference */
  return NULL;
}

double yprev_1(int step, size_t i) { /* src/difeq.c:274-283 */
  iterator *it = active_iterator;
  if (step <= (int) it->step0) {
    return it->y0[i];
  }
  const double *y = find_step(it, step);
  return y ? y[i] : na_real();
}

void yprev_all(int step, double *y) {
  iterator *it = active_iterator;
  if (step <= (int) it->step0) {
    memcpy(y, it->y0, it->n * sizeof(double));
    return;
  }
  const double *src = find_step(it, step);
  if (src) {
    memcpy(y, src, it->n * sizeof(double));
  }
}

void yprev_vec(int step, const size_t *idx, size_t nidx, double *y) {
  iterator *it = active_iterator;
  const double *src = step <= (int) it->step0 ? it->y0 : find_step(it, step);
  if (src) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = src[idx[i]];
    }
  }
}

void yprev_vec_int(int step, const int *idx, size_t nidx, double *y) {
  iterator *it = active_iterator;
  const double *src = step <= (int) it->step0 ? it->y0 : find_step(it, step);
  if (src) {
    for (size_t i = 0; i < nidx; ++i) {
      y[i] = src[idx[i]];
    }
  }
}

static void remember(iterator *it, const double *y) {
  if (!it->have_history) {
    return;
  }
  double *e = hist_head(&it->history);
  e[0] = (double) it->step; /* src/difeq.c:327-330 */
  memcpy(e + 1, y, it->n * sizeof(double));
  for (size_t i = 0; i < it->n_out; ++i) {
    e[1 + it->n + i] = na_real(); /* src/difeq.c:146 */
  }
  hist_commit(&it->history);
}

/* difeq_run, src/difeq.c:127-258, on an iterator that is fresh (restart == 0)
 * or carries the history of an earlier segment (difeq_continue). */
static int iterate_segment(iterator *it, target_fn *target, const double *y0, const void *data,
                           const uint64_t *steps, size_t n_steps, int return_initial,
                           int restart, double *y_out, double *out, double *y_final) {
  const size_t n = it->n, n_out = it->n_out;
  it->y0 = y0;
  it->error = 0;
  it->step = (long long) steps[0];
  /* difeq_data_reset, src/difeq.c:96-102: step0 is the oldest surviving entry's step */
  it->step0 = (it->have_history && it->history.used > 0)
                  ? (long long) hist_slot(&it->history, 0)[0]
                  : (long long) steps[0];
  double *y = (double *) calloc(2 * n + n_out + 1, sizeof(double));
  double *y_next = y + n, *o = y + 2 * n;
  memcpy(y, y0, n * sizeof(double));
  active_iterator = it;
  if (!(restart && it->have_history && it->history.used > 0)) {
    remember(it, y); /* the initial state is the first entry, src/difeq.c:140-156;
                        not committed again on a restart (first_entry is false) */
  }
  size_t i_step = 1;
  int pending_output = 0; /* the row just stored still lacks its outputs */
  if (return_initial) {
    memcpy(y_out, y, n * sizeof(double));
    y_out += n;
    pending_output = 1;
  }
  const long long last = (long long) steps[n_steps - 1];
  for (;;) {
    target(n, (size_t) it->step, y, y_next, n_out, o, data);
    if (it->error) {
      break;
    }
    it->step++;
    memcpy(y, y_next, n * sizeof(double));
    if (n_out > 0 && pending_output) { /* src/difeq.c:204-213 */
      memcpy(out, o, n_out * sizeof(double));
      out += n_out;
      pending_output = 0;
    }
    if (i_step < n_steps && it->step == (long long) steps[i_step]) {
      memcpy(y_out, y, n * sizeof(double));
      y_out += n;
      pending_output = 1;
      i_step++;
    }
    remember(it, y);
    if (it->step == last) {
      if (y_final) {
        memcpy(y_final, y, n * sizeof(double));
      }
      break;
    }
  }
  if (!it->error && n_out > 0 && pending_output) { /* src/difeq.c:249-255 */
    target(n, (size_t) it->step, y, y_next, n_out, out, data);
  }
  active_iterator = NULL;
  free(y);
  return it->error;
}

static void iterator_setup(iterator *it, size_t n, size_t n_out, size_t n_history) {
  memset(it, 0, sizeof *it);
  it->n = n;
  it->n_out = n_out;
  it->have_history = n_history > 0;
  if (it->have_history) {
    hist_init(&it->history, n_history, 1 + n + n_out);
  }
}

static void iterator_release(iterator *it) {
  if (it->have_history) {
    free(it->history.slots);
  }
}

/* one replicate with fresh state */
static int iterate(target_fn *target, size_t n, size_t n_out, const double *y0,
                   const void *data, const uint64_t *steps, size_t n_steps, size_t n_history,
                   int return_initial, double *y_out, double *out, double *y_final) {
  iterator it;
  iterator_setup(&it, n, n_out, n_history);
  const int failed = iterate_segment(&it, target, y0, data, steps, n_steps, return_initial, 0,
                                     y_out, out, y_final);
  iterator_release(&it);
  return failed;
}

/* One replicate, also returning the `history` attribute (src/r_difeq.c:178-186):
 * entries {step, y[n], out[n_out]} oldest first; the out columns are NA here (the
 * reference fills them through a pointer that lags the head, not restated). */
long oracle_difeq_history(const char *model, size_t n, size_t n_out, const double *y0,
                          const double *parms, const uint64_t *steps, size_t n_steps,
                          size_t n_history, int32_t grow_history, int32_t return_initial,
                          double *y_out, double *out, int32_t *code, double *history,
                          size_t max_entries) {
  const model_entry *m = find_model(model);
  if (m == NULL || m->target == NULL) {
    return -1;
  }
  const size_t nt = return_initial ? n_steps : n_steps - 1;
  double *scratch = (double *) calloc(n_out * nt + 1, sizeof(double));
  iterator it;
  iterator_setup(&it, n, n_out, n_history);
  if (it.have_history) {
    it.history.grow = grow_history != 0;
  }
  const int failed = iterate_segment(&it, m->target, y0, parms, steps, n_steps,
                                     return_initial != 0, 0, y_out, out ? out : scratch, NULL);
  *code = failed ? DDE_ERR_DIFEQ_HISTORY : DDE_OK_COMPLETE;
  long used = 0;
  if (it.have_history) {
    used = (long) it.history.used;
    for (size_t k = 0; history && k < it.history.used && k < max_entries; ++k) {
      memcpy(history + k * it.history.len, hist_slot(&it.history, k),
             it.history.len * sizeof(double));
    }
  }
  iterator_release(&it);
  free(scratch);
  return used;
}

/* difeq_continue, src/r_difeq.c:114-165: steps1 then steps2 on the same
 * iterator; results of the second segment (see ref_driver.c). */
int oracle_difeq_batch_continue(const char *model, size_t n, size_t n_out, size_t n_rep,
                                const double *y0, const double *parms, size_t n_parms,
                                size_t parms_stride, const uint64_t *steps1, size_t n_steps1,
                                const uint64_t *steps2, size_t n_steps2, const double *y_new,
                                size_t n_history, int32_t return_initial, double *y_out,
                                double *out, int32_t *code, double *y_final) {
  const model_entry *m = find_model(model);
  (void) n_parms;
  if (m == NULL || m->target == NULL) {
    return 1;
  }
  const size_t nt1 = return_initial ? n_steps1 : n_steps1 - 1;
  const size_t nt2 = return_initial ? n_steps2 : n_steps2 - 1;
  double *scratch_y = (double *) calloc(n * nt1 + n + 1, sizeof(double));
  double *scratch_o = (double *) calloc(n_out * (nt1 > nt2 ? nt1 : nt2) + 1, sizeof(double));
  double *y_end = scratch_y + n * nt1;
  memset(y_out, 0, sizeof(double) * n * nt2 * n_rep);
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt2 * n_rep);
  }
  for (size_t j = 0; j < n_rep; ++j) {
    iterator it;
    iterator_setup(&it, n, n_out, n_history);
    int failed = iterate_segment(&it, m->target, y0 + j * n, parms + j * parms_stride, steps1,
                                 n_steps1, return_initial != 0, 0, scratch_y, scratch_o, y_end);
    if (!failed && (long long) steps2[0] == it.step) {
      failed = iterate_segment(&it, m->target, y_new ? y_new + j * n : y_end,
                               parms + j * parms_stride, steps2, n_steps2, return_initial != 0, 1,
                               y_out + j * n * nt2, out ? out + j * n_out * nt2 : scratch_o,
                               y_final ? y_final + j * n : NULL);
      code[j] = failed ? DDE_ERR_DIFEQ_HISTORY : DDE_OK_COMPLETE;
    } else {
      code[j] = DDE_ERR_RESTART;
    }
    iterator_release(&it);
  }
  free(scratch_y);
  free(scratch_o);
  return 0;
}

int oracle_difeq_batch(const char *model, size_t n, size_t n_out, size_t n_rep, const double *y0,
                       const double *parms, size_t n_parms, size_t parms_stride,
                       const uint64_t *steps, size_t n_steps, size_t n_history,
                       int32_t return_initial, double *y_out, double *out, int32_t *code,
                       double *y_final) {
  const model_entry *m = find_model(model);
  (void) n_parms;
  if (m == NULL || m->target == NULL) {
    return 1;
  }
  if (n_steps < 2 || steps[n_steps - 1] == steps[0]) { /* src/difeq.c:85-87 */
    snprintf(last_message, sizeof last_message,
             "Initialisation failure: Beginning and end steps are the same");
    return 1;
  }
  for (size_t i = 0; i + 1 < n_steps; ++i) {
    if (steps[i + 1] <= steps[i]) { /* src/difeq.c:89-93 */
      snprintf(last_message, sizeof last_message,
               "Initialisation failure: Steps not strictly increasing");
      return 1;
    }
  }
  const size_t nt = return_initial ? n_steps : n_steps - 1;
  memset(y_out, 0, sizeof(double) * n * nt * n_rep);
  if (out) {
    memset(out, 0, sizeof(double) * n_out * nt * n_rep);
  }
  double *scratch = (double *) calloc(n_out * nt + 1, sizeof(double));
  const double t_start = now_seconds();
  for (size_t j = 0; j < n_rep; ++j) {
    const int failed = iterate(m->target, n, n_out, y0 + j * n, parms + j * parms_stride, steps,
                               n_steps, n_history, return_initial != 0, y_out + j * n * nt,
                               out ? out + j * n_out * nt : scratch,
                               y_final ? y_final + j * n : NULL);
    if (code) {
      code[j] = failed ? DDE_ERR_DIFEQ_HISTORY : DDE_OK_COMPLETE;
    }
  }
  last_seconds = now_seconds() - t_start;
  free(scratch);
  return 0;
}
