/* Age-structured SEIR with two delays, n = 512 (BASELINE.json configs[4];
 * SURVEY.md section 8d, C5): 128 age classes x (S, E, I, R), state layout
 * S[0,128) E[128,256) I[256,384) R[384,512).  parms = (beta, sigma, delta).
 * Rates scaled from the reference's SEIR example
 * (/root/reference/inst/examples/seir.c:8-10).
 *
 *   lambda_a(t) = (beta / (N_a sum w)) sum_{|d| <= 4} w_|d| I_{a+d}(t)   banded contacts,
 *                                                       N_a = N / 128 per class
 *   new_inf_a   = S_a lambda_a
 *   lag_inf_a   = S_a(t - 14) lambda_a(t - 14) exp(-14 b)          latent period 14
 *   waning_a    = delta R_a + sigma I_a(t - 21) exp(-21 b) / 8     loss of immunity, lag 21
 *   S_a' = b N / 128 - b S_a - new_inf_a + waning_a
 *   E_a' = new_inf_a - lag_inf_a - b E_a
 *   I_a' = lag_inf_a - (b + sigma) I_a
 *   R_a' = sigma I_a - b R_a - waning_a
 *
 * The S and I blocks at t - 14 and the I block at t - 21 are fetched with
 * ylag_vec_int (src/dopri.c:816-828).  Written with the cooperative macros of
 * <dde/dde.h>: sequential C for the reference solver, one age class per
 * thread when a thread block integrates the trajectory. */
#include <dde/dde.h>

#define SEIR_AGE_CLASSES 128
#define SEIR_AGE_BAND 4

DDE_MODEL_FN void seir_age(size_t n, double t, const double *y, double *dydt,
                           const void *data) {
  const double *p = (const double *) data;
  const double beta = p[0], sigma = p[1], delta = p[2];
  const double b = 0.1, N = 1e7;
  const double surv14 = 0.24659696394160649;  /* exp(-1.4) */
  const double surv21 = 0.12245642825298191;  /* exp(-2.1) */
  const double w[SEIR_AGE_BAND + 1] = {1.0, 0.5, 0.25, 0.125, 0.0625};
  const int A = SEIR_AGE_CLASSES;

  DDE_SCRATCH_BEGIN();
  DDE_SCRATCH(int, idx14, 2 * SEIR_AGE_CLASSES);
  DDE_SCRATCH(int, idx21, SEIR_AGE_CLASSES);
  DDE_SCRATCH(double, lag14, 2 * SEIR_AGE_CLASSES);
  DDE_SCRATCH(double, lag21, SEIR_AGE_CLASSES);
  DDE_FOR(a, 0, A) {
    idx14[a] = a;             /* S block */
    idx14[A + a] = 2 * A + a; /* I block */
    idx21[a] = 2 * A + a;
  }
  DDE_SYNC();
  ylag_vec_int(t - 14.0, idx14, 2 * SEIR_AGE_CLASSES, lag14);
  ylag_vec_int(t - 21.0, idx21, SEIR_AGE_CLASSES, lag21);

  const double *S = y, *E = y + A, *I = y + 2 * A, *R = y + 3 * A;
  const double *S14 = lag14, *I14 = lag14 + A;
  const double births = b * N / SEIR_AGE_CLASSES;
  const double scale = N / SEIR_AGE_CLASSES * 2.875; /* class size x sum of the weights */
  /* transmission per unit of weighted infectives, once per evaluation -- not a division per
     class and term: most classes' force of infection is exactly zero for most of the run
     (the epidemic front moves four classes per latent period), and a GPU's IEEE division
     leaves its fast path for a zero numerator */
  const double bs = beta / scale;
  DDE_FOR(a, 0, A) {
    double force = 0.0, force14 = 0.0;
    for (int d = -SEIR_AGE_BAND; d <= SEIR_AGE_BAND; ++d) {
      const int c = a + d;
      if (c >= 0 && c < A) {
        const double wd = w[d < 0 ? -d : d];
        force += wd * I[c];
        force14 += wd * I14[c];
      }
    }
    const double new_inf = S[a] * (bs * force);
    const double lag_inf = S14[a] * (bs * force14) * surv14;
    const double waning = delta * R[a] + sigma * lag21[a] * surv21 * 0.125;
    dydt[a] = births - b * S[a] - new_inf + waning;
    dydt[A + a] = new_inf - lag_inf - b * E[a];
    dydt[2 * A + a] = lag_inf - (b + sigma) * I[a];
    dydt[3 * A + a] = sigma * I[a] - b * R[a] - waning;
  }
}
