/* dde_model_def.h -- the compile-time description of a model: the struct the
 * solver kernels are templated on.  Split from dde_model.cuh (registration and
 * launch, which need the CUDA runtime) so that tests/emu can instantiate the same
 * kernels around the same model structs with a host compiler.
 *
 * A model struct M provides
 *   M::N, M::NP, M::N_OUT, M::N_EVENTS   compile-time shape (0 / -1: run time)
 *   M::deriv, M::output, M::event        the reference's model ABI, inlined
 *                                        (/root/reference/src/dopri.h:49-53)
 *   M::N_LAGS, M::lags                   optional (an extension): the times at which
 *                                        deriv(t, ...) reads the history,
 *                                            void lags(double t, const void *data, double *tlag)
 *                                        (the `lags` argument of dde23-style solvers).
 *   M::Scoped, M::deriv_scoped           with M::lags: the model compiled inside a struct
 *                                        (dde_model_scoped.inc), so that the kernel for one
 *                                        delay equation (dopri_dde1_kernel.cuh) can make the
 *                                        declared look-up itself, one evaluation ahead, and
 *                                        hand the model the answer.  A look-up deriv makes
 *                                        without having declared it ends the trajectory with
 *                                        DDE_ERR_LAG_UNDECLARED (-11): a wrong declaration
 *                                        is reported, never integrated.
 */
#ifndef DDE_MODEL_DEF_H
#define DDE_MODEL_DEF_H

#define DDE_DYN_OR(v, dyn) ((v) > 0 ? 0 : (dyn))

/* Event functions (src/dopri.h:53: void event(size_t n, double t, double *y,
 * void *data), applied at critical times, src/dopri.c:476-482): list a model's
 * events once,
 *     DDE_EVENT_TABLE(growth_events, identity, double_variables, halve_variables)
 * and register the model with DDE_REGISTER_DOPRI_EVENTS(..., growth_events, 3);
 * dde_dopri_batch_set_events() then names them by position.  Up to 8. */
#define DDE_EVENT_CASE_(k, f) \
  case k:                     \
    f(n, t, y, data);         \
    break;
#define DDE_EVENT_CASES_1(a) DDE_EVENT_CASE_(0, a)
#define DDE_EVENT_CASES_2(a, b) DDE_EVENT_CASES_1(a) DDE_EVENT_CASE_(1, b)
#define DDE_EVENT_CASES_3(a, b, c) DDE_EVENT_CASES_2(a, b) DDE_EVENT_CASE_(2, c)
#define DDE_EVENT_CASES_4(a, b, c, d) DDE_EVENT_CASES_3(a, b, c) DDE_EVENT_CASE_(3, d)
#define DDE_EVENT_CASES_5(a, b, c, d, e) DDE_EVENT_CASES_4(a, b, c, d) DDE_EVENT_CASE_(4, e)
#define DDE_EVENT_CASES_6(a, b, c, d, e, f) DDE_EVENT_CASES_5(a, b, c, d, e) DDE_EVENT_CASE_(5, f)
#define DDE_EVENT_CASES_7(a, b, c, d, e, f, g) \
  DDE_EVENT_CASES_6(a, b, c, d, e, f) DDE_EVENT_CASE_(6, g)
#define DDE_EVENT_CASES_8(a, b, c, d, e, f, g, h) \
  DDE_EVENT_CASES_7(a, b, c, d, e, f, g) DDE_EVENT_CASE_(7, h)
#define DDE_EVENT_PICK_(_1, _2, _3, _4, _5, _6, _7, _8, NAME, ...) NAME
#define DDE_EVENT_TABLE(TABLE, ...)                                                            \
  static __device__ __forceinline__ void TABLE(int which, size_t n, double t, double *y,       \
                                               void *data) {                                   \
    switch (which) {                                                                           \
      DDE_EVENT_PICK_(__VA_ARGS__, DDE_EVENT_CASES_8, DDE_EVENT_CASES_7, DDE_EVENT_CASES_6,    \
                      DDE_EVENT_CASES_5, DDE_EVENT_CASES_4, DDE_EVENT_CASES_3,                 \
                      DDE_EVENT_CASES_2, DDE_EVENT_CASES_1)(__VA_ARGS__)                       \
      default:                                                                                 \
        break;                                                                                 \
    }                                                                                          \
  }
#define DDE_NO_EVENTS(which, n, t, y, data) ((void) 0)

#define DDE_NO_LAGS(t, data, tlag) ((void) 0)

/* a model without declared lags: never launched on the scoped kernel */
#define DDE_NO_SCOPE_TYPE(NAME) ::dde::Dde1Scope
#define DDE_NO_SCOPE_CALL(sc, DERIV) DERIV
/* ... with: struct dde_scoped_NAME (dde_model_scoped.inc) */
#define DDE_SCOPE_TYPE(NAME) dde_scoped_##NAME
#define DDE_SCOPE_CALL(sc, DERIV) (sc).DERIV

#define DDE_DEFINE_DOPRI_MODEL(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, EVENTS, NEVENTS)        \
  DDE_DEFINE_DOPRI_MODEL_FULL(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, EVENTS, NEVENTS,         \
                              DDE_NO_LAGS, 0, DDE_NO_SCOPE_TYPE, DDE_NO_SCOPE_CALL)

#define DDE_DEFINE_DOPRI_MODEL_LAGS(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, EVENTS, NEVENTS,   \
                                    LAGS, NLAGS)                                               \
  DDE_DEFINE_DOPRI_MODEL_FULL(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, EVENTS, NEVENTS, LAGS,   \
                              NLAGS, DDE_SCOPE_TYPE, DDE_SCOPE_CALL)

#define DDE_DEFINE_DOPRI_MODEL_FULL(NAME, DERIV, NEQ, NPARMS, OUTPUT, NOUT, EVENTS, NEVENTS,   \
                                    LAGS, NLAGS, SCOPE_TYPE, SCOPE_CALL)                       \
  struct dde_dopri_model_##NAME {                                                              \
    typedef SCOPE_TYPE(NAME) Scoped;                                                           \
    static __device__ __forceinline__ void deriv_scoped(Scoped &sc, size_t n, double t,        \
                                                        const double *y, double *dydt,         \
                                                        const void *data) {                    \
      SCOPE_CALL(sc, DERIV)(n, t, (double *) y, dydt, (void *) data);                          \
    }                                                                                          \
    static constexpr int N = (NEQ), NP = (NPARMS), N_OUT = (NOUT), N_EVENTS = (NEVENTS);       \
    static constexpr int N_LAGS = (NLAGS);                                                     \
    static __device__ __forceinline__ void lags(double t, const void *data, double *tlag) {    \
      LAGS(t, data, tlag);                                                                     \
    }                                                                                          \
    static __device__ __forceinline__ void deriv(size_t n, double t, const double *y,          \
                                                 double *dydt, const void *data) {             \
      /* the casts: a model written for the reference may take y and data without const  \
         (inst/examples/seir.c:5), which C lets through a function pointer */                \
      DERIV(n, t, (double *) y, dydt, (void *) data);                                          \
    }                                                                                          \
    static __device__ __forceinline__ void output(size_t n, double t, const double *y,         \
                                                  size_t n_out, double *out,                   \
                                                  const void *data) {                          \
      OUTPUT(n, t, y, n_out, out, data);                                                       \
    }                                                                                          \
    static __device__ __forceinline__ void event(int which, size_t n, double t, double *y,     \
                                                 void *data) {                                 \
      EVENTS(which, n, t, y, data);                                                            \
    }                                                                                          \
  };

#define DDE_NO_OUTPUT(n, t, y, n_out, out, data) ((void) 0)

#endif /* DDE_MODEL_DEF_H */
