/* dopri_args.h -- kernel arguments of the dopri kernels: the per-batch
 * constants of dopri_data (/root/reference/src/dopri.h:55-161) plus array
 * bases.  Shared by the thread-per-trajectory kernels (dopri_kernels.cuh), the
 * block-per-trajectory kernel (dopri_coop_kernel.cuh) and the host layer. */
#ifndef DDE_DOPRI_ARGS_H
#define DDE_DOPRI_ARGS_H

namespace dde {

#ifndef DDE_DYN_MAXN
#define DDE_DYN_MAXN 16 /* largest n of a model compiled with runtime n */
#endif
#ifndef DDE_DYN_MAXP
#define DDE_DYN_MAXP 16
#endif
#ifndef DDE_DYN_MAXOUT
#define DDE_DYN_MAXOUT 8
#endif

/* Kernel arguments: the per-batch constants of dopri_data plus array bases. */
struct DopriArgs {
  /* problem */
  long long n_traj;
  int n, n_out, n_parms;
  long long parms_stride;
  const double *y0;    /* [n x n_traj] */
  const double *parms; /* [n_parms x n_traj] or one shared vector */
  const double *times; /* [n_times] */
  const double *tcrit; /* [n_tcrit] */
  int n_times, n_tcrit, tcrit_idx0;
  const int *event;    /* [n_tcrit] the model's event to apply at tcrit[i], -1: none; or NULL */
  /* controls (src/dopri.h:133-146) */
  double atol, rtol, step_size_min, step_size_max, step_size_initial;
  unsigned long long step_max_n, stiff_check;
  int n_history, step_size_min_allow, return_initial;
  double sign; /* src/dopri.c:173 */
  /* results */
  double *y_out, *out;
  int *stats, *code;
  double *t_last, *step_size;
  /* start-up values from dopri_init_kernel */
  double *init_k1; /* [n x n_traj] f(t0, y0) */
  double *init_h;  /* [n_traj] first step size */
  /* history */
  double *ring;          /* [n_traj][n_history][order*n + 2] */
  unsigned *ring_count;  /* [n_traj] entries committed (for export) */
  int win_staged;        /* lag-window coefficients in (dynamic) shared memory */
  /* restart (dopri_continue, src/r_dopri.c:193-262): the rings, ring_count,
     code, t_last and step_size of the previous run are inputs */
  int restart;
  double *y_final;       /* [n x n_traj] obj->y on return */
  /* work distribution: next trajectory index to hand out */
  unsigned long long *next_traj;
  /* ... and which trajectory the k-th one handed out is (NULL: the k-th) */
  const unsigned *order;
  /* streamed download (dde_dopri_batch_run_download): trajectories are handed
     out in index order, so the results of chunk c = [c << chunk_shift,
     (c + 1) << chunk_shift) are complete long before the kernel ends.  The
     thread that retires the last trajectory of a chunk raises chunk_flag[c]
     (pinned, mapped host memory) and the host starts that chunk's
     device-to-host copies while the kernel is still running.  NULL: off. */
  unsigned *chunk_done; /* [n_chunks] trajectories retired so far */
  int *chunk_flag;      /* [n_chunks] host-visible */
  int chunk_shift;
};

#if defined(__CUDACC__) || defined(DDE_HOST_EMU)
/* called by the one thread that has written a trajectory's last result */
__device__ __forceinline__ void chunk_retire(const DopriArgs &a, long long traj) {
  if (a.chunk_done == nullptr) {
    return;
  }
  __threadfence(); /* the results before the count */
  const long long c = traj >> a.chunk_shift;
  const long long first = c << a.chunk_shift;
  const long long end = first + (1LL << a.chunk_shift) < a.n_traj ? first + (1LL << a.chunk_shift)
                                                                 : a.n_traj;
  if (atomicAdd(&a.chunk_done[c], 1u) + 1u == (unsigned) (end - first)) {
    __threadfence_system();
    *(volatile int *) &a.chunk_flag[c] = 1;
  }
}
#endif

} /* namespace dde */

#endif /* DDE_DOPRI_ARGS_H */
