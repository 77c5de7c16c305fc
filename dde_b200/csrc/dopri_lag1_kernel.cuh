/* dopri_lag1_kernel.cuh -- dop853 for ONE delay equation with a declared lag, one
 * trajectory per thread: the stepping kernel of the Mackey-Glass batch
 * (BASELINE.json configs[1]).
 *
 * Same arithmetic, operation for operation, as dopri_tpt_kernel (dopri_kernels.cuh) and
 * therefore as the reference:
 *   adaptive loop            /root/reference/src/dopri.c:291-514
 *   dop853 step/error/dense  src/dopri_853.c:159-380
 *   controller               src/dopri.c:516-525
 *   stiffness test           src/dopri.c:668-693, src/dopri_853.c:382-394
 * What differs is the shape of the program, which is built around what the first
 * kernel's profile showed (profiles/r1g_*, profiles/r2b_*): twelve warps per SM cannot
 * hide a 300-instruction dependency chain per evaluation, and a hot path of ~40 KB of
 * code misses the instruction cache all the time.
 *
 *   * The sixteen right-hand-side evaluations of a step are one loop around ONE inlined
 *     copy of the model, followed by one warp-uniform switch that files the result and
 *     forms the next stage input (the stage counter is warp-uniform: lanes walk a step
 *     attempt in lock-step, each with its own t and h).  The hot loop is small enough
 *     for the instruction cache.  (Stage inputs as loops over (weight, stage) lists in
 *     constant memory were measured too, profiles/r2c: 12.9 M trajectories/s -- the
 *     address arithmetic and loop control around each two-instruction term cost 40 %
 *     more instructions.)
 *   * The stage values k2 .. k15 live in shared memory, [stage][thread] (conflict-free,
 *     immediate offsets off one base register); k1, y and the controller state stay in
 *     registers.  Fewer live registers than the unrolled kernel, no spills.
 *   * The model declares its lag (M::lags, dde_model_def.h), so the history look-ups of
 *     FOUR consecutive evaluations are answered together, ahead of them
 *     (lag_prefetch_group, lag_ctx.cuh): four independent look-up chains next to each
 *     other instead of one chain per evaluation in front of pow.  An evaluation's
 *     ylag_1 finds its answer in the memo if the model asks for exactly the predicted
 *     time, and takes the ordinary path otherwise.
 */
#ifndef DDE_DOPRI_LAG1_KERNEL_CUH
#define DDE_DOPRI_LAG1_KERNEL_CUH

#include "dopri_kernels.cuh"

namespace dde {

/* where a stage value lives: k1 and k16 in registers, the rest in shared memory */
enum { L1_K2 = 0, L1_K3, L1_K4, L1_K5, L1_K6, L1_K7, L1_K8, L1_K9, L1_K10, L1_K14, L1_K15, L1_NSLOT };
#define DDE_L1_K(slot) K[(size_t) (slot) * DDE_TPB]

#define DDE_L1_SMEM_BYTES (sizeof(double) * ::dde::L1_NSLOT * DDE_TPB)

template <class M>
__global__ void __launch_bounds__(DDE_TPB, DDE_TPB > 128 ? 1 : 3) dopri_lag1_kernel(const DopriArgs a) {
  static_assert(M::N == 1 && M::N_LAGS == 1 && M::NP >= 0, "one equation, one declared lag");
  constexpr int ORDER = 8;
  constexpr int PMAX = M::NP > 0 ? M::NP : 1;
  constexpr int OMAX = M::N_OUT > 0 ? M::N_OUT : (M::N_OUT == 0 ? 1 : DDE_DYN_MAXOUT);
  constexpr int CFG = 1 | (ORDER << 16) | (1 << 24) | DDE_CFG_MEMO;
  constexpr int NST_STEP = 11, NST = 16;
  const int n_out = M::N_OUT == 0 ? 0 : a.n_out;
  constexpr int hl = ORDER + 2;
  const int s = threadIdx.x;

  if (threadIdx.x == 0) {
    s_lag_u.sign = a.sign;
    s_lag_u.cap = a.n_history;
    s_lag_u.n = 1;
    s_lag_u.hl = hl;
    s_lag_u.order = ORDER;
    s_lag_u.staged = 1;
  }
  __syncthreads();

  /* the thread's column of the stage values, after the lag window's coefficients */
  double *const K = s_wc + (size_t) DDE_WIN * ORDER * DDE_TPB + s;

  /* method constants, src/dopri.c:67-78 and :142-153 */
  constexpr double step_factor_min = 0.333, step_factor_max = 6.0, step_factor_safe = 0.9;
  const double step_constant = 0.125 - 0.0 * 0.2;
  const double sign = a.sign;
  const double atol = a.atol, rtol = a.rtol;
  const int n_times = a.n_times;
  const int nt = a.return_initial ? n_times : n_times - 1;
  const double t_end = a.times[n_times - 1];

  /* ---- per-lane solver object (src/dopri.h:55-161) ---- */
  long long traj = -1;
  bool active = false, done = false;
  double p[PMAX];
  double y = 0.0, k1 = 0.0;
  double t = 0.0, h = 0.0, fac_old = 1e-4, h_save = 0.0;
  bool stop = false, last = false, reject = false;
  int times_idx = 1, tcrit_idx = 0;
  double t_out = 0.0;
  int n_eval = 0, n_step = 0, n_accept = 0, n_reject = 0;
  unsigned long long stiff_n_stiff = 0, stiff_n_nonstiff = 0;
  double *yo_next = nullptr;

#ifndef DDE_L1_SYNC
#define DDE_L1_SYNC 0
#endif
  bool exhausted = false; /* no trajectory left to take */
  for (;;) {
    if (DDE_L1_SYNC) {
      /* the block's warps start every step attempt together, so that they run the same
         code at about the same time and share what the instruction caches hold */
      if (__syncthreads_and(exhausted)) {
        break;
      }
    }
    if (!active && !exhausted) {
      /* persistent threads: take the next trajectory, or leave */
      traj = (long long) atomicAdd(a.next_traj, 1ULL);
      if (traj >= a.n_traj) {
        exhausted = true;
        if (!DDE_L1_SYNC) {
          break;
        }
      }
    }
    if (!active && !exhausted) {
      active = true;
      done = false;
#pragma unroll
      for (int i = 0; i < PMAX; ++i) {
        if (i < M::NP) {
          p[i] = a.parms[traj * a.parms_stride + i];
        }
      }
      y = a.y0[traj];
      /* dopri_data_reset, src/dopri.c:138-219 (time checks: once, on the host) */
      t = a.times[0];
      times_idx = 1;
      yo_next = a.y_out + ((size_t) traj * (size_t) nt + (a.return_initial ? 1u : 0u));
      t_out = n_times > 1 ? a.times[1] : 0.0;
      tcrit_idx = a.tcrit_idx0;
      n_step = n_accept = n_reject = 0;
      stiff_n_stiff = stiff_n_nonstiff = 0;
      s_ring[s] = a.ring + (size_t) traj * (size_t) a.n_history * (size_t) hl;
      s_y0_value[s] = y;
      const RingStart rs = ring_start(a, traj, s_ring[s], hl);
      s_t0[s] = rs.t0s;
      s_count[s] = rs.count;
      s_head[s] = rs.head;
      lag_window_reset(s);
      s_cfg[s] = CFG;
      s_memo_t[s] = na_real();
      s_code[s] = 0;
      s_error[s] = 0;
      fac_old = 1e-4;
      stop = last = reject = false;
      h_save = 0.0;
      /* k1 = f(t0, y0) and the first step size come from dopri_init_kernel */
      k1 = a.init_k1[traj];
      h = a.init_h[traj];
      n_eval = a.stats[4 * traj + 0];
      const int code0 = a.code[traj];
      if (code0 == -8 || code0 == -10) {
        s_code[s] = code0;
        done = true;
      } else if (code0 != 0) { /* a lag look-up failed during start-up */
        s_code[s] = code0;
        s_error[s] = 1;
      }
    }

    /* ---- one trip of the adaptive loop, src/dopri.c:343-510 ---- */
    bool run = active && !done;
    bool force_this_step = false;
    if (run) {
      if ((unsigned long long) n_step > a.step_max_n) {
        s_error[s] = 1;
        s_code[s] = -3; /* ERR_TOO_MANY_STEPS */
        done = true;
      } else if (fabs(h) <= a.step_size_min) {
        if (a.step_size_min_allow) {
          h = copysign(a.step_size_min, sign);
          force_this_step = true;
        } else {
          s_error[s] = 1;
          s_code[s] = -4; /* ERR_STEP_SIZE_TOO_SMALL */
          done = true;
        }
      } else if (fabs(h) <= fabs(t) * DBL_EPSILON) {
        s_error[s] = 1;
        s_code[s] = -5; /* ERR_STEP_SIZE_VANISHED */
        done = true;
      }
      run = !done;
    }
    if (run) {
      double t_stop = t_end;
      if (tcrit_idx < a.n_tcrit && sign * a.tcrit[tcrit_idx] < sign * t_end) {
        t_stop = a.tcrit[tcrit_idx];
      }
      if ((t + 1.01 * h - t_stop) * sign > 0.0) {
        h = t_stop - t;
        stop = true;
      }
      if ((t + 1.01 * h - t_end) * sign > 0.0) {
        h_save = h;
        h = t_end - t;
        last = true;
      }
      ++n_step;
      /* the window follows the declared lag: [lags(t), lags(t + h)] */
      double la[1], lb[1];
      M::lags(t, p, la);
      M::lags(t + h, p, lb);
      s_att_lo[s] = sign * t;
      lag_window_ensure_range<ORDER, true>(s, fmin(sign * la[0], sign * lb[0]),
                                           fmax(sign * la[0], sign * lb[0]));
    }

    /* ---- the evaluations of the step, in lock-step across the warp ---- */
    double yin = 0.0, tt = t, ysave = 0.0;
    double k4 = 0.0, k5 = 0.0; /* y1's slope and y1 after the attempt (also in their slots) */
    double h_new = 0.0;
    double zz[4];       /* the answers to four consecutive evaluations' look-ups */
    unsigned zok = 0;   /* which of them are answers */
    if (run) {          /* input of evaluation 0: src/dopri_853.c:178-181 */
      yin = y + h * DP8_A21 * k1;
      tt = stage_time<1>(0, t, h);
    }
    int my_stages = run ? NST_STEP : 0;
#pragma unroll 1
    for (int st = 0; st < NST; ++st) {
      if (DDE_L1_SYNC == 2) {
        __syncthreads();
      }
      if (st >= my_stages) {
        continue;
      }
      /* the look-ups of evaluations st .. st + 3, all at once (12 asks what 10 asked, so
         the last group is 13, 14, 15); evaluation st's answer is picked first */
      const int zi = st == 12 ? 2 : (st > 12 ? st - 13 : st & 3);
      double zq = zi == 0 ? zz[0] : (zi == 1 ? zz[1] : (zi == 2 ? zz[2] : zz[3]));
      bool zgood = (zok >> zi) & 1u;
      if ((st & 3) == 0) {
        double tl[4];
        switch (st) {
          case 0:
#pragma unroll
            for (int q = 0; q < 4; ++q) M::lags(stage_time<1>(q, t, h), p, &tl[q]);
            break;
          case 4:
#pragma unroll
            for (int q = 0; q < 4; ++q) M::lags(stage_time<1>(4 + q, t, h), p, &tl[q]);
            break;
          case 8:
#pragma unroll
            for (int q = 0; q < 4; ++q) M::lags(stage_time<1>(8 + q, t, h), p, &tl[q]);
            break;
          default:
#pragma unroll
            for (int q = 0; q < 4; ++q) M::lags(stage_time<1>(q < 3 ? 13 + q : 15, t, h), p, &tl[q]);
            break;
        }
        LagWin w;
        lag_win_read(s, w);
        const unsigned okg = lag_prefetch_group<ORDER, 4>(s, w, tl, zz);
        if (st != 12) {
          zq = zz[0];
          zgood = okg & 1u;
        }
        zok = okg;
      }

      /* -- dopri_eval, src/dopri.c:831-835: the one copy of the model -- */
      double kt;
      {
        double tl[1];
        M::lags(tt, p, tl);
        s_tt[s] = tt;
        s_memo_t[s] = zgood ? tl[0] : na_real();
        s_memo_v[s] = zq;
        __builtin_assume(s_cfg[s] == CFG);
        M::deriv((size_t) 1, tt, &yin, &kt, p);
      }
      ++n_eval;

      /* -- file the result and form the input of the next evaluation (one warp-uniform
            switch per evaluation), src/dopri_853.c:178-246,304,330-347 -- */
      bool attempt_done = false, step_done = false;
      switch (st) {
        case 0:
          DDE_L1_K(L1_K2) = kt;
          yin = y + h * (DP8_A31 * k1 + DP8_A32 * kt);
          tt = stage_time<1>(1, t, h);
          break;
        case 1:
          DDE_L1_K(L1_K3) = kt;
          yin = y + h * (DP8_A41 * k1 + DP8_A43 * kt);
          tt = stage_time<1>(2, t, h);
          break;
        case 2:
          DDE_L1_K(L1_K4) = kt;
          yin = y + h * (DP8_A51 * k1 + DP8_A53 * DDE_L1_K(L1_K3) + DP8_A54 * kt);
          tt = stage_time<1>(3, t, h);
          break;
        case 3:
          DDE_L1_K(L1_K5) = kt;
          yin = y + h * (DP8_A61 * k1 + DP8_A64 * DDE_L1_K(L1_K4) + DP8_A65 * kt);
          tt = stage_time<1>(4, t, h);
          break;
        case 4:
          DDE_L1_K(L1_K6) = kt;
          yin = y + h * (DP8_A71 * k1 + DP8_A74 * DDE_L1_K(L1_K4) + DP8_A75 * DDE_L1_K(L1_K5) +
                         DP8_A76 * kt);
          tt = stage_time<1>(5, t, h);
          break;
        case 5:
          DDE_L1_K(L1_K7) = kt;
          yin = y + h * (DP8_A81 * k1 + DP8_A84 * DDE_L1_K(L1_K4) + DP8_A85 * DDE_L1_K(L1_K5) +
                         DP8_A86 * DDE_L1_K(L1_K6) + DP8_A87 * kt);
          tt = stage_time<1>(6, t, h);
          break;
        case 6:
          DDE_L1_K(L1_K8) = kt;
          yin = y + h * (DP8_A91 * k1 + DP8_A94 * DDE_L1_K(L1_K4) + DP8_A95 * DDE_L1_K(L1_K5) +
                         DP8_A96 * DDE_L1_K(L1_K6) + DP8_A97 * DDE_L1_K(L1_K7) + DP8_A98 * kt);
          tt = stage_time<1>(7, t, h);
          break;
        case 7:
          DDE_L1_K(L1_K9) = kt;
          yin = y + h * (DP8_A101 * k1 + DP8_A104 * DDE_L1_K(L1_K4) + DP8_A105 * DDE_L1_K(L1_K5) +
                         DP8_A106 * DDE_L1_K(L1_K6) + DP8_A107 * DDE_L1_K(L1_K7) +
                         DP8_A108 * DDE_L1_K(L1_K8) + DP8_A109 * kt);
          tt = stage_time<1>(8, t, h);
          break;
        case 8:
          DDE_L1_K(L1_K10) = kt;
          yin = y + h * (DP8_A111 * k1 + DP8_A114 * DDE_L1_K(L1_K4) + DP8_A115 * DDE_L1_K(L1_K5) +
                         DP8_A116 * DDE_L1_K(L1_K6) + DP8_A117 * DDE_L1_K(L1_K7) +
                         DP8_A118 * DDE_L1_K(L1_K8) + DP8_A119 * DDE_L1_K(L1_K9) + DP8_A1110 * kt);
          tt = stage_time<1>(9, t, h);
          break;
        case 9: /* k2 is free again: stage 11's value goes there */
          DDE_L1_K(L1_K2) = kt;
          yin = y + h * (DP8_A121 * k1 + DP8_A124 * DDE_L1_K(L1_K4) + DP8_A125 * DDE_L1_K(L1_K5) +
                         DP8_A126 * DDE_L1_K(L1_K6) + DP8_A127 * DDE_L1_K(L1_K7) +
                         DP8_A128 * DDE_L1_K(L1_K8) + DP8_A129 * DDE_L1_K(L1_K9) +
                         DP8_A1210 * DDE_L1_K(L1_K10) + DP8_A1211 * kt);
          ysave = yin;
          tt = stage_time<1>(10, t, h);
          break;
        case 10: /* src/dopri_853.c:242-246 */
          DDE_L1_K(L1_K3) = kt;
          k4 = DP8_B1 * k1 + DP8_B6 * DDE_L1_K(L1_K6) + DP8_B7 * DDE_L1_K(L1_K7) +
               DP8_B8 * DDE_L1_K(L1_K8) + DP8_B9 * DDE_L1_K(L1_K9) + DP8_B10 * DDE_L1_K(L1_K10) +
               DP8_B11 * DDE_L1_K(L1_K2) + DP8_B12 * kt;
          k5 = y + h * k4;
          attempt_done = true;
          break;
        case 11: /* k4 = f(t_old, k5): only the stiffness test reads it */
          k4 = kt;
          yin = k5; /* evaluation 12, src/dopri_853.c:304 */
          tt = stage_time<1>(12, t, h);
          break;
        case 12: /* k4 = f(t + h, k5) */
          k4 = kt;
          DDE_L1_K(L1_K4) = kt;
          yin = y + h * (DP8_A141 * k1 + DP8_A147 * DDE_L1_K(L1_K7) + DP8_A148 * DDE_L1_K(L1_K8) +
                         DP8_A149 * DDE_L1_K(L1_K9) + DP8_A1410 * DDE_L1_K(L1_K10) +
                         DP8_A1411 * DDE_L1_K(L1_K2) + DP8_A1412 * DDE_L1_K(L1_K3) + DP8_A1413 * kt);
          tt = stage_time<1>(13, t, h);
          break;
        case 13:
          DDE_L1_K(L1_K14) = kt;
          yin = y + h * (DP8_A151 * k1 + DP8_A156 * DDE_L1_K(L1_K6) + DP8_A157 * DDE_L1_K(L1_K7) +
                         DP8_A158 * DDE_L1_K(L1_K8) + DP8_A1511 * DDE_L1_K(L1_K2) +
                         DP8_A1512 * DDE_L1_K(L1_K3) + DP8_A1513 * k4 + DP8_A1514 * kt);
          tt = stage_time<1>(14, t, h);
          break;
        case 14:
          DDE_L1_K(L1_K15) = kt;
          yin = y + h * (DP8_A161 * k1 + DP8_A166 * DDE_L1_K(L1_K6) + DP8_A167 * DDE_L1_K(L1_K7) +
                         DP8_A168 * DDE_L1_K(L1_K8) + DP8_A169 * DDE_L1_K(L1_K9) + DP8_A1613 * k4 +
                         DP8_A1614 * DDE_L1_K(L1_K14) + DP8_A1615 * kt);
          tt = stage_time<1>(15, t, h);
          break;
        default: /* 15: k16 = kt */
          step_done = true;
          break;
      }

      if (attempt_done) {
        if (s_error[s]) {
          done = true; /* a lag look-up failed during dopri_step, src/dopri.c:387-389 */
          my_stages = 0;
          continue;
        }
        /* dopri853_error, src/dopri_853.c:249-283: scaled by sign, not |h|, guarded by
           deno < 0; k3 is evaluation 10's value, kt */
        const double sk = atol + rtol * fmax(fabs(y), fabs(k5));
        double erri = k4 - DP8_BHH1 * k1 - DP8_BHH2 * DDE_L1_K(L1_K9) - DP8_BHH3 * kt;
        const double err2 = 0.0 + sq(erri / sk);
        erri = DP8_ER1 * k1 + DP8_ER6 * DDE_L1_K(L1_K6) + DP8_ER7 * DDE_L1_K(L1_K7) +
               DP8_ER8 * DDE_L1_K(L1_K8) + DP8_ER9 * DDE_L1_K(L1_K9) + DP8_ER10 * DDE_L1_K(L1_K10) +
               DP8_ER11 * DDE_L1_K(L1_K2) + DP8_ER12 * kt;
        double err = 0.0 + sq(erri / sk);
        double deno = err + 0.01 * err2;
        deno = deno < 0 ? 1.0 : deno;
        err = sign * err * sqrt(1.0 / ((double) 1 * deno));
        /* dopri_h_new, src/dopri.c:516-525 (step_beta = 0: pow(fac_old, 0) == 1) */
        const double fac11 = dde_pow_ctrl(err, step_constant);
        double fac = fac11 / 1.0;
        fac = fmax(1.0 / step_factor_max, fmin(1.0 / step_factor_min, fac / step_factor_safe));
        h_new = h / fac;
        if ((err <= 1) || force_this_step) {
          my_stages = NST;
          fac_old = fmax(err, 1e-4); /* src/dopri.c:398-399 */
          ++n_accept;
          /* input of evaluation 11: the redundant one at the OLD time, src/dopri.c:400-403
             (SURVEY.md A.1#2): it counts, and its look-up can fail */
          yin = k5;
          tt = stage_time<1>(11, t, h);
        } else {
          /* rejected, src/dopri.c:495-509 */
          h_new = h / fmin(1 / step_factor_min, fac11 / step_factor_safe);
          reject = true;
          if (n_accept >= 1) {
            ++n_reject;
          }
          last = false;
          stop = false;
          h = h_new;
        }
      }

      if (st == 11 && a.stiff_check != 0 &&
          (stiff_n_stiff > 0 || (unsigned long long) n_accept % a.stiff_check == 0)) {
        /* dopri_test_stiff, src/dopri.c:668-693, src/dopri_853.c:382-394 */
        const double stnum = 0.0 + sq(k4 - DDE_L1_K(L1_K3));
        const double stden = 0.0 + sq(k5 - ysave);
        const bool stiff = stden > 0 && fabs(h) * sqrt(stnum / stden) > 6.1;
        bool give_up = false;
        if (stiff) {
          stiff_n_nonstiff = 0;
          if (stiff_n_stiff++ >= 15) {
            give_up = true;
          }
        } else if (stiff_n_stiff > 0) {
          if (stiff_n_nonstiff++ >= 6) {
            stiff_n_stiff = 0;
          }
        }
        if (give_up) {
          s_error[s] = 1;
          s_code[s] = -7; /* ERR_STIFF */
          done = true;
          my_stages = 0;
          continue;
        }
      }

      if (step_done) {
        /* ---- the rest of an accepted step, src/dopri.c:410-494 ---- */
        /* dopri853_save_history, src/dopri_853.c:306-327,350-366 */
        double hd[ORDER];
        {
          const double ydiff = k5 - y;
          const double bspl = h * k1 - ydiff;
          hd[0] = y;
          hd[1] = ydiff;
          hd[2] = bspl;
          hd[3] = ydiff - h * k4 - bspl;
          const double k6 = DDE_L1_K(L1_K6), k7 = DDE_L1_K(L1_K7), k8 = DDE_L1_K(L1_K8),
                       k9 = DDE_L1_K(L1_K9), k10 = DDE_L1_K(L1_K10), k11 = DDE_L1_K(L1_K2),
                       k12 = DDE_L1_K(L1_K3), k14 = DDE_L1_K(L1_K14), k15 = DDE_L1_K(L1_K15);
          hd[4] = h * (DP8_D41 * k1 + DP8_D46 * k6 + DP8_D47 * k7 + DP8_D48 * k8 + DP8_D49 * k9 +
                       DP8_D410 * k10 + DP8_D411 * k11 + DP8_D412 * k12 + DP8_D413 * k4 +
                       DP8_D414 * k14 + DP8_D415 * k15 + DP8_D416 * kt);
          hd[5] = h * (DP8_D51 * k1 + DP8_D56 * k6 + DP8_D57 * k7 + DP8_D58 * k8 + DP8_D59 * k9 +
                       DP8_D510 * k10 + DP8_D511 * k11 + DP8_D512 * k12 + DP8_D513 * k4 +
                       DP8_D514 * k14 + DP8_D515 * k15 + DP8_D516 * kt);
          hd[6] = h * (DP8_D61 * k1 + DP8_D66 * k6 + DP8_D67 * k7 + DP8_D68 * k8 + DP8_D69 * k9 +
                       DP8_D610 * k10 + DP8_D611 * k11 + DP8_D612 * k12 + DP8_D613 * k4 +
                       DP8_D614 * k14 + DP8_D615 * k15 + DP8_D616 * kt);
          hd[7] = h * (DP8_D71 * k1 + DP8_D76 * k6 + DP8_D77 * k7 + DP8_D78 * k8 + DP8_D79 * k9 +
                       DP8_D710 * k10 + DP8_D711 * k11 + DP8_D712 * k12 + DP8_D713 * k4 +
                       DP8_D714 * k14 + DP8_D715 * k15 + DP8_D716 * kt);
        }
        k1 = k4; /* src/dopri.c:420-421 */
        y = k5;
        const double t_old = t;
        t += h;

        /* output rows, src/dopri.c:437-456: interpolate from the head */
        for (;;) {
          if (times_idx >= n_times) {
            break;
          }
          const double tq = t_out;
          if (!(last || sign * tq <= sign * t)) {
            break;
          }
          const double t_next = times_idx + 1 < n_times ? a.times[times_idx + 1] : 0.0;
          const double theta = (tq - t_old) / h;
          const double theta1 = 1 - theta;
          const double row = interp853(1, theta, theta1, hd);
          double *yo = yo_next;
          yo_next += 1;
          yo[0] = row;
          if (n_out > 0) {
            double o[OMAX];
            const size_t r = (size_t) traj * (size_t) nt +
                             (size_t) (times_idx - (a.return_initial ? 0 : 1));
            double *oo = a.out + r * (size_t) n_out;
            M::output((size_t) 1, tq, &row, (size_t) n_out, o, p);
#pragma unroll
            for (int i = 0; i < OMAX; ++i) {
              if (i < n_out) {
                oo[i] = o[i];
              }
            }
          }
          ++times_idx;
          t_out = t_next;
        }

        /* ring_buffer_head_advance, src/dopri.c:460 */
        {
          double *dst = s_ring[s] + (size_t) s_head[s] * (size_t) hl;
#pragma unroll
          for (int i = 0; i < ORDER; ++i) {
            dst[i] = hd[i];
          }
          dst[ORDER] = t_old;
          dst[ORDER + 1] = h;
          ring_advance<ORDER>(s, a.n_history, hd, ORDER, t_old, h);
        }

        if (last) {
          s_code[s] = 1; /* OK_COMPLETE */
          done = true;
          my_stages = 0;
          continue;
        }
        if (fabs(h_new) >= a.step_size_max) {
          h_new = copysign(a.step_size_max, sign);
        }
        if (reject) {
          h_new = copysign(fmin(fabs(h_new), fabs(h)), sign);
          reject = false;
        }
        if (stop) {
          /* critical times, src/dopri.c:476-491 */
          while (tcrit_idx < a.n_tcrit && sign * a.tcrit[tcrit_idx] <= sign * t) {
            if (M::N_EVENTS > 0 && a.event != nullptr) {
              const int ev = a.event[tcrit_idx];
              if (ev >= 0) {
                M::event(ev, (size_t) 1, t, &y, p);
              }
            }
            ++tcrit_idx;
          }
          stop = false;
        } else {
          h = h_new;
        }
      }
    }

    if (active && done) {
      /* retire: statistics as r_dopri_cleanup reads them, src/r_dopri.c:412-428 */
      a.stats[4 * traj + 0] = n_eval;
      a.stats[4 * traj + 1] = n_step;
      a.stats[4 * traj + 2] = n_accept;
      a.stats[4 * traj + 3] = n_reject;
      a.code[traj] = s_code[s];
      if (s_code[s] == 1) {
        a.step_size[traj] = h_save; /* obj->step_size_initial on return, src/dopri.c:463 */
      } else if (!a.restart) {
        a.step_size[traj] = a.step_size_initial;
      }
      if (s_code[s] != -10) { /* a trajectory that could not restart keeps its old state */
        a.t_last[traj] = t;
        if (a.y_final != nullptr) {
          a.y_final[traj] = y;
        }
      }
      a.ring_count[traj] = s_count[s];
      chunk_retire(a, traj);
      active = false;
    }
  }
}

} /* namespace dde */

#endif /* DDE_DOPRI_LAG1_KERNEL_CUH */
