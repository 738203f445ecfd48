// rls_solver_state.cuh — device-resident scalar state of one solve ("lane") and the
// scalar recurrences of FISTA / POGM / OptISTA / CGNR / ADMM, executed by the last block
// of the fused epilogue kernels (or by a 1-thread kernel) so that a whole solve! runs
// without a host round trip.  Every recurrence is evaluated in Float32 with individually
// rounded operations, exactly as the reference evaluates them in rT.
#pragma once
#include "rls_common.cuh"

struct DevState {
  // ---- configuration (constant during a solve) ----
  int kind, iterations, restart, n_cap, iterations_cg, vary_rho, n_reg, proj_mask;
  float rho, theta0, sigma_fac, rel_tol, abs_tol, tol_inner;
  float lam[4];
  double lam64[4];
  int lam_is_f64[4];
  float rho0[4];     // ADMM solver.rho (pristine copy, ADMM.jl:163)
  // ---- state ----
  int iteration, done;
  float theta, theta_old, theta_n, alpha, beta, gamma, gamma_old, sigma;
  float norm_x0, rel_res_norm, res_norm, thr;
  double save[8];
  long long b_len;   // length(b) for ADMM's sigma_abs (ADMM.jl:212)
  // CGNR
  float2 cg_alpha, cg_beta, cg_zeta;
  double rr, pp;
  // ADMM
  float a_rho[4], a_rk[4], a_sk[4], a_eps_pri[4], a_eps_dua[4], a_delta[4], a_thr[4], a_uscale[4];
  float sigma_abs;
  float cgi_res, cgi_prev, cgi_tol, cgi_beta;
  float2 cgi_alpha;
  int cgi_k, cgi_gate, cgi_last, cgi_total;
  // SplitBregman (ADMM-like inner iteration + Bregman update of β_y): SplitBregman.jl:19-46
  int sb, iterations_inner, sb_iter_cnt, sb_outer_gate, sb_total;
};

enum {
  STEP_NONE = 0,
  STEP_INIT,           // t[0] = |x0|^2  -> norm_x0 and per-solver scalar init
  STEP_FISTA_GRAD,     // t[0] = |res|^2 (non-elementwise prox: tail follows)
  STEP_FISTA_POST,     // t[0] = |res|^2, t[1] = Re(res . (x - xold))
  STEP_FISTA_TAIL,     // t[1] only; |res|^2 taken from save[0]
  STEP_POGM_PRE,
  STEP_POGM_GRAD,      // t[0] = |res|^2 -> save
  STEP_POGM_POST,      // t[0] = |res|^2 (or save), t[1..3] = Re(w.x), Re(w.z), Re(w.res)
  STEP_OPTISTA_PRE,
  STEP_OPTISTA_GRAD,
  STEP_OPTISTA_POST,
  STEP_CGNR_ALPHA,     // t[0..1] = p.v
  STEP_CGNR_BETA,      // t[0] = |r|^2
  STEP_CGNR_POST,      // t[0] = |p|^2
  STEP_ADMM_INIT,
  STEP_ADMM_ITER_BEGIN,
  STEP_ADMM_CG_INIT,   // t[0] = |r|^2
  STEP_ADMM_CG_BETA,
  STEP_ADMM_CG_ALPHA,  // t[0..1] = u.c
  STEP_ADMM_CG_POST,   // t[0] = |r|^2
  STEP_ADMM_TERM_SAVE, // gradient trafo: dual-domain sums parked in save[] until the pixel pass
  STEP_ADMM_TERM,      // residual bookkeeping of one term (index in `arg`; arg >= 16: merge save[])
  STEP_ADMM_ITER_END,
  STEP_SB_FINISH       // SplitBregman: counters after the (gated) Bregman update  SplitBregman.jl:264-271
};

#ifdef __CUDACC__
__device__ __forceinline__ float theta_next(float th, float four_or_eight) {
  // (1 + sqrt(1 + c*th^2)) / 2   FISTA.jl:180, POGM.jl:191-193, OptISTA.jl:170-174
  return fdiv(fadd(1.f, fsqrt(fadd(1.f, fmul(four_or_eight, fmul(th, th))))), 2.f);
}

__device__ __forceinline__ float thr_from(const DevState* S, int i, float factor) {
  // factor * λ with λ's upstream type: Float32 -> rounded product; Float64 -> product in
  // double then convert(T, ·)  (Regularization.jl:31)
  if (S->lam_is_f64[i]) return (float)((double)factor * S->lam64[i]);
  return fmul(factor, S->lam[i]);
}

__device__ inline void admm_update_done(DevState* S) {
  bool conv = true;
  for (int i = 0; i < S->n_reg; ++i) {
    if (S->a_rk[i] >= fadd(S->sigma_abs, fmul(S->rel_tol, S->a_eps_pri[i]))) conv = false;
    if (S->a_sk[i] >= fadd(S->sigma_abs, fmul(S->rel_tol, S->a_eps_dua[i]))) conv = false;
  }
  S->done = (conv || S->iteration >= S->iterations) ? 1 : 0;
}

__device__ inline bool admm_converged(const DevState* S) {
  for (int i = 0; i < S->n_reg; ++i) {
    if (S->a_rk[i] >= fadd(S->sigma_abs, fmul(S->rel_tol, S->a_eps_pri[i]))) return false;
    if (S->a_sk[i] >= fadd(S->sigma_abs, fmul(S->rel_tol, S->a_eps_dua[i]))) return false;
  }
  return true;
}
// done(::SplitBregman) = converged || (iteration == 1 && iter_cnt > iterations)   SplitBregman.jl:289
__device__ inline void sb_update_done(DevState* S) {
  S->done = (admm_converged(S) || (S->iteration == 1 && S->sb_iter_cnt > S->iterations)) ? 1 : 0;
}

// The scalar part of every solver step.  `t` are the fixed-order reduction totals.
__device__ inline void scalar_step(DevState* S, int step, int arg, const double* t) {
  switch (step) {
    case STEP_INIT: {
      S->norm_x0 = (float)sqrt(t[0]);
      S->iteration = 0;
      S->rel_res_norm = __int_as_float(0x7f800000);
      S->res_norm = __int_as_float(0x7f800000);
      S->theta = S->theta0;
      S->theta_old = S->theta0;
      if (S->kind == RLS_POGM) {
        S->sigma = 1.f;                       // gamma is NOT reset by init! (POGM.jl:138-164)
      } else if (S->kind == RLS_OPTISTA) {
        float tn = S->theta0;                 // OptISTA.jl:145-149
        for (int i = 0; i < S->iterations - 1; ++i) tn = theta_next(tn, 4.f);
        S->theta_n = theta_next(tn, 8.f);
      } else if (S->kind == RLS_CGNR) {
        S->rr = t[0];
        S->pp = t[0];                          // p = r at init
        S->cg_alpha = S->cg_beta = S->cg_zeta = make_float2(0.f, 0.f);
        // converged(): norm(x0)/z0 <= relTol  (CGNR.jl:181-183); NaN when z0 == 0 -> false
        float rel = fdiv(S->norm_x0, S->norm_x0);
        S->rel_res_norm = rel;
        S->res_norm = S->norm_x0;
        S->done = ((rel <= S->rel_tol) || 0 >= S->n_cap) ? 1 : 0;
        break;
      }
      S->done = (S->iterations <= 0) ? 1 : 0;
      if (S->kind == RLS_FISTA) S->thr = thr_from(S, 0, S->rho);
      break;
    }
    case STEP_FISTA_GRAD:
      // λ is re-normalised by init! AFTER the state was set up (FISTA.jl:128, rls_solver_set_reg): form ρλ for the
      // non-elementwise prox (L21 / TV) from the current λ, as the elementwise path does inside its kernel
      S->thr = thr_from(S, 0, S->rho);
      // fallthrough
    case STEP_POGM_GRAD:
    case STEP_OPTISTA_GRAD:
      S->save[0] = t[0];
      break;
    case STEP_FISTA_TAIL:
    case STEP_FISTA_POST: {
      const double rr = (step == STEP_FISTA_TAIL) ? S->save[0] : t[0];
      S->res_norm = (float)sqrt(rr);
      S->rel_res_norm = fdiv(S->res_norm, S->norm_x0);                     // FISTA.jl:156
      if (S->restart && (float)t[1] > 0.f) S->theta = 1.f;                 // :171-176
      S->theta_old = S->theta;                                             // :179-180
      S->theta = theta_next(S->theta_old, 4.f);
      S->iteration += 1;
      S->done = (S->rel_res_norm < S->rel_tol || S->iteration >= S->iterations) ? 1 : 0;
      break;
    }
    case STEP_POGM_PRE: {                                                   // POGM.jl:189-202
      S->theta_old = S->theta;
      const bool last = (S->iteration == S->iterations - 1) && S->restart;
      S->theta = theta_next(S->theta_old, last ? 8.f : 4.f);
      S->alpha = fdiv(fsub(S->theta_old, 1.f), S->theta);
      S->beta = fdiv(fmul(S->sigma, S->theta_old), S->theta);
      S->gamma_old = S->gamma;
      if (S->restart) S->gamma = fmul(S->rho, fadd(fadd(1.f, S->alpha), S->beta));
      else S->gamma = fdiv(fmul(S->rho, fsub(fadd(fmul(2.f, S->theta_old), S->theta), 1.f)), S->theta);
      S->thr = thr_from(S, 0, S->gamma);                                    // :216
      break;
    }
    case STEP_POGM_POST: {
      const double rr = (arg == 1) ? S->save[0] : t[0];
      S->res_norm = (float)sqrt(rr);
      S->rel_res_norm = fdiv(S->res_norm, S->norm_x0);                      // :185
      if (S->restart) {                                                     // :222-230
        float v = fsub(fdiv(fsub((float)t[1], (float)t[2]), S->gamma), (float)t[3]);
        if (v < 0.f) { S->sigma = 1.f; S->theta = 1.f; }
        else S->sigma = fmul(S->sigma, S->sigma_fac);
      }
      S->iteration += 1;
      S->done = (S->rel_res_norm < S->rel_tol || S->iteration >= S->iterations) ? 1 : 0;
      break;
    }
    case STEP_OPTISTA_PRE: {                                                // OptISTA.jl:168-176
      const float th = S->theta, tn2 = fmul(S->theta_n, S->theta_n);
      S->gamma = fmul(fdiv(fmul(2.f, th), tn2), fadd(fsub(tn2, fmul(2.f, fmul(th, th))), th));
      S->theta_old = th;
      S->theta = theta_next(th, (S->iteration == S->iterations - 1) ? 8.f : 4.f);
      S->alpha = fdiv(fsub(S->theta_old, 1.f), S->theta);
      S->beta = fdiv(S->theta_old, S->theta);
      S->thr = thr_from(S, 0, fmul(S->rho, S->gamma));                      // :190
      break;
    }
    case STEP_OPTISTA_POST: {
      const double rr = (arg == 1) ? S->save[0] : t[0];
      S->res_norm = (float)sqrt(rr);
      S->rel_res_norm = fdiv(S->res_norm, S->norm_x0);                      // :185
      S->iteration += 1;
      S->done = (S->rel_res_norm < S->rel_tol || S->iteration >= S->iterations) ? 1 : 0;
      break;
    }
    case STEP_CGNR_ALPHA: {                                                 // CGNR.jl:153-161
      const float nr = (float)sqrt(S->rr);
      const float zeta = fmul(nr, nr);
      S->cg_zeta = make_float2(zeta, 0.f);
      const bool cplx = arg != 0;
      float2 normvl = make_float2((float)t[0], cplx ? (float)t[1] : 0.f);
      bool lam_pos = S->lam_is_f64[0] ? (S->lam64[0] > 0.0) : (S->lam[0] > 0.f);
      if (lam_pos) {
        const float np = (float)sqrt(S->pp);
        const float np2 = fmul(np, np);
        if (S->lam_is_f64[0]) {
          // promoted to Float64 upstream; evaluate in double, round once
          double dr = (double)normvl.x + S->lam64[0] * (double)np2, di = (double)normvl.y;
          double den = dr * dr + di * di;
          S->cg_alpha = make_float2((float)((double)zeta * dr / den), (float)(-(double)zeta * di / den));
        } else {
          float2 den = make_float2(fadd(normvl.x, fmul(S->lam[0], np2)), normvl.y);
          S->cg_alpha = cplx ? cdiv(S->cg_zeta, den) : make_float2(fdiv(zeta, den.x), 0.f);
        }
      } else {
        S->cg_alpha = cplx ? cdiv(S->cg_zeta, normvl) : make_float2(fdiv(zeta, normvl.x), 0.f);
      }
      break;
    }
    case STEP_CGNR_BETA: {                                                  // :171
      S->rr = t[0];
      const bool cplx = arg != 0;
      float2 num = make_float2((float)t[0], 0.f);
      S->cg_beta = cplx ? cdiv(num, S->cg_zeta) : make_float2(fdiv(num.x, S->cg_zeta.x), 0.f);
      break;
    }
    case STEP_CGNR_POST: {
      S->pp = t[0];
      S->iteration += 1;
      S->res_norm = (float)sqrt(S->rr);
      S->rel_res_norm = fdiv(S->res_norm, S->norm_x0);
      S->done = ((S->rel_res_norm <= S->rel_tol) || S->iteration >= S->n_cap) ? 1 : 0;  // :181-185
      break;
    }
    // ------------------------------ ADMM ------------------------------------------
    case STEP_ADMM_INIT: {                                                  // ADMM.jl:207-217
      const float inf = __int_as_float(0x7f800000);
      for (int i = 0; i < 4; ++i) {
        S->a_rk[i] = inf; S->a_sk[i] = inf; S->a_eps_pri[i] = 0.f; S->a_eps_dua[i] = 0.f; S->a_delta[i] = inf;
        S->a_rho[i] = S->rho0[i]; S->a_uscale[i] = 1.f;
      }
      S->sigma_abs = (float)(sqrt((double)S->b_len) * (double)S->abs_tol);
      S->iteration = 0;
      S->cgi_k = 0; S->cgi_last = 0; S->cgi_total = 0;
      S->sb_outer_gate = 1; S->sb_total = 0;
      if (S->sb) {                                                           // SplitBregman.jl:198-199
        S->iteration = 1; S->sb_iter_cnt = 1;
        sb_update_done(S);
      } else {
        admm_update_done(S);
      }
      S->cgi_gate = 1;
      break;
    }
    case STEP_ADMM_ITER_BEGIN: {
      for (int i = 0; i < S->n_reg; ++i) {                                   // λ/(2ρ)  ADMM.jl:261 ; λ/ρ  SplitBregman.jl:236
        const float two_rho = S->sb ? S->a_rho[i] : fmul(2.f, S->a_rho[i]);
        S->a_thr[i] = S->lam_is_f64[i] ? (float)(S->lam64[i] / (double)two_rho) : fdiv(S->lam[i], two_rho);
      }
      S->cgi_gate = S->done;
      break;
    }
    case STEP_ADMM_TERM_SAVE: {
      for (int k = 0; k < 8; ++k) S->save[k] = t[k];
      break;
    }
    case STEP_ADMM_CG_INIT: {                                               // IterativeSolvers cg_iterator!
      S->cgi_res = (float)sqrt(t[0]);
      S->cgi_tol = fmaxf(fmul(S->tol_inner, S->cgi_res), 0.f);
      S->cgi_prev = 1.f;
      S->cgi_k = 0;
      S->cgi_gate = (S->done || !(S->cgi_k < S->iterations_cg && S->cgi_res > S->cgi_tol)) ? 1 : 0;
      S->cgi_beta = fdiv(fmul(S->cgi_res, S->cgi_res), fmul(S->cgi_prev, S->cgi_prev));
      break;
    }
    case STEP_ADMM_CG_ALPHA: {                                              // α = res² / (u·c), via a*inv(z)
      const float r2 = fmul(S->cgi_res, S->cgi_res);
      const bool cplx = arg != 0;
      if (cplx) {
        const double c = (double)(float)t[0], d = (double)(float)t[1];   // dot() returns ComplexF32, inv() widens it
        const double mag = 1.0 / (c * c + d * d);
        const float ir = (float)(c * mag), ii = (float)(-d * mag);
        S->cgi_alpha = make_float2(fmul(r2, ir), fmul(r2, ii));
      } else {
        S->cgi_alpha = make_float2(fdiv(r2, (float)t[0]), 0.f);
      }
      break;
    }
    case STEP_ADMM_CG_POST: {
      S->cgi_prev = S->cgi_res;
      S->cgi_res = (float)sqrt(t[0]);
      S->cgi_k += 1;
      S->cgi_gate = (S->done || !(S->cgi_k < S->iterations_cg && S->cgi_res > S->cgi_tol)) ? 1 : 0;
      S->cgi_beta = fdiv(fmul(S->cgi_res, S->cgi_res), fmul(S->cgi_prev, S->cgi_prev));
      break;
    }
    case STEP_ADMM_TERM: {
      // t = {|dx|², |dz|², |du|², |Φ'dz|², |Φx|², |z|², |Φx - z|², |Φ'u|²}   ADMM.jl:282-309
      const int i = arg & 15;
      double tt[8];
      for (int k = 0; k < 8; ++k) tt[k] = t[k];
      if (arg >= 16) { tt[1] = S->save[1]; tt[2] = S->save[2]; tt[4] = S->save[4]; tt[5] = S->save[5]; tt[6] = S->save[6]; }
      t = tt;
      const float rho = S->a_rho[i];
      const float delta_old = S->a_delta[i];
      S->a_delta[i] = fadd(fadd((float)sqrt(t[0]), (float)sqrt(t[1])), (float)sqrt(t[2]));
      // ADMM: ρ‖Φ'(z-zᵒˡᵈ)‖, ρ‖Φ'u‖ (ADMM.jl:290,299); SplitBregman: ‖ρΦ'(z-zᵒˡᵈ)‖, ‖ρΦ'u‖ — ρ already inside the sums
      S->a_sk[i] = S->sb ? (float)sqrt(t[3]) : fmul(rho, (float)sqrt(t[3]));
      S->a_eps_pri[i] = fmaxf((float)sqrt(t[4]), (float)sqrt(t[5]));
      S->a_rk[i] = (float)sqrt(t[6]);
      S->a_eps_dua[i] = S->sb ? (float)sqrt(t[7]) : fmul(rho, (float)sqrt(t[7]));
      S->a_uscale[i] = 1.f;
      if (S->sb) break;
      // rᵏ/ɛᵖʳⁱ > 10sᵏ/ɛᵈᵘᵃ parses as (10*sᵏ)/ɛᵈᵘᵃ ; Δ/Δᵒˡᵈ > 0.9 compares against the Float64 literal
      const float rp = fdiv(S->a_rk[i], S->a_eps_pri[i]), sd = fdiv(S->a_sk[i], S->a_eps_dua[i]);
      const float rp10 = fdiv(fmul(10.f, S->a_rk[i]), S->a_eps_pri[i]), sd10 = fdiv(fmul(10.f, S->a_sk[i]), S->a_eps_dua[i]);
      if ((S->vary_rho == RLS_VARY_RHO_BALANCE && rp > sd10) ||
          (S->vary_rho == RLS_VARY_RHO_PNP && (double)fdiv(S->a_delta[i], delta_old) > 0.9)) {
        S->a_rho[i] = fmul(rho, 2.f);
        S->a_uscale[i] = 0.5f;
      } else if (S->vary_rho == RLS_VARY_RHO_BALANCE && sd > rp10) {
        S->a_rho[i] = fdiv(rho, 2.f);
        S->a_uscale[i] = 2.f;
      }
      break;
    }
    case STEP_ADMM_ITER_END: {
      S->cgi_last = S->cgi_k;
      S->cgi_total += S->cgi_k;
      S->cgi_gate = 1;
      if (S->sb) {
        // Bregman update next iff converged || iteration >= iterationsInner   SplitBregman.jl:257
        S->sb_outer_gate = (admm_converged(S) || S->iteration >= S->iterations_inner) ? 0 : 1;
        break;
      }
      S->iteration += 1;
      admm_update_done(S);
      break;
    }
    case STEP_SB_FINISH: {                                                   // SplitBregman.jl:264-271
      if (S->sb_outer_gate == 0) { S->sb_iter_cnt += 1; S->iteration = 0; }
      S->iteration += 1;
      S->sb_total += 1;
      S->sb_outer_gate = 1;
      sb_update_done(S);
      break;
    }
    default:
      break;
  }
}
#endif
