// The Levenberg-Marquardt step policy as ONE function that runs on the device (k_lm_decide, a single thread)
// once per trial evaluation: gain ratio, accept / reject, step-length measurement, damping update, retry along
// the same direction, termination (scipy check_termination, scipy/optimize/_lsq/common.py:705) and the trace.
// Replaces the accept / Delta-update part of scipy's trf_no_bounds loop (scipy/optimize/_lsq/trf.py:509-560)
// behind the reference call at /root/reference/src/bundle_adjustment.py:319.
//
// The same rules are restated in oracle/line_ba_oracle.py (lm_solve, damping_update, backtrack_alpha); this
// header is plain C++ so that tests/test_device_math_on_host.py can compile it with g++ and drive it with the
// oracle's scalars.  Products and sums that feed a decision are written with mul()/add() so that the device
// does not contract them into FMAs (the oracle rounds every operation).
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef LBA_TRACE_MAX
#define LBA_TRACE_MAX 512
#endif

namespace lba {

#if defined(__CUDA_ARCH__)
#define LBA_POL_HD __host__ __device__
__device__ __forceinline__ double pol_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double pol_add(double a, double b) { return __dadd_rn(a, b); }
#elif defined(__CUDACC__)
#define LBA_POL_HD __host__ __device__
__host__ __device__ inline double pol_mul(double a, double b) { return a * b; }
__host__ __device__ inline double pol_add(double a, double b) { return a + b; }
#else
#define LBA_POL_HD
inline double pol_mul(double a, double b) { return a * b; }
inline double pol_add(double a, double b) { return a + b; }
#endif

// policy constants (oracle/line_ba_oracle.py: RATIO_GOOD ... ALPHA_HI)
constexpr double kRatioGood = 0.5, kLengthGood = 0.6;
constexpr int kStreakMax = 3;
constexpr double kSigmaGain = 1.15, kSigmaMin = 0.1;
constexpr double kLengthCap = 4.0, kGrowthMax = 10.0;
constexpr double kAlphaLo = 0.1, kAlphaHi = 0.5;

// Device-resident state of one lba_lm_solve.  Every kernel of the loop reads what it needs from here (which of
// the two evaluation sets is current, the damping, the step length, whether this pass solves or only retries),
// so that nothing about an iteration has to come from the host.
struct LmCtl {
  // options
  double ftol, xtol, gtol, lambda_min, lambda_max, min_ratio;
  int max_iter, max_evals, pcg_max_iter, max_backtrack;
  // state
  double lam, nu, sigma, alpha;
  double cost;        // cost of the current point
  double g_norm;      // inf-norm of the gradient at the point the last solve started from
  int cur;            // index of the current evaluation set (the trial is evaluated into cur ^ 1)
  int need_solve;     // 1: the next pass solves the damped system; 0: it retries the last direction at `alpha`
  int done, status;   // status: 0 max_iter / max_evals, 1 gtol, 2 ftol, 3 xtol, 4 ftol and xtol, -5 non-finite start
  int it, evals;      // damped solves, evaluations (the initial one included)
  int streak, bt;
  int n_acc, pcg_total;
  int cg_short;       // host-driven mode: the enqueued PCG steps did not reach the tolerance, more are needed
  int trace_len;
  int scaling_mode;   // consumed by k_line_factor: 0 keep Dl, 1 running maximum, 2 first Jacobian
  int q_last;         // PCG iterations of the last solve
  unsigned long long pass;   // passes decided so far (sequence number of the host mirror)
};

struct LmTrace {
  int it[LBA_TRACE_MAX], pcg[LBA_TRACE_MAX], accepted[LBA_TRACE_MAX];
  double cost[LBA_TRACE_MAX], cost_new[LBA_TRACE_MAX], lam[LBA_TRACE_MAX], pred[LBA_TRACE_MAX];
  double step_norm[LBA_TRACE_MAX], g_norm[LBA_TRACE_MAX], alpha[LBA_TRACE_MAX];
};

// what one trial produced (sums over all observations / parameters, already all-reduced)
struct LmTrial {
  double cost_new;          // cost at x + alpha * delta
  double rjd, jd2;          // sum r (J delta), sum (J delta)^2 for the FULL step delta
  double d2;                // |delta|^2
  double x2;                // |x|^2
  double g_norm;            // gradient inf-norm at x
  int pcg_iters, pcg_done;  // state of the PCG solve
};

LBA_POL_HD inline bool pol_finite(double v) { return v - v == 0.0; }

// One decision.  Returns true when a trial was actually judged (false: nothing to do or PCG steps missing).
LBA_POL_HD inline bool lm_decide(LmCtl& L, const LmTrial& t, LmTrace* tr) {
  if (L.done) return false;
  if (L.need_solve && !t.pcg_done && t.pcg_iters < L.pcg_max_iter) { L.cg_short = 1; return false; }
  L.cg_short = 0;
  L.pass += 1;
  if (L.need_solve) {
    // the gradient norm of the point this solve started from arrives with the trial: below gtol the solve is
    // discarded (same decision as testing before the solve, scipy trf.py:470-472)
    L.g_norm = t.g_norm;
    if (t.g_norm < L.gtol) { L.status = 1; L.done = 1; return false; }
    L.it += 1;
    L.pcg_total += t.pcg_iters;
    L.q_last = t.pcg_iters;
  }
  const double alpha = L.alpha, cost = L.cost, cost_new = t.cost_new;
  // pred = -alpha r.Jd - 1/2 alpha^2 |Jd|^2
  const double pred = pol_add(-pol_mul(alpha, t.rjd), -pol_mul(pol_mul(pol_mul(0.5, alpha), alpha), t.jd2));
  const double step_norm = pol_mul(alpha, sqrt(t.d2));
  const double x_norm = sqrt(t.x2);
  const double actual = pol_add(cost, -cost_new);
  const bool finite = pol_finite(cost_new);
  const double ratio = (finite && pred > 0.0) ? actual / pred : -1.0;
  const bool accepted = finite && ratio > L.min_ratio;
  const bool ftol_ok = finite && (actual < pol_mul(L.ftol, cost)) && (ratio > 0.25);
  const bool xtol_ok = step_norm < pol_mul(L.xtol, pol_add(L.xtol, x_norm));
  // best length along the direction: parabola through F(x), slope r.Jd, F(x + alpha d)
  bool have_len = false;
  double a_opt = 0.0;
  if (finite) {
    const double c2 = pol_add(pol_add(cost_new, -cost), -pol_mul(t.rjd, alpha)) / pol_mul(alpha, alpha);
    if (c2 > 0.0) { have_len = true; a_opt = -t.rjd / pol_mul(2.0, c2); }
  }
  if (tr && L.trace_len < LBA_TRACE_MAX) {
    const int k = L.trace_len++;
    tr->it[k] = L.it - 1; tr->pcg[k] = L.q_last; tr->accepted[k] = accepted ? 1 : 0;
    tr->cost[k] = cost; tr->cost_new[k] = cost_new; tr->lam[k] = L.lam; tr->pred[k] = pred;
    tr->step_norm[k] = step_norm; tr->g_norm[k] = L.g_norm; tr->alpha[k] = alpha;
  }
  L.evals += 1;
  if (accepted) {
    L.n_acc += 1;
    L.cur ^= 1;
    L.cost = cost_new;
    double a = have_len ? a_opt : kLengthCap;
    if (a > kLengthCap) a = kLengthCap;
    double fac;
    if (ratio >= kRatioGood && a >= kLengthGood) {
      L.streak += 1;
      const int s = L.streak < kStreakMax ? L.streak : kStreakMax;
      fac = s == 1 ? 0.5 : (s == 2 ? 0.25 : 0.125);
    } else {
      L.streak = 0;
      fac = 1.0;
      if (a < 1.0) { fac = 1.0 / pol_mul(a, a); if (fac > kGrowthMax) fac = kGrowthMax; }
    }
    double sg = pol_mul(kSigmaGain, a);
    if (sg < kSigmaMin) sg = kSigmaMin;
    if (sg > 1.0) sg = 1.0;
    L.sigma = sg;
    double lam = pol_mul(L.lam, fac);
    if (lam < L.lambda_min) lam = L.lambda_min;
    if (lam > L.lambda_max) lam = L.lambda_max;
    L.lam = lam;
    L.nu = 2.0;
    L.need_solve = 1; L.alpha = L.sigma; L.bt = 0; L.scaling_mode = 1;
  } else if (L.bt < L.max_backtrack) {
    L.bt += 1;
    double a = have_len ? a_opt : pol_mul(kAlphaHi, alpha);
    const double lo = pol_mul(kAlphaLo, alpha), hi = pol_mul(kAlphaHi, alpha);
    if (a < lo) a = lo;
    if (a > hi) a = hi;
    L.alpha = a;
    L.need_solve = 0;
  } else {
    double lam = pol_mul(pol_mul(L.lam, kGrowthMax), L.nu);
    if (lam > L.lambda_max) lam = L.lambda_max;
    L.lam = lam;
    L.nu = pol_mul(L.nu, 2.0);
    L.streak = 0;
    L.need_solve = 1; L.alpha = L.sigma; L.bt = 0; L.scaling_mode = 0;
  }
  if (ftol_ok && xtol_ok) L.status = 4;
  else if (ftol_ok) L.status = 2;
  else if (xtol_ok) L.status = 3;
  if (L.status) L.done = 1;
  if (!L.done) {
    if (L.need_solve && L.it >= L.max_iter) L.done = 1;
    if (L.max_evals > 0 && L.evals >= L.max_evals) L.done = 1;
  }
  return true;
}

}  // namespace lba
