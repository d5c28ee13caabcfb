// optim.cu -- host-side optimiser loop on top of the evaluation shim.
//
// The reference drives its cost callbacks with GSL's Fletcher-Reeves conjugate gradient
// (gsl_multimin_fdfminimizer_conjugate_fr; src/frontend/local_optim_contrast_gsl.cpp:74-233,
// src/backend/global_optim_contrast_gsl.cpp:15-145).  GSL is a system library that is neither vendored nor
// present in this image, so its algorithm is RESTATED here from GSL 2.x multimin/conjugate_fr.c and
// multimin/directional_minimize.c (set / iterate / take_step / intermediate_point / minimize, including
// the "trial step already reduces f => accept it and double the step" shortcut and the restart every n
// iterations), together with the reference's own stopping rules (relative cost change < tolfun, |g| <
// epsabs_grad, <= 50 iterations).  PARITY UNPINNED against real GSL iterates; pinned against an independent
// restatement in oracle/gsl_fr.py driven by the CPU oracle cost (tests/test_gpu_optim.py).
// Pure host code: every cost evaluation goes through cmaxb_fe_eval / cmaxb_be_eval on the GPU.
#include <cmath>
#include <functional>
#include <vector>

#include "capi_common.cuh"

using namespace cmaxb;

namespace {

enum { kGslSuccess = 0, kGslContinue = -2, kGslEnoprog = 27 };

struct Cost {
  int n = 0;
  // f: value only; df: gradient only (the reference's df recomputes the value too); fdf: both.  cost = -contrast.
  std::function<int(const double*, double*)> f;
  std::function<int(const double*, double*, double*)> fdf;
  int f_evals = 0, g_evals = 0;      // value / gradient REQUESTS of the minimiser (what GSL would have called)
  int launches = 0;                  // cost evaluations actually run
  int rc = CMAXB_OK;
  // fused trials (cmaxb_opt_params.fused_trials): GSL evaluates every accepted line-search point twice -- f(x) for the
  // test, then df(x) at the SAME x (conjugate_fr.c / directional_minimize.c; the reference's df recomputes the value as
  // well, local_optim_contrast_gsl.cpp:65-70).  On the device a value + gradient evaluation costs less than a value-only
  // one plus a second value + gradient one, so every trial is evaluated with its gradient and remembered: the df request
  // at an already evaluated point is answered from the memo.  Same requests, same iterates; only the launches change.
  bool fuse = false;
  std::vector<double> memo_x, memo_g;
  double memo_f = 0;
  bool memo_ok = false;
  void run_fdf(const std::vector<double>& x, double* v, std::vector<double>& g) {
    if (rc == CMAXB_OK) rc = fdf(x.data(), v, g.data());
    ++launches;
    if (fuse) { memo_x = x; memo_g = g; memo_f = *v; memo_ok = rc == CMAXB_OK; }
  }
  double eval_f(const std::vector<double>& x) {
    double v = 0;
    ++f_evals;
    if (fuse) {
      if (memo_ok && memo_x == x) return memo_f;
      std::vector<double> g(x.size());
      run_fdf(x, &v, g);
      return v;
    }
    if (rc == CMAXB_OK) rc = f(x.data(), &v);
    ++launches;
    return v;
  }
  void eval_df(const std::vector<double>& x, std::vector<double>& g) {
    double v = 0;
    ++g_evals;
    if (fuse && memo_ok && memo_x == x) { g = memo_g; return; }
    run_fdf(x, &v, g);
  }
  void eval_fdf(const std::vector<double>& x, double* v, std::vector<double>& g) {
    ++f_evals; ++g_evals;
    if (fuse && memo_ok && memo_x == x) { *v = memo_f; g = memo_g; return; }
    run_fdf(x, v, g);
  }
};

double nrm2(const std::vector<double>& v) {   // gsl_blas_dnrm2 -> gslcblas cblas_dnrm2 (source_nrm2_r.h): scaled sum of squares
  double scale = 0.0, ssq = 1.0;
  for (double x : v) {
    if (x != 0.0) {
      const double ax = std::fabs(x);
      if (scale < ax) { ssq = 1.0 + ssq * (scale / ax) * (scale / ax); scale = ax; }
      else { ssq += (ax / scale) * (ax / scale); }
    }
  }
  return scale * std::sqrt(ssq);
}
double dot(const std::vector<double>& a, const std::vector<double>& b) {
  double s = 0;
  for (size_t i = 0; i < a.size(); ++i) s += a[i] * b[i];
  return s;
}

// directional_minimize.c: take_step
void take_step(const std::vector<double>& x, const std::vector<double>& p, double step, double lambda,
               std::vector<double>& x1, std::vector<double>& dx) {
  for (size_t i = 0; i < x.size(); ++i) { dx[i] = 0.0 + (-step * lambda) * p[i]; x1[i] = x[i] + 1.0 * dx[i]; }
}

// directional_minimize.c: intermediate_point
void intermediate_point(Cost& c, const std::vector<double>& x, const std::vector<double>& p, double lambda, double pg,
                        double stepa, double stepc, double fa, double fc, std::vector<double>& x1, std::vector<double>& dx,
                        std::vector<double>& gradient, double* step, double* f) {
  (void)stepa;
  double stepb, fb;
  for (;;) {
    const double u = std::fabs(pg * lambda * stepc);
    stepb = 0.5 * stepc * u / ((fc - fa) + u);
    take_step(x, p, stepb, lambda, x1, dx);
    if (x == x1) {   // trial point does not move from the initial point: fast exit
      *step = 0;
      *f = fa;
      c.eval_df(x1, gradient);
      return;
    }
    fb = c.eval_f(x1);
    if (fb >= fa && stepb > 0.0) {   // downhill step failed, reduce the step size and try again
      fc = fb;
      stepc = stepb;
      if (c.rc != CMAXB_OK) return;
      continue;
    }
    break;
  }
  *step = stepb;
  *f = fb;
  c.eval_df(x1, gradient);
}

// directional_minimize.c: minimize (Brent-like line minimisation, at most 10 trials)
void minimize(Cost& c, const std::vector<double>& x, const std::vector<double>& p, double lambda, double stepa, double stepb,
              double stepc, double fa, double fb, double fc, double tol, std::vector<double>& x1, std::vector<double>& dx1,
              std::vector<double>& x2, std::vector<double>& dx2, std::vector<double>& gradient, double* step, double* f,
              double* gnorm) {
  double u = stepb, v = stepa, w = stepc;
  double fu = fb, fv = fa, fw = fc;
  double old2 = std::fabs(w - v), old1 = std::fabs(v - u);
  double stepm, fm, pg, gnorm1;
  int iter = 0;
  x2 = x1;
  dx2 = dx1;
  *f = fb;
  *step = stepb;
  *gnorm = nrm2(gradient);
  for (;;) {
    iter++;
    if (iter > 10 || c.rc != CMAXB_OK) return;   // MAX ITERATIONS
    {
      const double dw = w - u, dv = v - u;
      double du = 0.0;
      const double e1 = ((fv - fu) * dw * dw + (fu - fw) * dv * dv);
      const double e2 = 2.0 * ((fv - fu) * dw + (fu - fw) * dv);
      if (e2 != 0.0) du = e1 / e2;
      if (du > 0.0 && du < (stepc - stepb) && std::fabs(du) < 0.5 * old2) stepm = u + du;
      else if (du < 0.0 && du > (stepa - stepb) && std::fabs(du) < 0.5 * old2) stepm = u + du;
      else if ((stepc - stepb) > (stepb - stepa)) stepm = 0.38 * (stepc - stepb) + stepb;
      else stepm = stepb - 0.38 * (stepb - stepa);
    }
    take_step(x, p, stepm, lambda, x1, dx1);
    fm = c.eval_f(x1);
    if (fm > fb) {
      if (fm < fv) { w = v; v = stepm; fw = fv; fv = fm; }
      else if (fm < fw) { w = stepm; fw = fm; }
      if (stepm < stepb) { stepa = stepm; fa = fm; }
      else { stepc = stepm; fc = fm; }
      continue;
    } else {   // fm <= fb
      old2 = old1;
      old1 = std::fabs(u - stepm);
      w = v; v = u; u = stepm;
      fw = fv; fv = fu; fu = fm;
      x2 = x1;
      dx2 = dx1;
      c.eval_df(x1, gradient);
      pg = dot(p, gradient);
      gnorm1 = nrm2(gradient);
      *f = fm;
      *step = stepm;
      *gnorm = gnorm1;
      if (std::fabs(pg * lambda / gnorm1) < tol) return;   // SUCCESS
      if (stepm < stepb) { stepc = stepb; fc = fb; stepb = stepm; fb = fm; }
      else { stepa = stepb; fa = fb; stepb = stepm; fb = fm; }
      continue;
    }
  }
}

// conjugate_fr.c state + set + iterate
struct ConjugateFr {
  int iter = 0;
  double step = 0, max_step = 0, tol = 0, pnorm = 0, g0norm = 0;
  std::vector<double> x1, dx1, x2, p, g0;
  std::vector<double> x, gradient, dx;
  double f = 0;

  void set(Cost& c, const std::vector<double>& x0, double step_size, double tol_) {
    const size_t n = x0.size();
    x1.assign(n, 0); dx1.assign(n, 0); x2.assign(n, 0); p.assign(n, 0); g0.assign(n, 0); dx.assign(n, 0); gradient.assign(n, 0);
    x = x0;
    iter = 0; step = step_size; max_step = step_size; tol = tol_;
    c.eval_fdf(x, &f, gradient);
    p = gradient;        // the gradient is the initial direction
    g0 = gradient;
    const double gnorm = nrm2(gradient);
    pnorm = gnorm; g0norm = gnorm;
  }

  int iterate(Cost& c) {
    const double fa = f;
    double fb, fc, dir, g1norm, pg;
    const double stepa = 0.0;
    double stepb;
    const double stepc = step;
    if (pnorm == 0.0 || g0norm == 0.0) { dx.assign(dx.size(), 0.0); return kGslEnoprog; }
    pg = dot(p, gradient);                       // which direction is downhill, +p or -p
    dir = (pg >= 0.0) ? +1.0 : -1.0;
    take_step(x, p, stepc, dir / pnorm, x1, dx);  // trial point x_c = x - step * p
    fc = c.eval_f(x1);
    if (c.rc != CMAXB_OK) return kGslEnoprog;
    if (fc < fa) {                                // success: reduced the function value
      step = stepc * 2.0;
      f = fc;
      x = x1;
      c.eval_df(x1, gradient);
      return kGslSuccess;
    }
    intermediate_point(c, x, p, dir / pnorm, pg, stepa, stepc, fa, fc, x1, dx1, gradient, &stepb, &fb);
    if (stepb == 0.0 || c.rc != CMAXB_OK) return kGslEnoprog;
    minimize(c, x, p, dir / pnorm, stepa, stepb, stepc, fa, fb, fc, tol, x1, dx1, x2, dx, gradient, &step, &f, &g1norm);
    x = x2;
    iter = (iter + 1) % (int)x.size();           // new conjugate direction
    if (iter == 0) {
      p = gradient;
      pnorm = g1norm;
    } else {
      const double beta = -std::pow(g1norm / g0norm, 2.0);   // p' = g1 - beta * p
      for (size_t i = 0; i < p.size(); ++i) p[i] = (-beta) * p[i];
      for (size_t i = 0; i < p.size(); ++i) p[i] += 1.0 * gradient[i];
      pnorm = nrm2(p);
    }
    g0norm = g1norm;
    g0 = gradient;
    return kGslSuccess;
  }
};

// The reference's loop around gsl_multimin_fdfminimizer_iterate (identical in FE and BE up to the constants).
int run_reference_loop(Cost& c, const std::vector<double>& x0, const cmaxb_opt_params& prm, std::vector<double>& x_out,
                       cmaxb_opt_result* res) {
  ConjugateFr s;
  c.fuse = prm.fused_trials != 0;
  s.set(c, x0, prm.initial_step, prm.line_tol);
  if (c.rc != CMAXB_OK) return c.rc;
  const double initial_cost = s.f;
  double cost_new = 1e9, cost_old = 1e9;
  int iter = 0, status = kGslContinue, stop = 0;
  do {
    iter++;
    cost_old = cost_new;
    status = s.iterate(c);
    if (c.rc != CMAXB_OK) return c.rc;
    if (status == kGslSuccess) {
      cost_new = s.f;                                        // gsl_multimin_fdfminimizer_minimum
      if (std::fabs(1 - cost_new / (cost_old + 1e-7)) < prm.tolfun) { stop = 1; break; }   // progress tolerance reached
      status = kGslContinue;
    }
    if (nrm2(s.gradient) < prm.epsabs_grad) { stop = 2; break; }                           // gsl_multimin_test_gradient
    if (status != kGslContinue) { stop = 3; break; }                                       // iteration made no progress
  } while (status == kGslContinue && iter < prm.max_iterations);
  x_out = s.x;
  if (res) {
    res->cost_initial = initial_cost;
    res->cost_final = s.f;
    res->iterations = iter;
    res->f_evals = c.f_evals;
    res->g_evals = c.g_evals;
    res->stop_reason = stop;   // 0 iteration limit, 1 cost stagnation, 2 gradient norm, 3 no progress (GSL_ENOPROG)
    res->cost_launches = c.launches;
  }
  return CMAXB_OK;
}

}  // namespace

// The same loop over a caller-supplied cost (to be MINIMISED): the solver without the CUDA cost, for costs that live elsewhere
// and for pinning this loop on the CPU (tests/test_optim.py drives it with the oracle cost and compares with the reference's own
// *_optim_contrast_gsl.cpp compiled over a GSL stand-in).
extern "C" int cmaxb_optimize_callback(int n, const double* x0, cmaxb_cost_f f, cmaxb_cost_fdf fdf, void* user,
                                       const cmaxb_opt_params* params, double* x_out, cmaxb_opt_result* result) {
  if (n <= 0 || !x0 || !f || !fdf || !params || !x_out) return set_error(CMAXB_ERR_INVALID, "null argument / no parameters");
  Cost c;
  c.n = n;
  c.f = [=](const double* x, double* v) { return f(x, n, user, v); };
  c.fdf = [=](const double* x, double* v, double* g) { return fdf(x, n, user, v, g); };
  std::vector<double> xs(x0, x0 + n), xo;
  CMAXB_TRY(run_reference_loop(c, xs, *params, xo, result));
  for (int i = 0; i < n; ++i) x_out[i] = xo[i];
  return CMAXB_OK;
}

extern "C" int cmaxb_fe_optimize(cmaxb_fe* fe, const double omega0[3], const cmaxb_opt_params* params, double omega_out[3],
                                 cmaxb_opt_result* result) {
  if (!fe || !omega0 || !omega_out) return set_error(CMAXB_ERR_INVALID, "null argument");
  // local_optim_contrast_gsl.cpp:106-122: step 0.1, tol 0.05, <= 50 iterations, |g| < 1e-3, rel. change < 1e-4
  cmaxb_opt_params prm{0.1, 0.05, 50, 1e-3, 1e-4, 0};
  if (params) prm = *params;
  Cost c;
  c.n = 3;
  c.f = [fe](const double* x, double* f) { double v; int rc = cmaxb_fe_eval(fe, x, &v, nullptr); *f = -v; return rc; };
  c.fdf = [fe](const double* x, double* f, double* g) {
    double v, gr[3];
    int rc = cmaxb_fe_eval(fe, x, &v, gr);
    *f = -v;
    for (int i = 0; i < 3; ++i) g[i] = -gr[i];
    return rc;
  };
  std::vector<double> x0(omega0, omega0 + 3), xo;
  CMAXB_TRY(run_reference_loop(c, x0, prm, xo, result));
  for (int i = 0; i < 3; ++i) omega_out[i] = xo[i];
  return CMAXB_OK;
}

extern "C" int cmaxb_be_optimize(cmaxb_be* be, const double* x0, int n, const cmaxb_opt_params* params, double* x_out,
                                 cmaxb_opt_result* result) {
  if (!be || !x_out || n <= 0) return set_error(CMAXB_ERR_INVALID, "null argument / no parameters");
  // global_optim_contrast_gsl.cpp:41-53: step 0.1, tol 0.1, <= 50 iterations, |g| < 1e-4, rel. change < 1e-4; x0 = 0
  cmaxb_opt_params prm{0.1, 0.1, 50, 1e-4, 1e-4, 0};
  if (params) prm = *params;
  Cost c;
  c.n = n;
  c.f = [be, n](const double* x, double* f) { double v; int rc = cmaxb_be_eval(be, x, n, &v, nullptr); *f = -v; return rc; };
  c.fdf = [be, n](const double* x, double* f, double* g) {
    double v;
    int rc = cmaxb_be_eval(be, x, n, &v, g);
    *f = -v;
    for (int i = 0; i < n; ++i) g[i] = -g[i];
    return rc;
  };
  std::vector<double> xs(n, 0.0), xo;
  if (x0) xs.assign(x0, x0 + n);
  CMAXB_TRY(run_reference_loop(c, xs, prm, xo, result));
  for (int i = 0; i < n; ++i) x_out[i] = xo[i];
  return CMAXB_OK;
}
