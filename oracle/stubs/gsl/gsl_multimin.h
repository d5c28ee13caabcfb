// TEST INFRASTRUCTURE -- stand-in for the part of GSL's multimin interface the reference drives
// (gsl_multimin_fdfminimizer_conjugate_fr: alloc / set / iterate / x / minimum / free, gsl_multimin_test_gradient).
// GSL is neither vendored by the reference nor installed here, so the Fletcher-Reeves minimiser behind this interface is a
// RESTATEMENT of GSL 2.x multimin/conjugate_fr.c + directional_minimize.c (set, iterate, take_step, intermediate_point, minimize)
// -- the same algorithm as oracle/gsl_fr.py and csrc/optim.cu, written a third time against GSL's own data structures.
// PARITY UNPINNED against real GSL.  What compiling the reference's *_optim_contrast_gsl*.cpp on top of it pins is the reference's
// OWN loop around the minimiser: callbacks, sign conventions, stopping rules, result hand-over.
#pragma once
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "gsl_vector.h"

enum { GSL_SUCCESS = 0, GSL_FAILURE = -1, GSL_CONTINUE = -2, GSL_ENOPROG = 27 };

typedef struct {
  double (*f)(const gsl_vector* x, void* params);
  void (*df)(const gsl_vector* x, void* params, gsl_vector* df);
  void (*fdf)(const gsl_vector* x, void* params, double* f, gsl_vector* df);
  size_t n;
  void* params;
} gsl_multimin_function_fdf;

typedef struct { const char* name; } gsl_multimin_fdfminimizer_type;

typedef struct {
  int iter;
  double step, max_step, tol;
  gsl_vector *x1, *dx1, *x2, *p, *g0;
  double pnorm, g0norm;
} gslstub_fr_state;

typedef struct {
  const gsl_multimin_fdfminimizer_type* type;
  gsl_multimin_function_fdf* fdf;
  double f;
  gsl_vector* x;
  gsl_vector* gradient;
  gsl_vector* dx;
  void* state;
} gsl_multimin_fdfminimizer;

static const gsl_multimin_fdfminimizer_type gslstub_fr_type = {"conjugate_fr"};
static const gsl_multimin_fdfminimizer_type gslstub_other_type = {"not-implemented"};
static const gsl_multimin_fdfminimizer_type* const gsl_multimin_fdfminimizer_conjugate_fr = &gslstub_fr_type;
static const gsl_multimin_fdfminimizer_type* const gsl_multimin_fdfminimizer_conjugate_pr = &gslstub_other_type;
static const gsl_multimin_fdfminimizer_type* const gsl_multimin_fdfminimizer_vector_bfgs = &gslstub_other_type;
static const gsl_multimin_fdfminimizer_type* const gsl_multimin_fdfminimizer_vector_bfgs2 = &gslstub_other_type;

/* gsl_blas_dnrm2 -> gslcblas cblas_dnrm2 (source_nrm2_r.h): scaled sum of squares */
static inline double gslstub_nrm2(const gsl_vector* v) {
  double scale = 0.0, ssq = 1.0;
  for (size_t i = 0; i < v->size; ++i) {
    const double x = v->data[i];
    if (x != 0.0) {
      const double ax = fabs(x);
      if (scale < ax) { ssq = 1.0 + ssq * (scale / ax) * (scale / ax); scale = ax; }
      else { ssq += (ax / scale) * (ax / scale); }
    }
  }
  return scale * sqrt(ssq);
}
static inline double gslstub_dot(const gsl_vector* a, const gsl_vector* b) { double s = 0; for (size_t i = 0; i < a->size; ++i) s += a->data[i] * b->data[i]; return s; }
static inline void gslstub_copy(gsl_vector* d, const gsl_vector* s) { memcpy(d->data, s->data, sizeof(double) * s->size); }
static inline int gslstub_equal(const gsl_vector* a, const gsl_vector* b) { for (size_t i = 0; i < a->size; ++i) if (a->data[i] != b->data[i]) return 0; return 1; }   /* gsl_vector_equal */
#include <stdio.h>
static inline double gslstub_eval_f(gsl_multimin_function_fdf* fdf, const gsl_vector* x) {
  const double v = fdf->f(x, fdf->params);
  if (getenv("GSLSTUB_TRACE")) { fprintf(stderr, "f  "); for (size_t i = 0; i < x->size && i < 3; ++i) fprintf(stderr, "%.17g ", x->data[i]); fprintf(stderr, "-> %.17g\n", v); }
  return v;
}
static inline void gslstub_eval_df(gsl_multimin_function_fdf* fdf, const gsl_vector* x, gsl_vector* g) {
  fdf->df(x, fdf->params, g);
  if (getenv("GSLSTUB_TRACE")) { fprintf(stderr, "df "); for (size_t i = 0; i < x->size && i < 3; ++i) fprintf(stderr, "%.17g ", x->data[i]); fprintf(stderr, "-> g0 %.17g\n", g->data[0]); }
}
#define GSLSTUB_F(fdf, x) gslstub_eval_f((fdf), (x))
#define GSLSTUB_DF(fdf, x, g) gslstub_eval_df((fdf), (x), (g))

static inline gsl_multimin_fdfminimizer* gsl_multimin_fdfminimizer_alloc(const gsl_multimin_fdfminimizer_type* T, size_t n) {
  gsl_multimin_fdfminimizer* s = (gsl_multimin_fdfminimizer*)malloc(sizeof(*s));
  gslstub_fr_state* st = (gslstub_fr_state*)malloc(sizeof(*st));
  s->type = T; s->fdf = 0; s->f = 0;
  s->x = gsl_vector_alloc(n); s->gradient = gsl_vector_alloc(n); s->dx = gsl_vector_alloc(n);
  st->x1 = gsl_vector_alloc(n); st->dx1 = gsl_vector_alloc(n); st->x2 = gsl_vector_alloc(n); st->p = gsl_vector_alloc(n); st->g0 = gsl_vector_alloc(n);
  s->state = st;
  return s;
}
static inline void gsl_multimin_fdfminimizer_free(gsl_multimin_fdfminimizer* s) {
  gslstub_fr_state* st = (gslstub_fr_state*)s->state;
  gsl_vector_free(st->x1); gsl_vector_free(st->dx1); gsl_vector_free(st->x2); gsl_vector_free(st->p); gsl_vector_free(st->g0);
  free(st);
  gsl_vector_free(s->x); gsl_vector_free(s->gradient); gsl_vector_free(s->dx);
  free(s);
}
// conjugate_fr_set
static inline int gsl_multimin_fdfminimizer_set(gsl_multimin_fdfminimizer* s, gsl_multimin_function_fdf* fdf, const gsl_vector* x,
                                                double step_size, double tol) {
  gslstub_fr_state* st = (gslstub_fr_state*)s->state;
  s->fdf = fdf;
  gslstub_copy(s->x, x);
  gsl_vector_set_zero(s->dx);
  st->iter = 0; st->step = step_size; st->max_step = step_size; st->tol = tol;
  fdf->fdf(s->x, fdf->params, &s->f, s->gradient);
  gslstub_copy(st->p, s->gradient);
  gslstub_copy(st->g0, s->gradient);
  st->pnorm = st->g0norm = gslstub_nrm2(s->gradient);
  return GSL_SUCCESS;
}
// directional_minimize.c: take_step
static inline void gslstub_take_step(const gsl_vector* x, const gsl_vector* p, double step, double lambda, gsl_vector* x1, gsl_vector* dx) {
  for (size_t i = 0; i < x->size; ++i) { dx->data[i] = 0.0 + (-step * lambda) * p->data[i]; x1->data[i] = x->data[i] + 1.0 * dx->data[i]; }
}
// directional_minimize.c: intermediate_point
static inline void gslstub_intermediate_point(gsl_multimin_function_fdf* fdf, const gsl_vector* x, const gsl_vector* p, double lambda, double pg,
                                              double stepa, double stepc, double fa, double fc, gsl_vector* x1, gsl_vector* dx, gsl_vector* gradient,
                                              double* step, double* f) {
  double stepb, fb;
  (void)stepa;
  for (;;) {
    const double u = fabs(pg * lambda * stepc);
    stepb = 0.5 * stepc * u / ((fc - fa) + u);
    gslstub_take_step(x, p, stepb, lambda, x1, dx);
    if (gslstub_equal(x, x1)) { *step = 0; *f = fa; GSLSTUB_DF(fdf, x1, gradient); return; }
    fb = GSLSTUB_F(fdf, x1);
    if (fb >= fa && stepb > 0.0) { fc = fb; stepc = stepb; continue; }
    *step = stepb; *f = fb;
    GSLSTUB_DF(fdf, x1, gradient);
    return;
  }
}
// directional_minimize.c: minimize
static inline void gslstub_minimize(gsl_multimin_function_fdf* fdf, const gsl_vector* x, const gsl_vector* p, double lambda, double stepa, double stepb,
                                    double stepc, double fa, double fb, double fc, double tol, gsl_vector* x1, gsl_vector* dx1, gsl_vector* x2,
                                    gsl_vector* dx2, gsl_vector* gradient, double* step, double* f, double* gnorm) {
  double u = stepb, v = stepa, w = stepc, fu = fb, fv = fa, fw = fc;
  double old2 = fabs(w - v), old1 = fabs(v - u);
  double stepm, fm, pg, gnorm1;
  int iter = 0;
  gslstub_copy(x2, x1);
  gslstub_copy(dx2, dx1);
  *f = fb; *step = stepb; *gnorm = gslstub_nrm2(gradient);
mid_trial:
  iter++;
  if (iter > 10) return;
  {
    const double dw = w - u, dv = v - u;
    double du = 0.0;
    const double e1 = ((fv - fu) * dw * dw + (fu - fw) * dv * dv);
    const double e2 = 2.0 * ((fv - fu) * dw + (fu - fw) * dv);
    if (e2 != 0.0) du = e1 / e2;
    if (du > 0.0 && du < (stepc - stepb) && fabs(du) < 0.5 * old2) stepm = u + du;
    else if (du < 0.0 && du > (stepa - stepb) && fabs(du) < 0.5 * old2) stepm = u + du;
    else if ((stepc - stepb) > (stepb - stepa)) stepm = 0.38 * (stepc - stepb) + stepb;
    else stepm = stepb - 0.38 * (stepb - stepa);
  }
  gslstub_take_step(x, p, stepm, lambda, x1, dx1);
  fm = GSLSTUB_F(fdf, x1);
  if (fm > fb) {
    if (fm < fv) { w = v; v = stepm; fw = fv; fv = fm; }
    else if (fm < fw) { w = stepm; fw = fm; }
    if (stepm < stepb) { stepa = stepm; fa = fm; } else { stepc = stepm; fc = fm; }
    goto mid_trial;
  } else if (fm <= fb) {
    old2 = old1; old1 = fabs(u - stepm);
    w = v; v = u; u = stepm;
    fw = fv; fv = fu; fu = fm;
    gslstub_copy(x2, x1);
    gslstub_copy(dx2, dx1);
    GSLSTUB_DF(fdf, x1, gradient);
    pg = gslstub_dot(p, gradient);
    gnorm1 = gslstub_nrm2(gradient);
    *f = fm; *step = stepm; *gnorm = gnorm1;
    if (fabs(pg * lambda / gnorm1) < tol) return;
    if (stepm < stepb) { stepc = stepb; fc = fb; stepb = stepm; fb = fm; }
    else { stepa = stepb; fa = fb; stepb = stepm; fb = fm; }
    goto mid_trial;
  }
}
// conjugate_fr_iterate
static inline int gsl_multimin_fdfminimizer_iterate(gsl_multimin_fdfminimizer* s) {
  gslstub_fr_state* st = (gslstub_fr_state*)s->state;
  gsl_multimin_function_fdf* fdf = s->fdf;
  gsl_vector *x = s->x, *gradient = s->gradient, *dx = s->dx, *x1 = st->x1, *dx1 = st->dx1, *x2 = st->x2, *p = st->p, *g0 = st->g0;
  const double pnorm = st->pnorm, g0norm = st->g0norm;
  double fa = s->f, fb, fc, dir, stepa = 0.0, stepb, stepc = st->step, tol = st->tol, g1norm, pg;
  if (pnorm == 0.0 || g0norm == 0.0) { gsl_vector_set_zero(dx); return GSL_ENOPROG; }
  pg = gslstub_dot(p, gradient);
  dir = (pg >= 0.0) ? +1.0 : -1.0;
  gslstub_take_step(x, p, stepc, dir / pnorm, x1, dx);
  fc = GSLSTUB_F(fdf, x1);
  if (fc < fa) {
    st->step = stepc * 2.0;
    s->f = fc;
    gslstub_copy(x, x1);
    GSLSTUB_DF(fdf, x1, gradient);
    return GSL_SUCCESS;
  }
  gslstub_intermediate_point(fdf, x, p, dir / pnorm, pg, stepa, stepc, fa, fc, x1, dx1, gradient, &stepb, &fb);
  if (stepb == 0.0) return GSL_ENOPROG;
  gslstub_minimize(fdf, x, p, dir / pnorm, stepa, stepb, stepc, fa, fb, fc, tol, x1, dx1, x2, dx, gradient, &st->step, &s->f, &g1norm);
  gslstub_copy(x, x2);
  st->iter = (st->iter + 1) % (int)x->size;
  if (st->iter == 0) { gslstub_copy(p, gradient); st->pnorm = g1norm; }
  else {
    const double beta = -pow(g1norm / g0norm, 2.0);
    for (size_t i = 0; i < p->size; ++i) p->data[i] = (-beta) * p->data[i];            /* gsl_blas_dscal(-beta, p) */
    for (size_t i = 0; i < p->size; ++i) p->data[i] = p->data[i] + 1.0 * gradient->data[i];   /* gsl_blas_daxpy(1.0, gradient, p) */
    st->pnorm = gslstub_nrm2(p);
  }
  st->g0norm = g1norm;
  gslstub_copy(g0, gradient);
  return GSL_SUCCESS;
}
static inline gsl_vector* gsl_multimin_fdfminimizer_x(const gsl_multimin_fdfminimizer* s) { return s->x; }
static inline double gsl_multimin_fdfminimizer_minimum(const gsl_multimin_fdfminimizer* s) { return s->f; }
static inline gsl_vector* gsl_multimin_fdfminimizer_gradient(const gsl_multimin_fdfminimizer* s) { return s->gradient; }
static inline int gsl_multimin_test_gradient(const gsl_vector* g, double epsabs) { return gslstub_nrm2(g) < epsabs ? GSL_SUCCESS : GSL_CONTINUE; }
