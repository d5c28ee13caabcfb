"""TEST INFRASTRUCTURE.  Independent Python restatement of the optimiser the reference drives its cost with:
GSL 2.x `gsl_multimin_fdfminimizer_conjugate_fr` (multimin/conjugate_fr.c + directional_minimize.c) and the
reference's loop around it (src/frontend/local_optim_contrast_gsl.cpp:119-204,
src/backend/global_optim_contrast_gsl.cpp:51-113).  GSL is not vendored by the reference and not installed
here: PARITY UNPINNED against real GSL iterates; used to check the library's C++ loop (csrc/optim.cu)."""
import math

import numpy as np

GSL_SUCCESS, GSL_CONTINUE, GSL_ENOPROG = 0, -2, 27


class Counter:
    def __init__(self, f, fdf):
        self._f, self._fdf, self.f_evals, self.g_evals = f, fdf, 0, 0

    def f(self, x):
        self.f_evals += 1
        return float(self._f(x))

    def df(self, x):
        self.g_evals += 1
        return np.array(self._fdf(x)[1], dtype=np.float64)

    def fdf(self, x):
        self.f_evals += 1
        self.g_evals += 1
        v, g = self._fdf(x)
        return float(v), np.array(g, dtype=np.float64)


def _dot(a, b):
    """Sequential sum of products, left to right -- the operation order of the reference CBLAS ddot and of the C / C++
    restatements (oracle/stubs/gsl/gsl_multimin.h, csrc/optim.cu); numpy's BLAS dot may fuse or re-associate."""
    s = 0.0
    for u, v in zip(a.tolist(), b.tolist()):
        s += u * v
    return s


def _nrm2(a):
    """gsl_blas_dnrm2 -> gslcblas cblas_dnrm2 (source_nrm2_r.h): scaled sum of squares."""
    scale, ssq = 0.0, 1.0
    for x in a.tolist():
        if x != 0.0:
            ax = abs(x)
            if scale < ax:
                ssq = 1.0 + ssq * (scale / ax) * (scale / ax)
                scale = ax
            else:
                ssq += (ax / scale) * (ax / scale)
    return scale * math.sqrt(ssq)


def _take_step(x, p, step, lam):
    dx = 0.0 + (-step * lam) * p
    return x + 1.0 * dx, dx


def _intermediate_point(c, x, p, lam, pg, stepa, stepc, fa, fc):
    while True:
        u = abs(pg * lam * stepc)
        stepb = 0.5 * stepc * u / ((fc - fa) + u)
        x1, dx = _take_step(x, p, stepb, lam)
        if np.array_equal(x, x1):
            return 0.0, fa, x1, dx, c.df(x1)
        fb = c.f(x1)
        if fb >= fa and stepb > 0.0:
            fc, stepc = fb, stepb
            continue
        return stepb, fb, x1, dx, c.df(x1)


def _minimize(c, x, p, lam, stepa, stepb, stepc, fa, fb, fc, tol, x1, dx1, gradient):
    u, v, w = stepb, stepa, stepc
    fu, fv, fw = fb, fa, fc
    old2, old1 = abs(w - v), abs(v - u)
    x2, dx2 = x1.copy(), dx1.copy()
    f, step, gnorm = fb, stepb, _nrm2(gradient)
    it = 0
    while True:
        it += 1
        if it > 10:
            return x2, dx2, gradient, step, f, gnorm
        dw, dv, du = w - u, v - u, 0.0
        e1 = (fv - fu) * dw * dw + (fu - fw) * dv * dv
        e2 = 2.0 * ((fv - fu) * dw + (fu - fw) * dv)
        if e2 != 0.0:
            du = e1 / e2
        if du > 0.0 and du < (stepc - stepb) and abs(du) < 0.5 * old2:
            stepm = u + du
        elif du < 0.0 and du > (stepa - stepb) and abs(du) < 0.5 * old2:
            stepm = u + du
        elif (stepc - stepb) > (stepb - stepa):
            stepm = 0.38 * (stepc - stepb) + stepb
        else:
            stepm = stepb - 0.38 * (stepb - stepa)
        x1, dx1 = _take_step(x, p, stepm, lam)
        fm = c.f(x1)
        if fm > fb:
            if fm < fv:
                w, v, fw, fv = v, stepm, fv, fm
            elif fm < fw:
                w, fw = stepm, fm
            if stepm < stepb:
                stepa, fa = stepm, fm
            else:
                stepc, fc = stepm, fm
            continue
        old2, old1 = old1, abs(u - stepm)
        w, v, u = v, u, stepm
        fw, fv, fu = fv, fu, fm
        x2, dx2 = x1.copy(), dx1.copy()
        gradient = c.df(x1)
        pg = _dot(p, gradient)
        gnorm1 = _nrm2(gradient)
        f, step, gnorm = fm, stepm, gnorm1
        if abs(pg * lam / gnorm1) < tol:
            return x2, dx2, gradient, step, f, gnorm
        if stepm < stepb:
            stepc, fc, stepb, fb = stepb, fb, stepm, fm
        else:
            stepa, fa, stepb, fb = stepb, fb, stepm, fm


def minimize_fr(f, fdf, x0, initial_step=0.1, line_tol=0.05, max_iterations=50, epsabs_grad=1e-3, tolfun=1e-4):
    """f(x)->cost, fdf(x)->(cost, grad).  Returns (x, stats dict) like cmaxb_*_optimize."""
    c = Counter(f, fdf)
    x = np.array(x0, dtype=np.float64)
    n = len(x)
    it_state, step, tol = 0, initial_step, line_tol
    fval, gradient = c.fdf(x)
    p, g0 = gradient.copy(), gradient.copy()
    pnorm = g0norm = _nrm2(gradient)
    cost_initial = fval
    cost_new = cost_old = 1e9
    it, stop = 0, 0
    trace = [x.copy()]
    while True:
        it += 1
        cost_old = cost_new
        # ---- conjugate_fr_iterate
        fa, stepa, stepc = fval, 0.0, step
        if pnorm == 0.0 or g0norm == 0.0:
            status = GSL_ENOPROG
        else:
            pg = _dot(p, gradient)
            direction = 1.0 if pg >= 0.0 else -1.0
            x1, dx = _take_step(x, p, stepc, direction / pnorm)
            fc = c.f(x1)
            if fc < fa:
                step, fval, x = stepc * 2.0, fc, x1
                gradient = c.df(x1)
                status = GSL_SUCCESS
            else:
                stepb, fb, x1, dx1, gradient = _intermediate_point(c, x, p, direction / pnorm, pg, stepa, stepc, fa, fc)
                if stepb == 0.0:
                    status = GSL_ENOPROG
                else:
                    x2, dx, gradient, step, fval, g1norm = _minimize(c, x, p, direction / pnorm, stepa, stepb, stepc, fa, fb, fc,
                                                                      tol, x1, dx1, gradient)
                    x = x2
                    it_state = (it_state + 1) % n
                    if it_state == 0:
                        p, pnorm = gradient.copy(), g1norm
                    else:
                        beta = -math.pow(g1norm / g0norm, 2.0)
                        p = (-beta) * p
                        p = p + 1.0 * gradient
                        pnorm = _nrm2(p)
                    g0norm, g0 = g1norm, gradient.copy()
                    status = GSL_SUCCESS
        trace.append(x.copy())
        # ---- the reference's stopping rules
        if status == GSL_SUCCESS:
            cost_new = fval
            if abs(1 - cost_new / (cost_old + 1e-7)) < tolfun:
                stop = 1
                break
            status = GSL_CONTINUE
        if _nrm2(gradient) < epsabs_grad:
            stop = 2
            break
        if status != GSL_CONTINUE:
            stop = 3
            break
        if not (status == GSL_CONTINUE and it < max_iterations):
            break
    return x, {"cost_initial": cost_initial, "cost_final": fval, "iterations": it, "f_evals": c.f_evals, "g_evals": c.g_evals,
               "stop_reason": stop, "trace": np.array(trace)}
