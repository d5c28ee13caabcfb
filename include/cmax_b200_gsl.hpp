// cmax_b200_gsl.hpp -- reference-side binding: the GSL f / df / fdf callback triple of CMax-SLAM
// re-expressed over libcmax_b200.so.  Header-only C++17; meant to be compiled INSIDE the reference
// tree (it needs <gsl/gsl_vector.h>, dlopen).  See INTEGRATION.md for the exact edit sites.
//
// Replaces the bodies of
//   local_contrast_fdf / _f / _df    src/frontend/local_optim_contrast_gsl.cpp:20-70
//   global_contrast_fdf / _f / _df   src/backend/global_optim_contrast_gsl_analytical.cpp:17-81
// keeping their signatures and sign convention (they return -contrast and -gradient).
#pragma once
#include <dlfcn.h>
#include <gsl/gsl_vector.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "cmax_b200.h"

namespace cmaxb_gsl {

// Function table resolved once with dlopen/dlsym (e.g. in CMaxSLAM::CMaxSLAM, src/cmax_slam.cpp:85-96).
struct Api {
  void* dl = nullptr;
  decltype(&cmaxb_fe_create) fe_create = nullptr;
  decltype(&cmaxb_fe_destroy) fe_destroy = nullptr;
  decltype(&cmaxb_fe_set_packet) fe_set_packet = nullptr;
  decltype(&cmaxb_fe_eval) fe_eval = nullptr;
  decltype(&cmaxb_fe_get_iwe) fe_get_iwe = nullptr;
  decltype(&cmaxb_be_create) be_create = nullptr;
  decltype(&cmaxb_be_destroy) be_destroy = nullptr;
  decltype(&cmaxb_be_set_window) be_set_window = nullptr;
  decltype(&cmaxb_be_eval) be_eval = nullptr;
  decltype(&cmaxb_be_get_il) be_get_il = nullptr;
  decltype(&cmaxb_be_get_alpha) be_get_alpha = nullptr;
  decltype(&cmaxb_last_error) last_error = nullptr;

  bool load(const char* path = "libcmax_b200.so") {
    dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!dl) { std::fprintf(stderr, "cmax_b200: %s\n", dlerror()); return false; }
    bool ok = true;
#define CMAXB_SYM(field, name) field = reinterpret_cast<decltype(field)>(dlsym(dl, name)); ok = ok && field != nullptr;
    CMAXB_SYM(fe_create, "cmaxb_fe_create") CMAXB_SYM(fe_destroy, "cmaxb_fe_destroy")
    CMAXB_SYM(fe_set_packet, "cmaxb_fe_set_packet") CMAXB_SYM(fe_eval, "cmaxb_fe_eval")
    CMAXB_SYM(fe_get_iwe, "cmaxb_fe_get_iwe") CMAXB_SYM(be_create, "cmaxb_be_create")
    CMAXB_SYM(be_destroy, "cmaxb_be_destroy") CMAXB_SYM(be_set_window, "cmaxb_be_set_window")
    CMAXB_SYM(be_eval, "cmaxb_be_eval") CMAXB_SYM(be_get_il, "cmaxb_be_get_il")
    CMAXB_SYM(be_get_alpha, "cmaxb_be_get_alpha") CMAXB_SYM(last_error, "cmaxb_last_error")
#undef CMAXB_SYM
    return ok;
  }
};

// `params` of the gsl_multimin_function_fdf: what the estimator object hands to GSL instead of `this`.
struct FeParams { const Api* api; cmaxb_fe* fe; };
struct BeParams { const Api* api; cmaxb_be* be; int n_params; };

[[noreturn]] inline void die(const Api* api, const char* where) {
  // the reference aborts on its glog CHECKs; keep that behaviour at the seam
  std::fprintf(stderr, "cmax_b200 %s: %s\n", where, api->last_error());
  std::abort();
}

// ---- front-end: drop-in bodies for local_optim_contrast_gsl.cpp:20-70 -------------------------------
inline void local_contrast_fdf(const gsl_vector* v, void* ptr, double* f, gsl_vector* df) {
  auto* p = static_cast<FeParams*>(ptr);
  const double omega[3] = {gsl_vector_get(v, 0), gsl_vector_get(v, 1), gsl_vector_get(v, 2)};
  double contrast = 0, grad[3] = {0, 0, 0};
  if (p->api->fe_eval(p->fe, omega, &contrast, df ? grad : nullptr) != CMAXB_OK) die(p->api, "fe_eval");
  *f = -contrast;                                   // change sign: minimize -contrast
  if (df) for (int i = 0; i < 3; ++i) gsl_vector_set(df, i, -grad[i]);
}
inline double local_contrast_f(const gsl_vector* v, void* ptr) {
  double cost;
  local_contrast_fdf(v, ptr, &cost, nullptr);
  return cost;
}
inline void local_contrast_df(const gsl_vector* v, void* ptr, gsl_vector* df) {
  double cost;
  local_contrast_fdf(v, ptr, &cost, df);
}

// ---- back-end: drop-in bodies for global_optim_contrast_gsl_analytical.cpp:17-81 ---------------------
inline void global_contrast_fdf(const gsl_vector* v, void* adata, double* f, gsl_vector* df) {
  auto* p = static_cast<BeParams*>(adata);
  std::vector<double> x(p->n_params), g(df ? p->n_params : 0);
  for (int i = 0; i < p->n_params; ++i) x[i] = gsl_vector_get(v, i);
  double contrast = 0;
  if (p->api->be_eval(p->be, x.data(), p->n_params, &contrast, df ? g.data() : nullptr) != CMAXB_OK) die(p->api, "be_eval");
  *f = -contrast;
  if (df) for (int i = 0; i < p->n_params; ++i) gsl_vector_set(df, i, -g[i]);
}
inline double global_contrast_f(const gsl_vector* v, void* adata) {
  double cost;
  global_contrast_fdf(v, adata, &cost, nullptr);
  return cost;
}
inline void global_contrast_df(const gsl_vector* v, void* adata, gsl_vector* df) {
  double cost;
  global_contrast_fdf(v, adata, &cost, df);
}

}  // namespace cmaxb_gsl
