// TEST INFRASTRUCTURE -- stand-in for the gsl_vector subset the reference uses (GSL is not installed here)
#pragma once
#include <stddef.h>
#include <stdlib.h>
typedef struct { size_t size; size_t stride; double* data; void* block; int owner; } gsl_vector;
static inline gsl_vector* gsl_vector_alloc(size_t n) {
  gsl_vector* v = (gsl_vector*)malloc(sizeof(gsl_vector));
  v->size = n; v->stride = 1; v->data = (double*)calloc(n ? n : 1, sizeof(double)); v->block = 0; v->owner = 1;
  return v;
}
static inline void gsl_vector_free(gsl_vector* v) { if (v) { free(v->data); free(v); } }
static inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i * v->stride] = x; }
static inline void gsl_vector_set_zero(gsl_vector* v) { for (size_t i = 0; i < v->size; ++i) v->data[i * v->stride] = 0.0; }
