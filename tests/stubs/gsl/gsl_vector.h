/* Minimal stand-in for <gsl/gsl_vector.h> (GSL is not installed here): just enough to compile
 * include/cmax_b200_gsl.hpp in tests/test_gsl_adapter.py.  Field layout follows GSL 2.x. */
#pragma once
#include <stddef.h>
typedef struct { size_t size; size_t stride; double* data; void* block; int owner; } gsl_vector;
static inline double gsl_vector_get(const gsl_vector* v, size_t i) { return v->data[i * v->stride]; }
static inline void gsl_vector_set(gsl_vector* v, size_t i, double x) { v->data[i * v->stride] = x; }
