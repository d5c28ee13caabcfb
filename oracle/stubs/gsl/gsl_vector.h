// TEST INFRASTRUCTURE stub: the type named in the reference's callback declarations (the GSL solve itself is not compiled)
#pragma once
#include <stddef.h>
typedef struct { size_t size; size_t stride; double* data; void* block; int owner; } gsl_vector;
