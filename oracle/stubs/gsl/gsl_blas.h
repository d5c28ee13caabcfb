// TEST INFRASTRUCTURE stub (the reference includes gsl_blas.h but calls none of it)
#pragma once
#include "gsl_vector.h"
