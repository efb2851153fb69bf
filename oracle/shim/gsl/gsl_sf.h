/* oracle/shim/gsl/gsl_sf.h -- TEST INFRASTRUCTURE: included by the reference, nothing used. */
#ifndef BARCODE_ORACLE_SHIM_GSL_SF_H
#define BARCODE_ORACLE_SHIM_GSL_SF_H
#include "gsl_math.h"
#endif
