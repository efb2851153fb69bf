/* oracle/shim/gsl/gsl_errno.h -- TEST INFRASTRUCTURE: placeholder. */
#ifndef BARCODE_ORACLE_SHIM_GSL_ERRNO_H
#define BARCODE_ORACLE_SHIM_GSL_ERRNO_H
#endif
