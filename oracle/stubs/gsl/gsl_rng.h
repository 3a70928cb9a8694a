/* Stub so that the reference's gfl_utils.h parses without GSL installed.
 * Only the opaque type name is needed by the prototypes in gfl_utils.h; no
 * GSL symbol is referenced by gfl_tf.c (the only file compiled into oracle/_ref). */
#ifndef ORACLE_STUB_GSL_RNG_H
#define ORACLE_STUB_GSL_RNG_H
typedef struct gsl_rng_stub_s gsl_rng;
#endif
