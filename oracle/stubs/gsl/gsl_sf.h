/* stub: see gsl_rng.h */
