// TEST INFRASTRUCTURE ONLY, NOT Rcpp: forwards to the stub in oracle/stubs/Rcpp.h.
#include <Rcpp.h>
