#define MFHN_NUMBER double
#define MFHN_RUN_GENERIC run_generic_f64
#include "k_generic.inc"
