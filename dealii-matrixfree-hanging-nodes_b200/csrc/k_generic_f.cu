#define MFHN_NUMBER float
#define MFHN_RUN_GENERIC run_generic_f32
#include "k_generic.inc"
