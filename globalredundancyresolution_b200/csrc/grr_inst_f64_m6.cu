// Instantiation unit: arithmetic type double, 6 residual rows.
#define GRR_REAL double
#define GRR_M 6
#define GRR_OPS_FN ops_f64_m6
#include "grr_inst.inc"
