// Instantiation unit: arithmetic type double, 3 residual rows.
#define GRR_REAL double
#define GRR_M 3
#define GRR_OPS_FN ops_f64_m3
#include "grr_inst.inc"
