// Instantiation unit: arithmetic type float, 6 residual rows.
#define GRR_REAL float
#define GRR_M 6
#define GRR_OPS_FN ops_f32_m6
#include "grr_inst.inc"
