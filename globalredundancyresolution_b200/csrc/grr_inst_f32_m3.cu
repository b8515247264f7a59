// Instantiation unit: arithmetic type float, 3 residual rows.
#define GRR_REAL float
#define GRR_M 3
#define GRR_OPS_FN ops_f32_m3
#include "grr_inst.inc"
