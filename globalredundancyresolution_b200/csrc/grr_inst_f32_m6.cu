// Instantiation unit: arithmetic type float, 6 residual rows.
#define GRR_REAL float
#define GRR_M 6
#define GRR_OPS_FN ops_f32_m6
#define GRR_FOR_EACH_N(X) X(6) X(7)
#include "grr_inst.inc"
