// Instantiation unit: arithmetic type float, 3 residual rows.
#define GRR_REAL float
#define GRR_M 3
#define GRR_OPS_FN ops_f32_m3
#define GRR_FOR_EACH_N(X) X(2) X(3) X(5) X(10) X(20) X(6) X(7)
#include "grr_inst.inc"
