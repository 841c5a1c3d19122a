// Single translation unit of libgantts_b200.so (kernels in different files launch each other's
// helpers, so they are compiled together rather than with relocatable device code).
#include "core.cu"
#include "mlpg.cu"
#include "losses.cu"
#include "linear_simt.cu"
#include "gemm_tc.cu"
#include "mlp_tc.cu"
#include "linear.cu"
#include "optim.cu"
#include "gan_step.cu"
#include "lstm.cu"
#include "sru.cu"
