// pr_ensemble_m8.cu - instantiations of the fused ensemble kernel with 8 node(s) per lane (5 warps per CTA).
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_FAMILY(32, 8, 7, 0)
