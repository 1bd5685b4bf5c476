// pr_ensemble_m1.cu - instantiations of the fused ensemble kernel with 1 node(s) per lane (12 warps per CTA).
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_FAMILY(32, 1, 16, 0)
