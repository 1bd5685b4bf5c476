// pr_ensemble_m4.cu - instantiations of the fused ensemble kernel with 4 node(s) per lane (12 warps per CTA).
#include "pr_ensemble_kernel.cuh"

#ifndef PR_W4
#define PR_W4 16
#endif
PR_DEFINE_ENSEMBLE_FAMILY(32, 4, PR_W4)
