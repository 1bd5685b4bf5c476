// pr_ensemble_g16m4.cu - fused ensemble kernel, 2 members per warp (16 lanes each), 4 node(s) per lane:
// reaches of up to 61 nodes.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_FAMILY(16, 4, 16, 0)
