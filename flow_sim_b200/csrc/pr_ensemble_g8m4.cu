// pr_ensemble_g8m4.cu - fused ensemble kernel, 4 members per warp (8 lanes each), 4 node(s) per lane:
// reaches of up to 29 nodes.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_FAMILY(8, 4, 16, 0)
