// pr_ensemble_g8m1.cu - fused ensemble kernel, 4 members per warp (8 lanes each), 1 node(s) per lane:
// reaches of up to 8 nodes.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_FAMILY(8, 1, 16, 0)
