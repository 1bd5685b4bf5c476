// pr_ensemble_g8m2.cu - fused ensemble kernel, 4 members per warp (8 lanes each), 2 node(s) per lane:
// reaches of up to 15 nodes.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_FAMILY(8, 2, 16, 0)
