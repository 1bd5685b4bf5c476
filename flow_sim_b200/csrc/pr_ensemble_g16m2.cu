// pr_ensemble_g16m2.cu - fused ensemble kernel, 2 members per warp (16 lanes each), 2 node(s) per lane:
// reaches of up to 31 nodes.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_FAMILY(16, 2, 16, 0)
