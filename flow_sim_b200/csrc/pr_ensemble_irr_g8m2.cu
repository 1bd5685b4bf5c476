// pr_ensemble_irr_g8m2.cu - fused ensemble kernel for reaches with IrregularSection (polyline) nodes: 8 lanes per member (4 members per warp), 2 node(s) per lane.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_IRREGULAR(8, 2, 8)
