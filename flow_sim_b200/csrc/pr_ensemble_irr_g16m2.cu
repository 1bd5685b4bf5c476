// pr_ensemble_irr_g16m2.cu - fused ensemble kernel for reaches with IrregularSection (polyline) nodes: 16 lanes per member (2 members per warp), 2 node(s) per lane.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_IRREGULAR(16, 2, 8)
