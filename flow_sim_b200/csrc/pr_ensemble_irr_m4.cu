// pr_ensemble_irr_m4.cu - fused ensemble kernel for reaches with IrregularSection (polyline) nodes: 32 lanes per member, 4 node(s) per lane.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_IRREGULAR(32, 4, 8)
