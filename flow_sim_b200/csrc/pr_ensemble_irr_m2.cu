// pr_ensemble_irr_m2.cu - fused ensemble kernel for reaches with IrregularSection (polyline) nodes: 32 lanes per member, 2 node(s) per lane.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_IRREGULAR(32, 2, 8)
