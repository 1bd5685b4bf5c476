// pr_ensemble_irr_m8.cu - fused ensemble kernel for reaches with IrregularSection (polyline) nodes: 32 lanes per member, 8 node(s) per lane.
#include "pr_ensemble_kernel.cuh"

PR_DEFINE_ENSEMBLE_IRREGULAR(32, 8, 6)
