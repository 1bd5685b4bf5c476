// pr_ensemble_m4.cu - instantiations of the fused ensemble kernel with 4 node(s) per lane (16 warps per CTA).
#include "pr_ensemble_kernel.cuh"

#ifndef PR_W4
#define PR_W4 16
#endif
PR_DEFINE_ENSEMBLE_FAMILY(32, 4, PR_W4, 1)

#ifdef PR_VARIANT_SHIM   // tuning builds (tools/ab_build.sh): this family alone, loaded through PR_M4_VARIANT
extern "C" int pr_variant_launch_m4(const pr::DevParams* p, int curv, cudaStream_t s) {
  return pr::launch_ensemble_family<32, 4, PR_W4>(*p, curv != 0, s);
}
#endif
