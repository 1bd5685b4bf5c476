// pr_long_v0.cu - long-reach (tiled) path, kernels and driver loop for <compound, curvature, irregular> = <false, false, false>.
#include "pr_long_kernels.cuh"

template int pr::long_reach_run_t<false, false, false>(const pr::DevParams&, cudaStream_t, std::atomic<long long>&, std::string&);
