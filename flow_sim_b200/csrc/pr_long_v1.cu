// pr_long_v1.cu - long-reach (tiled) path, kernels and driver loop for <compound, curvature, irregular> = <true, false, false>.
#include "pr_long_kernels.cuh"

template int pr::long_reach_run_t<true, false, false>(const pr::DevParams&, cudaStream_t, std::atomic<long long>&, std::string&);
