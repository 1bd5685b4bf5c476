// pr_long_v4.cu - long-reach (tiled) path, kernels and driver loop for <compound, curvature, irregular> = <true, true, true>.
#include "pr_long_kernels.cuh"

template int pr::long_reach_run_t<true, true, true>(const pr::DevParams&, cudaStream_t, std::atomic<long long>&, std::string&);
