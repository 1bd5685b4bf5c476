// pr_long_v3.cu - long-reach (tiled) path, kernels and driver loop for <compound, curvature, irregular> = <true, false, true>.
#include "pr_long_kernels.cuh"

template int pr::long_reach_run_t<true, false, true>(const pr::DevParams&, cudaStream_t, std::atomic<long long>&, std::string&);
