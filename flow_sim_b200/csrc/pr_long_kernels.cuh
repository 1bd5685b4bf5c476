// pr_long_kernels.cuh - long-reach path (N > 249 nodes: state does not fit one warp's registers).
#pragma once
#include <atomic>
#include <string>

#include "pr_device.cuh"

namespace pr {

inline int long_reach_run(const DevParams& p, bool, cudaStream_t, std::atomic<long long>&, std::string& err) {
  char buf[160];
  snprintf(buf, sizeof buf, "n_nodes=%d: the long-reach (multi-CTA block cyclic reduction) path is not built yet", p.N);
  err = buf;
  return PR_ERR_UNSUPPORTED;
}

}  // namespace pr
