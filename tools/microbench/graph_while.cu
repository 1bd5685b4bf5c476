#include <cuda_runtime.h>
#include <cstdio>
__global__ void setc(cudaGraphConditionalHandle h, int* counter) {
  if (threadIdx.x == 0) { int c = atomicAdd(counter, 1); cudaGraphSetConditional(h, c < 5 ? 1u : 0u); }
}
int main() {
  cudaGraph_t g; cudaGraphCreate(&g, 0);
  cudaGraphConditionalHandle h; cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
  cudaGraphNodeParams np = {}; np.type = cudaGraphNodeTypeConditional; np.conditional.handle = h; np.conditional.type = cudaGraphCondTypeWhile; np.conditional.size = 1;
  cudaGraphNode_t node; cudaError_t e = cudaGraphAddNode(&node, g, nullptr, 0, &np);
  printf("add node: %s\n", cudaGetErrorString(e));
  cudaGraph_t body = np.conditional.phGraph_out[0];
  int* ctr; cudaMalloc(&ctr, 4); cudaMemset(ctr, 0, 4);
  cudaStream_t s; cudaStreamCreate(&s);
  cudaGraphNode_t kn; cudaKernelNodeParams kp = {}; void* args[] = {&h, &ctr}; kp.func = (void*)setc; kp.gridDim = dim3(1); kp.blockDim = dim3(32); kp.kernelParams = args;
  e = cudaGraphAddKernelNode(&kn, body, nullptr, 0, &kp); printf("kernel node: %s\n", cudaGetErrorString(e));
  cudaGraphExec_t ex; e = cudaGraphInstantiate(&ex, g, 0); printf("inst: %s\n", cudaGetErrorString(e));
  e = cudaGraphLaunch(ex, s); cudaStreamSynchronize(s); int hc = 0; cudaMemcpy(&hc, ctr, 4, cudaMemcpyDeviceToHost); printf("launch: %s, counter %d (expect 6)\n", cudaGetErrorString(e), hc);
  return 0;
}
