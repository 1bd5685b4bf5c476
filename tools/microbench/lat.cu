// Dependent-issue latencies on sm_100a (cycles per op in a serial chain, one warp).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void lat(double* out, long long* cyc, double a, double b, int n) {
  double x = threadIdx.x * 1e-3 + 1.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); x = fma(x, a, b); }
  long long t1 = clock64();
  double y = x;
  for (int i = 0; i < n; ++i) { y = y * a; y = y * a; y = y * a; y = y * a; }
  long long t2 = clock64();
  double z = y;
  for (int i = 0; i < n; ++i) { z = z + b; z = z + b; z = z + b; z = z + b; }
  long long t3 = clock64();
  double w = z;
  for (int i = 0; i < n; ++i) {
    w = __shfl_xor_sync(0xffffffffu, w, 1); w = __shfl_xor_sync(0xffffffffu, w, 2);
    w = __shfl_xor_sync(0xffffffffu, w, 4); w = __shfl_xor_sync(0xffffffffu, w, 8);
  }
  long long t4 = clock64();
  double r = w + 3.0;
  for (int i = 0; i < n; ++i) {
    double s; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(r)); r = s + 1.5;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(r)); r = s + 1.5;
  }
  long long t5 = clock64();
  // 2 independent chains interleaved
  double p0 = r, p1 = r + 1;
  for (int i = 0; i < n; ++i) { p0 = fma(p0, a, b); p1 = fma(p1, a, b); p0 = fma(p0, a, b); p1 = fma(p1, a, b); }
  long long t6 = clock64();
  double q0 = p0, q1 = p1, q2 = p0 + 1, q3 = p1 + 1;
  for (int i = 0; i < n; ++i) { q0 = fma(q0, a, b); q1 = fma(q1, a, b); q2 = fma(q2, a, b); q3 = fma(q3, a, b); }
  long long t7 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = q0 + q1 + q2 + q3 + x + y + z + w;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 64);
  const int n = 4096;
  for (int warps = 1; warps <= 4; warps *= 2) {
    lat<<<1, 32 * warps>>>(out, cyc, 1.0000001, 1e-9, n);
    cudaDeviceSynchronize();
    long long h[7]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    printf("warps/CTA=%d (1 CTA): DFMA chain %.2f cyc/op, DMUL %.2f, DADD %.2f, SHFL.64 %.2f, RCP64H+DADD %.2f, 2xDFMA ILP %.2f cyc/op, 4xDFMA ILP %.2f cyc/op\n",
           warps, h[0] / (4.0 * n), h[1] / (4.0 * n), h[2] / (4.0 * n), h[3] / (4.0 * n), h[4] / (2.0 * n), h[5] / (4.0 * n), h[6] / (4.0 * n));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
