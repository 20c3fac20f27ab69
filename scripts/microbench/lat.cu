// Dependent-issue latencies on sm_100a that shape the leaf kernels: DFMA, DADD, DMUL, SHFL, LDS, MUFU.RSQ64H/RCP64H.
// One warp, one block; clock64 around a long dependent chain.  nvcc -arch=sm_100a -O3 lat.cu -o lat
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_dfma(double* out, long long* cyc, double a, double b) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dfma_ilp(double* out, long long* cyc, double a, double b, int dummy) {
  double x0 = out[threadIdx.x], x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0+4, x5=x0+5, x6=x0+6, x7=x0+7;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) { x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
    x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b); }
  long long t1 = clock64();
  out[threadIdx.x] = x0 + x1 + x2 + x3+x4+x5+x6+x7; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd(double* out, long long* cyc, double a) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = x + a;
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc) {
  double x = out[threadIdx.x];
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_tp(double* out, long long* cyc) {  // 8 independent double shuffles per iteration
  double x[8];
  for (int j = 0; j < 8; ++j) x[j] = out[threadIdx.x] + j;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __shfl_sync(0xffffffffu, x[j], (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) p = idx[p];
  long long t1 = clock64();
  out[threadIdx.x] = p; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rsqrt(double* out, long long* cyc) {
  double x = out[threadIdx.x] + 2.0;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) { double r; asm volatile("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp(double* out, long long* cyc) {
  double x = out[threadIdx.x] + 2.0;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) { double r; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r; }
  long long t1 = clock64();
  out[threadIdx.x] = x; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
// throughput with W warps per SM sub-partition: 8 independent DFMA chains per thread
__global__ void k_dfma_tp(double* out, long long* cyc, double a, double b) {
  double x[8];
  for (int j = 0; j < 8; ++j) x[j] = out[threadIdx.x] + j;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fma(x[j], a, b);
  long long t1 = clock64();
  double s = 0; for (int j = 0; j < 8; ++j) s += x[j];
  out[threadIdx.x] = s; if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8); cudaMemset(out, 0, 8 * 1024);
  long long h;
#define RUN(name, threads, per, ...) for (int rep = 0; rep < 2; ++rep) { name<<<1, threads>>>(__VA_ARGS__); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); } \
  printf("%-28s threads=%4d  cycles/op = %.2f\n", #name, threads, (double)h / (N * per));
  RUN(k_dfma, 32, 1, out, cyc, 1.0000001, 1e-9)
  RUN(k_dfma_ilp, 32, 8, out, cyc, 1.0000001, 1e-9, 0)
  RUN(k_dadd, 32, 1, out, cyc, 1e-9)
  RUN(k_shfl, 32, 1, out, cyc)
  RUN(k_shfl_tp, 32, 8, out, cyc)
  RUN(k_shfl_tp, 128, 8, out, cyc)
  RUN(k_shfl_tp, 256, 8, out, cyc)
  RUN(k_lds, 32, 1, out, cyc)
  RUN(k_rsqrt, 32, 1, out, cyc)
  RUN(k_rcp, 32, 1, out, cyc)
  RUN(k_dfma_tp, 32, 8, out, cyc, 1.0000001, 1e-9)
  RUN(k_dfma_tp, 128, 8, out, cyc, 1.0000001, 1e-9)
  RUN(k_dfma_tp, 256, 8, out, cyc, 1.0000001, 1e-9)
  RUN(k_dfma_tp, 512, 8, out, cyc, 1.0000001, 1e-9)
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
