// Latency micro-benchmarks for the building blocks of k_main (one warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
#include "../../liesel_ptm_b200/csrc/ptm_kernels.cuh"
using namespace ptm;

__global__ void k_dfma(double* out, long long* cyc, int n) {
  double a = out[0], b = out[1];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = fma(a, b, b); a = fma(a, b, b); a = fma(a, b, b); a = fma(a, b, b); }
  long long t1 = clock64();
  out[2] = a; if (threadIdx.x == 0) cyc[0] = (t1 - t0);
}
__global__ void k_lds(double* out, long long* cyc, int n) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i * 17 + 5) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { p = idx[p]; p = idx[p]; p = idx[p]; p = idx[p]; }
  long long t1 = clock64();
  out[3] = p; if (threadIdx.x == 0) cyc[1] = (t1 - t0);
}
__global__ void k_dispatch(double* out, long long* cyc, const int* jj, int n) {
  __shared__ int sj[1024];
  __shared__ double sb[1024 * 4];
  for (int i = threadIdx.x; i < 1024; i += 32) { sj[i] = jj[i]; sb[i*4]=0.1; sb[i*4+1]=0.2; sb[i*4+2]=0.3; sb[i*4+3]=0.4; }
  __syncthreads();
  double acc[20];
#pragma unroll
  for (int j = 0; j < 20; ++j) acc[j] = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    double b0, b1, b2, b3; load4(sb + i * 4, b0, b1, b2, b3);
    Dispatch<double, 20>::scatter(acc, sj[i], b0, b1, b2, b3, 1.5);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 20; ++j) s += acc[j];
  out[4] = s; if (threadIdx.x == 0) cyc[2] = (t1 - t0);
}
__global__ void k_gather(double* out, long long* cyc, const int* jj, int n) {
  __shared__ int sj[1024];
  __shared__ double sb[1024 * 4];
  for (int i = threadIdx.x; i < 1024; i += 32) { sj[i] = jj[i]; sb[i*4]=0.1; sb[i*4+1]=0.2; sb[i*4+2]=0.3; sb[i*4+3]=0.4; }
  __syncthreads();
  double g[20];
#pragma unroll
  for (int j = 0; j < 20; ++j) g[j] = out[j + 8];
  double v = 0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    double b0, b1, b2, b3; load4(sb + i * 4, b0, b1, b2, b3);
    v = Dispatch<double, 20>::gather(g, sj[i], b0, b1, b2, b3, v);
  }
  long long t1 = clock64();
  out[5] = v; if (threadIdx.x == 0) cyc[3] = (t1 - t0);
}
__global__ void k_exp(double* out, long long* cyc, int n) {
  double a = out[0];
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = exp(-a); }
  long long t1 = clock64();
  out[6] = a; if (threadIdx.x == 0) cyc[4] = (t1 - t0);
}
__global__ void k_div(double* out, long long* cyc, int n) {
  double a = out[0] + 1.0;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) { a = 1.0 / a + 0.5; }
  long long t1 = clock64();
  out[7] = a; if (threadIdx.x == 0) cyc[5] = (t1 - t0);
}
__global__ void k_dispatch_pf(double* out, long long* cyc, const int* jj, int n) {
  // same as k_dispatch with the operands of step i+1 loaded before the jump of step i
  __shared__ int sj[1024];
  __shared__ double sb[1024 * 4];
  for (int i = threadIdx.x; i < 1024; i += 32) { sj[i] = jj[i]; sb[i*4]=0.1; sb[i*4+1]=0.2; sb[i*4+2]=0.3; sb[i*4+3]=0.4; }
  __syncthreads();
  double acc[20];
#pragma unroll
  for (int j = 0; j < 20; ++j) acc[j] = 0;
  long long t0 = clock64();
  double b0, b1, b2, b3; load4(sb, b0, b1, b2, b3);
  int j0 = sj[0];
  for (int i = 0; i < n; ++i) {
    double c0, c1, c2, c3; load4(sb + ((i + 1) & 1023) * 4, c0, c1, c2, c3);
    const int j1 = sj[(i + 1) & 1023];
    Dispatch<double, 20>::scatter(acc, j0, b0, b1, b2, b3, 1.5);
    b0 = c0; b1 = c1; b2 = c2; b3 = c3; j0 = j1;
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < 20; ++j) s += acc[j];
  out[9] = s; if (threadIdx.x == 0) cyc[6] = (t1 - t0);
}
__global__ void k_fwdbwd_pf(double* out, long long* cyc, const int* jj, int n) {
  // forward gather (registers g) + backward scatter (registers acc) per step, prefetched operands
  __shared__ int sj[1024];
  __shared__ double sb[1024 * 4];
  for (int i = threadIdx.x; i < 1024; i += 32) { sj[i] = jj[i]; sb[i*4]=0.1; sb[i*4+1]=0.2; sb[i*4+2]=0.3; sb[i*4+3]=0.4; }
  __syncthreads();
  double acc[20], g[20];
#pragma unroll
  for (int j = 0; j < 20; ++j) { acc[j] = 0; g[j] = out[j + 8]; }
  long long t0 = clock64();
  double b0, b1, b2, b3; load4(sb, b0, b1, b2, b3);
  int j0 = sj[0];
  double v = 0;
  for (int i = 0; i < n; ++i) {
    double c0, c1, c2, c3; load4(sb + ((i + 1) & 1023) * 4, c0, c1, c2, c3);
    const int j1 = sj[(i + 1) & 1023];
    const double f = Dispatch<double, 20>::gather(g, j0, b0, b1, b2, b3, 0.0);
    v += f;
    Dispatch<double, 20>::scatter(acc, sj[(i + 512) & 1023], b0, b1, b2, b3, 1.5);
    b0 = c0; b1 = c1; b2 = c2; b3 = c3; j0 = j1;
  }
  long long t1 = clock64();
  double s = v;
#pragma unroll
  for (int j = 0; j < 20; ++j) s += acc[j];
  out[10] = s; if (threadIdx.x == 0) cyc[7] = (t1 - t0);
}
int main() {
  double* out; long long* cyc; int* jj;
  cudaMallocManaged(&out, 64 * 8); cudaMallocManaged(&cyc, 128); cudaMallocManaged(&jj, 1024 * 4);
  for (int i = 0; i < 64; ++i) out[i] = 0.3 + 0.01 * i;
  for (int i = 0; i < 1024; ++i) jj[i] = (i * 7919) % 17;
  int n = 1024;
  for (int rep = 0; rep < 2; ++rep) {
    k_dfma<<<1, 32>>>(out, cyc, n); k_lds<<<1, 32>>>(out, cyc, n); k_dispatch<<<1, 32>>>(out, cyc, jj, n);
    k_gather<<<1, 32>>>(out, cyc, jj, n); k_exp<<<1, 32>>>(out, cyc, n); k_div<<<1, 32>>>(out, cyc, n);
    k_dispatch_pf<<<1, 32>>>(out, cyc, jj, n); k_fwdbwd_pf<<<1, 32>>>(out, cyc, jj, n);
    cudaDeviceSynchronize();
  }
  printf("dependent DFMA        : %.1f cycles\n", cyc[0] / (4.0 * n));
  printf("dependent LDS (32-bit): %.1f cycles\n", cyc[1] / (4.0 * n));
  printf("bwd dispatch step     : %.1f cycles (load4 + LDS jj + brx + 4 DFMA independent)\n", cyc[2] / (double)n);
  printf("fwd dispatch step     : %.1f cycles (load4 + LDS jj + brx + 4 DFMA chained)\n", cyc[3] / (double)n);
  printf("bwd dispatch, prefetch : %.1f cycles\n", cyc[6] / (double)n);
  printf("fwd + bwd, prefetch    : %.1f cycles (two jump tables per step)\n", cyc[7] / (double)n);
  printf("dependent exp()       : %.1f cycles\n", cyc[4] / (double)n);
  printf("dependent 1/x + add   : %.1f cycles\n", cyc[5] / (double)n);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
