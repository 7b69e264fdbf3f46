// dmma.cu -- throughput of the FP64 tensor-core MMA (mma.sync m8n8k4 f64) on one SM and
// whether it shares a pipe with DFMA.  One CTA per SM, W warps, each warp runs ILP
// independent accumulator chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma dmma.cu && ./dmma
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int ILP, int DF>
__global__ void k_dmma(double* out, long long* cyc, int n) {
  double c[ILP > 0 ? ILP : 1][2], f[DF > 0 ? DF : 1];
  const double a = out[threadIdx.x & 31], b = out[32 + (threadIdx.x & 31)];
#pragma unroll
  for (int i = 0; i < (ILP > 0 ? ILP : 1); ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < (DF > 0 ? DF : 1); ++i) f[i] = a;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
    for (int i = 0; i < DF; ++i) f[i] = fma(f[i], b, a);
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < (DF > 0 ? DF : 1); ++i) s += f[i];
  out[64 + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int ILP, int DF>
void run(int warps, double* d_out, long long* d_cyc, const char* label) {
  const int n = 4096;
  k_dmma<ILP, DF><<<1, warps * 32>>>(d_out, d_cyc, n);
  k_dmma<ILP, DF><<<1, warps * 32>>>(d_out, d_cyc, n);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  const double mm = (double)n * ILP * warps, ff = (double)n * DF * warps;
  printf("%-28s warps %2d  cycles %9lld  DMMA/clk/SM %.3f (%.0f flop/clk)  DFMA warp-inst/clk/SM %.3f\n", label, warps, c,
         mm / c, mm * 512 / c, ff / c);
}

int main() {
  double* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 8 * 4096); cudaMalloc(&d_cyc, 8 * 256);
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 + 1e-3 * i;
  cudaMemcpy(d_out, h, sizeof h, cudaMemcpyHostToDevice);
  for (int w : {1, 4, 8, 16, 32}) run<8, 0>(w, d_out, d_cyc, "DMMA only, ILP 8");
  for (int w : {4, 16}) run<1, 0>(w, d_out, d_cyc, "DMMA only, ILP 1 (latency)");
  for (int w : {4, 16, 32}) run<0, 8>(w, d_out, d_cyc, "DFMA only, ILP 8");
  for (int w : {4, 16, 32}) run<8, 8>(w, d_out, d_cyc, "DMMA x8 + DFMA x8");
  for (int w : {4, 16, 32}) run<8, 32>(w, d_out, d_cyc, "DMMA x8 + DFMA x32");
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
