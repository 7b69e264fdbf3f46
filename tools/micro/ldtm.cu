// ldtm.cu -- throughput / latency of tcgen05.ld.32x32b.x8 (8 consecutive 32-bit columns of the
// thread's own TMEM lane, warp-uniform column) from W warps of one CTA, next to the same
// 32 B per lane gathered from shared memory with four LDS.64.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm ldtm.cu && ./ldtm
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int ILP>
__global__ void __launch_bounds__(1024, 1) k_ldtm(long long* cyc, double* out, int n) {
  __shared__ uint32_t s_t;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_t)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_t + ((uint32_t)((warp & 3) * 32) << 16);
  // fill this warp's lane quarter: column c of lane l holds l + c
  if (warp < 4)
    for (int c = 0; c < 512; c += 8) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) v[j] = __float_as_uint((float)(lane + c + j));
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem + c), "r"(v[0]),
                   "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
    }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float acc = 0.f;
  uint32_t col = (warp * 8) & 255;
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
    uint32_t v[ILP][8];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[i][0]), "=r"(v[i][1]), "=r"(v[i][2]), "=r"(v[i][3]), "=r"(v[i][4]), "=r"(v[i][5]),
                     "=r"(v[i][6]), "=r"(v[i][7])
                   : "r"(tmem + ((col + i * 24) & 255)));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += __uint_as_float(v[i][0]) + __uint_as_float(v[i][7]);
    col = (col + 8 + (__float_as_uint(acc) & 8)) & 255;  // data-dependent next column
  }
  const long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(s_t));
}

template <int ILP>
__global__ void __launch_bounds__(1024, 1) k_lds(long long* cyc, double* out, int n) {
  extern __shared__ double tab[];  // [256][32]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 256 * 32; i += blockDim.x) tab[i] = (double)i;
  __syncthreads();
  double acc = 0.0;
  int row = warp & 127;
  const long long t0 = clock64();
  for (int it = 0; it < n; ++it) {
    double v[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      const double* p = tab + ((row + i * 12) & 127) * 32 + lane;
      v[i][0] = p[0]; v[i][1] = p[32]; v[i][2] = p[64]; v[i][3] = p[96];
    }
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc += v[i][0] + v[i][3] + v[i][1] * v[i][2];
    row = (row + 4 + ((int)acc & 4)) & 127;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  long long* d_c; double* d_o;
  cudaMalloc(&d_c, 64); cudaMalloc(&d_o, 8 * 1024);
  const int n = 4096;
  cudaFuncSetAttribute(k_lds<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaFuncSetAttribute(k_lds<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int w : {1, 4, 8, 16, 32}) {
    long long c1, c4, s1, s4;
    k_ldtm<1><<<1, 32 * w>>>(d_c, d_o, n); cudaDeviceSynchronize(); cudaMemcpy(&c1, d_c, 8, cudaMemcpyDeviceToHost);
    k_ldtm<4><<<1, 32 * w>>>(d_c, d_o, n); cudaDeviceSynchronize(); cudaMemcpy(&c4, d_c, 8, cudaMemcpyDeviceToHost);
    k_lds<1><<<1, 32 * w, 65536>>>(d_c, d_o, n); cudaDeviceSynchronize(); cudaMemcpy(&s1, d_c, 8, cudaMemcpyDeviceToHost);
    k_lds<4><<<1, 32 * w, 65536>>>(d_c, d_o, n); cudaDeviceSynchronize(); cudaMemcpy(&s4, d_c, 8, cudaMemcpyDeviceToHost);
    printf("warps %2d | LDTM.x8: dependent %.1f cyc/step; x4 per step %.1f cyc -> %.3f loads/clk/SM (%.0f B/clk)"
           " | 4xLDS.64: dependent %.1f; x4 %.1f -> %.3f gathers/clk/SM (%.0f B/clk)\n",
           w, (double)c1 / n, (double)c4 / n, 4.0 * w * n / c4, 4.0 * w * n / c4 * 1024, (double)s1 / n, (double)s4 / n,
           4.0 * w * n / s4, 4.0 * w * n / s4 * 1024);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
