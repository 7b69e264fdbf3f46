// umma_ts.cu -- known-answer test + timing of tcgen05.mma kind::tf32 with the A operand in
// TENSOR MEMORY (written with tcgen05.st, lane = row, column = k) and B from shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_ts umma_ts.cu && ./umma_ts
// D[128 x N] = A[128 x K] * B[N x K]^T, K = 32 (four MMAs of K = 8).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

constexpr int M = 128, KT = 32;
constexpr int LBO = 128;
constexpr int SBO = (KT / 4) * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(LBO >> 4) << 16) | ((uint64_t)(SBO >> 4) << 32) | ((uint64_t)1 << 46);
}
__device__ __forceinline__ int tile_off(int row, int k) { return (row >> 3) * SBO + (k >> 2) * LBO + (row & 7) * 16 + (k & 3) * 4; }

template <int N>
__global__ void __launch_bounds__(128, 1) k_test(const float* A, const float* B, float* D, long long* cyc, int rep) {
  __shared__ __align__(128) unsigned char sB[256 * KT * 4];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * KT; i += 128) *reinterpret_cast<float*>(sB + tile_off(i / KT, i % KT)) = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t colA = 256;  // A lives in columns [256, 256 + KT), D in [0, N)
  // A -> TMEM: thread = row, 8 columns per store
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < KT; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(A[row * KT + c0 + j]);
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + colA + c0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]), "r"(v[1]),
                 "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  auto mma_ts = [&](uint32_t d, uint32_t a, uint64_t db, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(db), "r"(idesc), "r"(acc)
        : "memory");
  };
  auto wait = [&](int b, uint32_t par) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; ++it)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(done) : "r"(smem_u32(&s_bar[b])), "r"(par) : "memory");
    if (!done) __trap();
  };
  if (tid == 0) {
    for (int ks = 0; ks < KT / 8; ++ks) mma_ts(tmem, tmem + colA + ks * 8, make_desc(smem_u32(sB) + ks * 2 * LBO), ks > 0);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[0])) : "memory");
  }
  wait(0, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // timing: rep x (KT/8) MMAs into 2 alternating accumulators (cols 0.. and 128..)
  if (tid == 0) {
    const long long t0 = clock64();
    for (int r = 0; r < rep; ++r)
      for (int ks = 0; ks < KT / 8; ++ks)
        mma_ts(tmem + (r & 1) * 128, tmem + colA + ks * 8, make_desc(smem_u32(sB) + ks * 2 * LBO), 1u);
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar[1])) : "memory");
    wait(1, 0);
    cyc[0] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

template <int N>
int run() {
  float *hA = new float[M * KT], *hB = new float[N * KT], *hD = new float[M * N];
  srand(1);
  for (int i = 0; i < M * KT; ++i) hA[i] = (float)((rand() % 33) - 16) / 8.0f;
  for (int i = 0; i < N * KT; ++i) hB[i] = (float)((rand() % 33) - 16) / 16.0f;
  float *dA, *dB, *dD; long long* dc;
  cudaMalloc(&dA, M * KT * 4); cudaMalloc(&dB, N * KT * 4); cudaMalloc(&dD, M * N * 4); cudaMalloc(&dc, 8);
  cudaMemcpy(dA, hA, M * KT * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * KT * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, M * N * 4);
  const int rep = 256;
  k_test<N><<<1, 128>>>(dA, dB, dD, dc, rep);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  long long c; cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < KT; ++k) ref += (double)hA[m * KT + k] * hB[n * KT + k];
      const double err = fabs(ref - hD[m * N + n]);
      if (!(err <= 1e-5)) { if (bad < 3) printf("mismatch m=%d n=%d got %g want %g\n", m, n, hD[m * N + n], ref); ++bad; }
      if (err > maxerr) maxerr = err;
    }
  printf("TS tf32 M=128 N=%3d: max abs err %.3g, mismatches %d; %.1f cycles/MMA (K=8) over %d MMAs (math floor %d)\n", N, maxerr,
         bad, (double)c / (rep * KT / 8), rep * KT / 8, 128 * N / 256);
  return bad != 0;
}

int main() { int b = run<16>(); b |= run<32>(); b |= run<64>(); b |= run<128>(); return b; }
