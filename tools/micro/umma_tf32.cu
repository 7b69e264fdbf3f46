// umma_tf32.cu -- known-answer test of the tcgen05 pieces the tensor-core XtWX route uses:
// shared-memory matrix descriptors (K-major, no swizzle), the kind::tf32 instruction
// descriptor, TMEM allocation, tcgen05.commit -> mbarrier, tcgen05.ld 32x32b.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_tf32 umma_tf32.cu && ./umma_tf32
// D[128 x N] = A[128 x K] * B[N x K]^T with K = 16 (two MMAs of K = 8), 3xTF32 split optional.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

constexpr int M = 128, N = 80, KT = 16;
constexpr int LBO = 128;             // bytes between the two 16-byte K chunks of one MMA
constexpr int SBO = (KT / 4) * 128;  // bytes between 8-row groups

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((LBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((SBO >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, layout type 0 (no swizzle)
}

__device__ __forceinline__ int tile_off(int row, int k) {  // byte offset of element (row, k)
  return (row >> 3) * SBO + (k >> 2) * LBO + (row & 7) * 16 + (k & 3) * 4;
}

__global__ void __launch_bounds__(128, 1) k_test(const float* A, const float* B, float* D) {
  __shared__ __align__(128) unsigned char sA[M * KT * 4];
  __shared__ __align__(128) unsigned char sB[N * KT * 4];
  __shared__ __align__(8) uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < M * KT; i += 128) {
    const int r = i / KT, k = i % KT;
    *reinterpret_cast<float*>(sA + tile_off(r, k)) = A[i];
  }
  for (int i = tid; i < N * KT; i += 128) {
    const int r = i / KT, k = i % KT;
    *reinterpret_cast<float*>(sB + tile_off(r, k)) = B[i];
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy (MMA)
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  if (tid == 0) {
    // instruction descriptor: D f32, A/B tf32, both K-major, N >> 3, M >> 4
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    for (int ks = 0; ks < KT / 8; ++ks) {
      const uint64_t da = make_desc(smem_u32(sA) + ks * 2 * LBO);
      const uint64_t db = make_desc(smem_u32(sB) + ks * 2 * LBO);
      const uint32_t acc = ks > 0;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar))
                 : "memory");
  }
  // bounded wait on phase 0
  uint32_t done = 0;
  for (int it = 0; it < (1 << 22) && !done; ++it)
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(smem_u32(&s_bar)), "r"(0u)
        : "memory");
  if (!done) __trap();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // warp w reads TMEM lanes 32w .. 32w+31; thread = row
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem));
}


// ---- timing: REP back-to-back MMAs (same operands, one accumulator) for several N ----
template <int NN>
__global__ void __launch_bounds__(128, 1) k_time(long long* out, int rep) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t s_t;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 * KT * 4 + 256 * KT * 4) / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 1.0f;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&s_t)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_t;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NN >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint64_t da = make_desc(smem_u32(sm)), db = make_desc(smem_u32(sm) + 128 * KT * 4);
    t0 = clock64();
    for (int r = 0; r < rep; ++r)
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
          "l"(da), "l"(db), "r"(idesc), "r"(1u)
          : "memory");
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                 : "memory");
    uint32_t done = 0;
    while (!done)
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
          : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    t1 = clock64();
    out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

template <int NN>
void time_n() {
  long long* d; cudaMalloc(&d, 8);
  const int smem = 128 * KT * 4 + 256 * KT * 4;
  cudaFuncSetAttribute(k_time<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep : {64, 512}) {
    k_time<NN><<<1, 128, smem>>>(d, rep);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    printf("SS tf32 M=128 N=%3d K=8: %4d MMAs in %7lld cycles = %.1f cycles/MMA (math floor %d)\n", NN, rep, c,
           (double)c / rep, 128 * NN / 256);
  }
  cudaFree(d);
}

int main() {
  float *hA = new float[M * KT], *hB = new float[N * KT], *hD = new float[M * N];
  srand(1);
  for (int i = 0; i < M * KT; ++i) hA[i] = (float)((rand() % 33) - 16) / 8.0f;  // exact in tf32
  for (int i = 0; i < N * KT; ++i) hB[i] = (float)((rand() % 33) - 16) / 16.0f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, M * KT * 4); cudaMalloc(&dB, N * KT * 4); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA, M * KT * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB, N * KT * 4, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xFF, M * N * 4);
  k_test<<<1, 128>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(hD, dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0; int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < KT; ++k) ref += (double)hA[m * KT + k] * hB[n * KT + k];
      const double err = fabs(ref - hD[m * N + n]);
      if (!(err <= 1e-5)) { if (bad < 5) printf("mismatch m=%d n=%d got %g want %g\n", m, n, hD[m * N + n], ref); ++bad; }
      if (err > maxerr) maxerr = err;
    }
  time_n<16>(); time_n<32>(); time_n<80>(); time_n<128>(); time_n<256>();
  printf("umma_tf32: max abs err %.3g, mismatches %d of %d\n", maxerr, bad, M * N);
  return bad != 0;
}
