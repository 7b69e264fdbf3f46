// ptm_main_tc.cuh -- tensor-core route of the fused logp(+gradient) sweep for FLOAT problems.
//
// The two contractions of the sweep have a chain-independent operand (dist.py:718-740 over
// gam/var.py:157-162 and its transpose in reverse mode):
//
//   forward   eta_q[chain, obs]  = sum_k Gamma[chain, k] * B[obs, k]      q = mu, sigma
//   backward  dGamma[chain, k]  += sum_obs G_q[chain, obs] * B[obs, k]    G_q = dl/deta_q
//
// with k = (smooth term, B-spline coefficient) and B the banded design expanded to its dense
// row (4 non-zeros per term and observation).  Both run on tcgen05.mma kind::tf32 with the
// 3xTF32 split (hi*hi + hi*lo + lo*hi, hi = cvt.rna.tf32, lo = x - hi), M = 128 chains:
//
//   k_main_tc  (forward + point-wise part):  A = Gamma hi / lo in TENSOR MEMORY (written once per
//       work item with tcgen05.st, lane = chain), B = the expanded design tile [NT obs x K] in
//       shared memory (core-matrix layout, written by the stager warps: only the four entries
//       per term and observation change from tile to tile), D = eta_mu, eta_sigma [128 x NT]
//       in TMEM.  16 eval warps (4 per chain quarter) read their eta values with tcgen05.ld,
//       release the accumulators, and run the point-wise part of k_main (transformation,
//       log-Jacobian, declared derivatives, gradient w.r.t. the B-spline coefficients of h);
//       dl/deta goes to HBM as G[obs][q][chain] (8 bytes per evaluation).
//   k_grad_tc  (backward):  A = G hi / lo in TMEM (loader warps: coalesced read, split,
//       tcgen05.st), B = the expanded design transposed [K coefs x 32 obs] in shared memory,
//       D = dGamma [128 chains x K] in TMEM, promoted to the double partials of k_post every
//       `flush` tiles.
//
// TMEM budget (512 columns): 2 K (Gamma hi, lo) + 2 NT (eta)  and  2 K16 + 256 (G tiles), so the
// single-launch form takes models with K = sum of bases <= 224 (c3: 5 + 5 terms x 20 -> 208,
// NT = 48 by shared memory).  Larger models run in several passes (make_plan, ptm_capi.cu):
// groups of whole terms of one predictor (<= 224 coefficients) go through the ETAONLY
// instantiation, which writes / accumulates its partial predictor to E[q] in HBM; the main launch
// carries the remaining terms and adds the external parts (ext_mask); k_grad_tc runs once per
// group.  tools/micro/umma_ts.cu is the known-answer test of the TMEM-operand MMA form used here.
#pragma once
#include <cstdint>
#include <type_traits>

#include "ptm_kernels.cuh"
#include "ptm_xtwx_tc.cuh"

// the eta-only instantiation of k_main_tc leaves the point-wise accumulators unreferenced
#pragma nv_diag_suppress 177

namespace ptm {

constexpr int kMtcEvalWarps = 16;  // 4 chain quarters x 4 observation groups
constexpr int kMtcMaxK = 224;
constexpr int kGtcKT = 32;         // observations per backward K tile
#ifndef PTM_GTC_HINT
#define PTM_GTC_HINT 100   // mbarrier try_wait suspend hint (ns) of k_grad_tc: every wait of that kernel sits
#define PTM_GTC_SLEEP 0    // on the serial chain copy -> split -> MMA, so its warps poll tightly
#endif
constexpr int kGtcStages = 3;      // G tiles in flight (bulk async copies into shared memory)
constexpr int kGtcTileBytes = kGtcKT * 2 * 128 * 4;  // one G tile: [32 obs][2 predictors][128 chains] floats

struct MainTcArgs {
  MainArgs<float> a;
  float* G;            // [CB][Gld][2][128] dl/deta_mu, dl/deta_sigma of (chain block, observation), chain fastest
  long long Gld;       // observations per chain block in G (shard length padded by one backward tile)
  int CB, Cpad;        // chain blocks of 128
  int Kt;              // forward K extent (both predictor blocks, each padded to 8)
  int Kq[2], kbase[2]; // forward K extent / first column of the mu and sigma blocks
  int NT;              // observations per forward tile (48 or 64)
  int NSW;             // stager warps of the forward kernel
  int NSWb;            // stager warps of the backward kernel
  int kcol[kMaxTerms]; // first forward K column of each term
  unsigned char kterm[kMtcMaxK], kj[kMtcMaxK];  // forward K column -> (term, coefficient); 255 = padding
  // backward
  int Nq[2], nbase[2]; // D columns / first D column of the predictor blocks (padded to 16)
  int ncol[kMaxTerms]; // first D column of each term
  short dslot[2 * kMtcMaxK + 32];  // D column -> partial slot (-1 = padding)
  int flush;           // backward tiles between promotions to double
  int mode;            // kModeLogp / kModeGrad
  // multi-pass form for models whose coefficients do not fit tensor memory at once: eta-only
  // launches (ETAONLY kernel) carry groups of <= 224 coefficients of ONE predictor and write /
  // accumulate its partial predictor to HBM; the main launch carries the remaining terms on the
  // tensor cores and adds the external parts
  float* E[2];         // [CB][Gld][128] partial eta_mu / eta_sigma of the eta-only launches
  int e_acc;           // eta-only launch: add to E instead of overwriting it
  int ext_mask;        // main launch: bit q = add E[q] to predictor q
};

__device__ __forceinline__ void mtc_mma_ts(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a_tmem), "l"(db),
      "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint64_t mtc_desc(uint32_t addr, int sbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(128 >> 4) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);
}
__device__ __forceinline__ void mtc_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mtc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mtc_st8(uint32_t addr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void mtc_ld4(uint32_t addr, uint32_t (&v)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
               : "r"(addr));
}
// Wait on an mbarrier phase with a back-off between polls: a spinning warp would take issue
// slots from the eval warps of its scheduler (bounded like tc_wait: a protocol bug traps).
template <int HINT_NS = 2000, int SLEEP_NS = 200>
__device__ __forceinline__ void mtc_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 24); ++it) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"((uint32_t)HINT_NS)
        : "memory");
    if (done) return;
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
  }
  __trap();
}
// exp / reciprocal / log of the float point-wise part: one MUFU each (ex2.approx, rcp.approx,
// lg2.approx: <= 2 ulp; the float parity bar is 1e-5)
__device__ __forceinline__ float mtc_exp(float x) {
  // exp(x) = 2^t (1 + e ln 2): t = fl(x log2 e), e = the exactly recovered rounding error of
  // that product plus the low word of log2 e -- the argument of the approximate ex2 is then
  // exact to ~2^-31 instead of |x| 2^-24, which removes the systematic part of the error
  const float t = x * 1.4426950216293335f;
  float e = fmaf(x, 1.4426950216293335f, -t);
  e = fmaf(x, 1.9259629911266175e-8f, e);
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t));
  return fmaf(y, e * 0.69314718055994530942f, y);
}
__device__ __forceinline__ float mtc_rcp(float x) {
  // (a Newton step on top changes nothing measurable: the float gradient error of this route
  //  sits in the f32 accumulation of G'B on the tensor cores, tools/mtc_acc.py)
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float mtc_log(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.69314718055994530942f;
}
// packed float32x2 arithmetic (Blackwell FFMA2 / FADD2 / FMUL2: two observations of a chain per
// instruction in the point-wise part)
__device__ __forceinline__ float2 p2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 p1(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 padd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 pmul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 psub(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// ---------------------------------------------------------------------------------------
// forward + point-wise part
// ---------------------------------------------------------------------------------------
template <int NT, bool GRAD, int MAXT, bool ETAONLY = false>
__global__ void __launch_bounds__(MAXT, 1) k_main_tc(const MainTcArgs g) {
  const MainArgs<float>& a = g.a;
  constexpr int OPW = NT / 4;  // observations per eval warp and tile
  extern __shared__ __align__(128) unsigned char smem_mtc[];
  __shared__ __align__(8) uint64_t s_bar[5];  // [0] eta ready (all MMAs of the tile), [1] eta consumed,
                                              // [2], [3] mu / sigma block of the B tile filled, [4] mu MMAs done
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, KK = a.KK, Kt = g.Kt, NSW = g.NSW;
  const int NCC = KK * 5 + 6, NGV = KK + 4;
  const int sbo = (Kt >> 2) * 128, bbytes = NT * Kt * 4;
  unsigned char* s_B = smem_mtc;                                            // [hi, lo][NT x Kt]
  float* s_cc = reinterpret_cast<float*>(s_B + 2 * bbytes);                 // [NCC + 4][128]
  float* s_gv = s_cc + (NCC + 4) * 128;                                     // [16][NGV][32]
  float* s_kn = s_gv + (GRAD ? kMtcEvalWarps * NGV * 32 : 0);               // [T][4][kKnotStride]
  // [16][kRegSlots][32] doubles of the item-end reduction, ALIASED onto the per-warp eta buffers
  // [16][2 OPW][32] + [32]: the reduction scratch is written after an item's last tile and read
  // before the next item's first one (eval warps only, between their named barriers)
  double* s_red = reinterpret_cast<double*>(s_kn + T * 4 * kKnotStride + ((T * 4 * kKnotStride) & 1));
  float* s_eta = reinterpret_cast<float*>(s_red);

  const bool is_eval = warp < kMtcEvalWarps;
  const int sidx = warp - kMtcEvalWarps;
  const bool is_stager = sidx >= 0 && sidx < NSW;
  const bool is_mma = sidx == NSW;
  const int q4 = warp & 3, sub = (warp >> 2) & 3;

  for (int t = warp; t < T; t += blockDim.x >> 5) {
    float* kn = s_kn + t * 4 * kKnotStride;
    const float* gk = a.knots + t * kKnotStride;
    const int nbt = a.nb[t];
    for (int i = lane; i < nbt + 4; i += 32) {
      kn[i] = gk[i];
      kn[kKnotStride + i] = i + 1 < nbt + 4 ? 1.f / (gk[i + 1] - gk[i]) : 0.f;
      kn[2 * kKnotStride + i] = i + 2 < nbt + 4 ? 1.f / (gk[i + 2] - gk[i]) : 0.f;
      kn[3 * kKnotStride + i] = i + 3 < nbt + 4 ? 1.f / (gk[i + 3] - gk[i]) : 0.f;
    }
  }
  // the design tile starts as zeros; afterwards only the entries a tile wrote are cleared
  for (int i = tid; i < (2 * bbytes) >> 4; i += blockDim.x) reinterpret_cast<float4*>(s_B)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[1])), "r"(kMtcEvalWarps));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[2])), "r"(NSW));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[3])), "r"(NSW));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[4])));
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
  const uint32_t barE = smem_u32(&s_bar[0]), barC = smem_u32(&s_bar[1]), barB = smem_u32(&s_bar[2]);  // barB + 8: sigma block
  const uint32_t barM0 = smem_u32(&s_bar[4]);
  const uint32_t etaCol = (uint32_t)(2 * Kt);
  const TrafoScal<float> ts = a.ts;
  uint32_t tg = 0;  // tiles this CTA has gone through (phase parity of all three barriers)

  // stager state: K column each (owned term, pass) wrote last, -1 = nothing to clear
  int prev_k[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) prev_k[i][0] = prev_k[i][1] = -1;

  // The roles run their own copies of the work-item loop (they meet at the CTA barriers only).
  auto run_items = [&](auto role_c) {
  constexpr int ROLE = decltype(role_c)::value;  // 0 = eval warps, 1 = stager / MMA / idle warps
  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int cb = item / a.M, mchunk = item - cb * a.M;
    const long long o_begin = a.obs_begin + (long long)mchunk * a.chunk_len;
    long long o_end = o_begin + a.chunk_len;
    if (o_end > a.obs_end) o_end = a.obs_end;
    const long long len = o_end > o_begin ? o_end - o_begin : 0;
    const int ntiles = (int)((len + NT - 1) / NT);
    int chain = cb * 128 + q4 * 32 + lane;
    if (chain >= a.C) chain = a.C - 1;
    const int cg = cb * 4 + q4;
    (void)cg;

    float beta0 = 0.f, gamma0 = 0.f;
    if constexpr (ROLE == 0) {
      // Gamma hi / lo of the 128 chains -> TMEM (thread = chain, 8 columns per store)
      for (int c0 = sub * 8; c0 < Kt; c0 += 32) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = c0 + j;
          const int t = g.kterm[k];
          const float v = t != 255 ? a.Gamma[((long long)t * kMaxNbases + g.kj[k]) * a.C + chain] : 0.f;
          const float vh = tf32_round(v);
          hi[j] = __float_as_uint(vh);
          lo[j] = __float_as_uint(tf32_round(v - vh));
        }
        const uint32_t addr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)c0;
        mtc_st8(addr, hi);
        mtc_st8(addr + (uint32_t)Kt, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      for (int s = sub; s < NCC; s += 4) s_cc[s * 128 + q4 * 32 + lane] = a.cc[(long long)s * a.C + chain];
      if (sub == 0) {
        ChainBoundary<float> cbd;
        make_boundary(ts, a.cc[(long long)(KK * 5 + 0) * a.C + chain], a.cc[(long long)(KK * 5 + 1) * a.C + chain],
                      a.cc[(long long)(KK * 5 + 2) * a.C + chain], a.cc[(long long)(KK * 5 + 3) * a.C + chain], cbd);
        s_cc[(NCC + 0) * 128 + q4 * 32 + lane] = cbd.constL;
        s_cc[(NCC + 1) * 128 + q4 * 32 + lane] = cbd.constR;
        s_cc[(NCC + 2) * 128 + q4 * 32 + lane] = cbd.ztailL;
        s_cc[(NCC + 3) * 128 + q4 * 32 + lane] = cbd.ztailR;
      }
      beta0 = a.cc[(long long)(KK * 5 + 4) * a.C + chain];
      gamma0 = a.cc[(long long)(KK * 5 + 5) * a.C + chain];
      if (GRAD) {
        float* gvp = s_gv + (warp * NGV) * 32 + lane;
        for (int s = 0; s < NGV; ++s) gvp[s * 32] = 0.f;
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if constexpr (ROLE == 1) {
    if (is_stager) {
      // ---- design tile: clear what the previous tile wrote, write this tile's four entries
      //      per (term, observation), hi and lo halves
      // covariate values of the next tile are requested one tile ahead: the fill of the design
      // tile sits on the serial chain MMA(s) -> fill(s+1) -> MMA(s+1) of the single-buffered tile
      float xn[8][2];
      auto fetch_x = [&](int s) {
        const long long base = o_begin + (long long)s * NT;
#pragma unroll
        for (int ti = 0; ti < 8; ++ti) {
          const int t = sidx + ti * NSW;
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const long long gi = base + p * 32 + lane;
            xn[ti][p] = 0.f;
            if (t < T && p * 32 + lane < NT && gi < o_end) xn[ti][p] = a.X[(long long)t * a.n + gi];
          }
        }
      };
      if (ntiles > 0) fetch_x(0);
      for (int s = 0; s < ntiles; ++s, ++tg) {
        float xc[8][2];
#pragma unroll
        for (int ti = 0; ti < 8; ++ti) { xc[ti][0] = xn[ti][0]; xc[ti][1] = xn[ti][1]; }
        if (s + 1 < ntiles) fetch_x(s + 1);
        const long long base = o_begin + (long long)s * NT;
        // the two predictor blocks of the tile are filled, multiplied and released one after the
        // other: the MMAs of the mu block run while the sigma block is being written, and the mu
        // block of the next tile is rewritten while the sigma MMAs of this one still run
        // (the logp-only sweep is paced by this chain: 9.4 -> 7.6 ms on c3.  With the gradient only
        //  the fill side is split -- 14.7 -> 14.2 ms; rewriting the mu block of the next tile under
        //  the running sigma MMAs costs that form 1.8 ms, so there both blocks wait for barE)
        constexpr int NPASS = 2;
#pragma unroll 1
        for (int q = 0; q < NPASS; ++q) {
        if (tg > 0) mtc_wait((!GRAD && q == 0) ? barM0 : barE, (tg - 1) & 1);  // the MMAs that read the block have completed
#pragma unroll
        for (int ti = 0; ti < 8; ++ti) {
          const int t = sidx + ti * NSW;
          if (t >= T) break;
          if (g.kcol[t] < 0) continue;  // several-pass form: the term belongs to another launch
          if ((g.Kq[1] > 0 && g.kcol[t] >= g.kbase[1]) != (q == 1)) continue;  // the other block
          const float* kn = s_kn + t * 4 * kKnotStride;
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const int o = p * 32 + lane;
            if (o >= NT) continue;
            unsigned char* rowp = s_B + (o >> 3) * sbo + (o & 7) * 16;
            const int pk = prev_k[ti][p];
            if (pk >= 0) {
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const int k = pk + m;
                float* e = reinterpret_cast<float*>(rowp + (k >> 2) * 128 + (k & 3) * 4);
                e[0] = 0.f;
                *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(e) + bbytes) = 0.f;
              }
            }
            const long long gi = base + o;
            int nk = -1;
            if (gi < o_end) {
              float bb[4];
              const int jj = basis_window_fast(xc[ti][p], kn, kn + kKnotStride, kn + 2 * kKnotStride, kn + 3 * kKnotStride,
                                               a.nb[t], a.inv_dk[t], bb);
              nk = g.kcol[t] + jj;
#pragma unroll
              for (int m = 0; m < 4; ++m) {
                const int k = nk + m;
                const float vh = tf32_round(bb[m]);
                float* e = reinterpret_cast<float*>(rowp + (k >> 2) * 128 + (k & 3) * 4);
                e[0] = vh;
                *reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(e) + bbytes) = tf32_round(bb[m] - vh);
              }
            }
            prev_k[ti][p] = nk;
          }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          mtc_arrive(barB + 8 * q);
        }
        }
      }
    } else if (is_mma) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NT >> 3) << 17) | (8u << 24);
      for (int s = 0; s < ntiles; ++s, ++tg) {
        if (lane == 0) {
          if (tg > 0) mtc_wait(barC, (tg - 1) & 1);    // eval warps have read the previous eta
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            mtc_wait(barB + 8 * q, tg & 1);            // this block of the design tile is written
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem + etaCol + (uint32_t)(q * NT);
            const int nks = g.Kq[q] >> 3;
            for (int ks = 0; ks < nks; ++ks) {
              const int k0 = g.kbase[q] + ks * 8;
              const uint32_t a_hi = tmem + (uint32_t)k0, a_lo = a_hi + (uint32_t)Kt;
              const uint32_t baddr = smem_u32(s_B) + (uint32_t)((k0 >> 2) * 128);
              const uint64_t b_hi = mtc_desc(baddr, sbo), b_lo = mtc_desc(baddr + (uint32_t)bbytes, sbo);
              mtc_mma_ts(d, a_hi, b_hi, idesc, ks > 0 ? 1u : 0u);
              mtc_mma_ts(d, a_hi, b_lo, idesc, 1u);
              mtc_mma_ts(d, a_lo, b_hi, idesc, 1u);
            }
            mtc_commit(q == 0 ? barM0 : barE);
          }
        }
        __syncwarp();
      }
    } else {
      tg += (uint32_t)ntiles;  // idle warps keep the tile count in step
    }
    } else {
      // =========================== eval warp ==========================================
      float* gvp = s_gv + (warp * NGV) * 32 + lane;
      const float* ccl = s_cc + q4 * 32 + lane;
      (void)gvp; (void)ccl;
      double a_z2 = 0, a_ls = 0, a_gb0 = 0, a_gg0 = 0, a_gvl = 0, a_gdl = 0, a_gvr = 0, a_gdr = 0;
      double a_logd = 0;
      const float sixth = 1.f / 6.f;
      const bool ext0 = (g.ext_mask & 1) != 0, ext1 = (g.ext_mask & 2) != 0;
      const bool mma0 = g.Kq[0] > 0, mma1 = g.Kq[1] > 0, hasq0 = mma0 || ext0, hasq1 = mma1 || ext1;
      float* Gp = g.G + (long long)cb * g.Gld * 256 + q4 * 32 + lane;
      const long long eoff = (long long)cb * g.Gld * 128 + q4 * 32 + lane;
      const float* Ex0 = g.E[0] + eoff;
      const float* Ex1 = g.E[1] + eoff;
      float* Eout = g.E[mma0 ? 0 : 1] + eoff;  // eta-only launch: its one predictor
      const uint32_t addr = tmem + ((uint32_t)(q4 * 32) << 16) + etaCol + (uint32_t)(sub * OPW);
      constexpr int EBS = 2 * OPW * 32;  // floats of the eta buffer: [2 predictors][OPW][32 chains]
      float* eb = s_eta + warp * EBS;
      (void)eb;
      // eta of tile s2 -> the warp's shared buffer (+ the external parts of the multi-pass form), then
      // the accumulators go back to the MMA warp
      auto copy_eta = [&](int s2, float* ebuf, uint32_t par) {
        const long long b2 = o_begin + (long long)s2 * NT + sub * OPW;
        float ex0[OPW], ex1[OPW];
#pragma unroll
        for (int c = 0; c < OPW; ++c) { ex0[c] = 0.f; ex1[c] = 0.f; }
        if (ext0 || ext1) {  // in flight during the wait
#pragma unroll
          for (int c = 0; c < OPW; ++c)
            if (b2 + c < o_end) {
              if (ext0) ex0[c] = __ldcs(Ex0 + (b2 + c - a.obs_begin) * 128);
              if (ext1) ex1[c] = __ldcs(Ex1 + (b2 + c - a.obs_begin) * 128);
            }
        }
        mtc_wait(barE, par);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int c = 0; c < OPW; c += 4) {
          uint32_t v[4], w[4];
          mtc_ld4(addr + c, v);
          mtc_ld4(addr + NT + c, w);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            ebuf[(c + j) * 32 + lane] = (mma0 ? __uint_as_float(v[j]) : 0.f) + ex0[c + j];
            ebuf[(OPW + c + j) * 32 + lane] = (mma1 ? __uint_as_float(w[j]) : 0.f) + ex1[c + j];
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mtc_arrive(barC);
      };
      (void)copy_eta;
      // this warp's responses of a tile: one coalesced load, requested a tile ahead
      float yreg = 0.f;
      if (ntiles > 0 && lane < OPW && o_begin + sub * OPW + lane < o_end) yreg = a.y[o_begin + sub * OPW + lane];
      for (int s = 0; s < ntiles; ++s, ++tg) {
        const long long base = o_begin + (long long)s * NT + sub * OPW;
        if constexpr (ETAONLY) {
        mtc_wait(barE, tg & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          // eta of the one predictor this launch carries -> HBM, nothing else
          const uint32_t ecol = addr + (g.Kq[0] > 0 ? 0u : (uint32_t)NT);
#pragma unroll
          for (int c = 0; c < OPW; c += 4) {
            uint32_t v[4];
            mtc_ld4(ecol + c, v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (base + c + j < o_end) {
                float* ep = Eout + (base + c + j - a.obs_begin) * 128;
                *ep = g.e_acc ? *ep + __uint_as_float(v[j]) : __uint_as_float(v[j]);
              }
          }
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mtc_arrive(barC);
        }
        if constexpr (!ETAONLY) {
        copy_eta(s, eb, tg & 1);
        // the next tile's responses, requested a tile ahead
        float ynext = 0.f;
        if (s + 1 < ntiles && lane < OPW && base + NT + lane < o_end) ynext = a.y[base + NT + lane];
        float t_z2 = 0.f, t_ls = 0.f, t_gb0 = 0.f, t_gg0 = 0.f, t_ld = 0.f;
        const int cnt = (int)((o_end - base) < OPW ? (o_end - base) : OPW);
        auto eval_one = [&](int i) {
          const long long gi = base + i;
          {
            const float yv = __shfl_sync(0xffffffffu, yreg, i);  // lane i holds the response of observation i
            const float em = beta0 + (hasq0 ? eb[i * 32 + lane] : 0.f);
            const float es = gamma0 + (hasq1 ? eb[(OPW + i) * 32 + lane] : 0.f);
            const float einv = mtc_exp(-es);
            const float r = (yv - em) * einv;
            const float rc = fminf(fmaxf(r, ts.a), ts.b);  // NaN -> a
            const bool core = rc == r;
            const int code = core ? 2 : region_code(ts, r);
            CoreEval<float> ev;
            core_locate_fast(ts, a.grid, rc, ev);
            const float* cp = ccl + (ev.kk * 5) * 128;
            core_eval(ts, cp[0], cp[128], cp[256], cp[384], cp[512], ev);
            float z = ev.z, d = ev.d, dzdr = ev.d, dddr = ev.d2;
            float cfac = 0.f, qfac = 0.f;
            if (!core) {
              if (code == 1) {
                const float dL = ccl[(KK * 5 + 1) * 128], constL = ccl[(NCC + 0) * 128];
                const float poly = r * ts.a - 0.5f * r * r;
                const float t1 = r - poly * ts.inv_w;
                z = (ts.inv_w * poly + dL * t1) + constL;
                const float q = (ts.a - r) * ts.inv_w;
                d = (1.f - q) * dL + q;
                dzdr = d; dddr = (dL - 1.f) * ts.inv_w;
                const float poly0 = ts.a * ts.a - 0.5f * ts.a * ts.a;
                cfac = t1 - (ts.a - poly0 * ts.inv_w);
                qfac = 1.f - q;
              } else if (code == 3) {
                const float dR = ccl[(KK * 5 + 3) * 128], constR = ccl[(NCC + 1) * 128];
                const float poly = 0.5f * r * r - r * ts.b;
                const float t1 = r - poly * ts.inv_w;
                z = (ts.inv_w * poly + dR * t1) + constR;
                const float q = (r - ts.b) * ts.inv_w;
                d = (1.f - q) * dR + q;
                dzdr = d; dddr = (1.f - dR) * ts.inv_w;
                const float poly0 = 0.5f * ts.b * ts.b - ts.b * ts.b;
                cfac = t1 - (ts.b - poly0 * ts.inv_w);
                qfac = 1.f - q;
              } else if (code == 0) {
                z = ccl[(NCC + 2) * 128] - (ts.min_eps - r);
                d = 1.f; dzdr = 1.f; dddr = 0.f;
                cfac = cL_of(ts, ts.min_eps);
              } else {
                z = ccl[(NCC + 3) * 128] + (r - ts.max_eps);
                d = 1.f; dzdr = 1.f; dddr = 0.f;
                cfac = cR_of(ts, ts.max_eps);
              }
            }
            const float tiny = tiny_of<float>();
            const bool above = d > tiny;
            const float dcl = above ? d : tiny;
            t_z2 = fmaf(z, z, t_z2);
            t_ls += es;
            t_ld += mtc_log(dcl);
            if (!GRAD) {  // intercept statistics ride along with the logp-only sweep
              a_gb0 += (double)einv;
              a_gg0 += (double)r;
              a_gvl += (double)(r * r);
              a_gdl += (double)(einv * einv);
              a_gvr += (double)(r * einv);
            } else {
              const float lz = -z;
              const float ld = above ? mtc_rcp(dcl) : 0.f;
              const float gr = fmaf(lz, dzdr, ld * dddr);
              const float gmu = -gr * einv;
              const float gls = fmaf(-gr, r, -1.f);
              float* gp2 = Gp + (gi - a.obs_begin) * 256;
              gp2[0] = gmu;
              gp2[128] = gls;
              t_gb0 += gmu;
              t_gg0 += gls;
              float w[5];
              core_backward(ts, ev, core ? lz : 0.f, core ? ld : 0.f, w);
              const bool lastp = ev.kk + 1 >= KK;
              const float w4 = lastp ? 0.f : w[4];
              const float g3 = w[3] - w4;
              float* gp = gvp + ev.kk * 32;
              gp[0] += (w[0] - g3) * sixth + (w[2] - w[1]) * 0.5f;
              gp[32] += (4.f * w[0] - w4) * sixth - w[2] + g3 * 0.5f;
              gp[64] += w[0] * sixth + (w[1] + w[2] - g3 + w4) * 0.5f;
              gp[96] += g3 * sixth - w4 * 0.5f;
              gp[128] += w4 * sixth;
              if (!core) {
                const double dv = (double)fmaf(lz, cfac, ld * qfac);
                if (code <= 1) { a_gvl += (double)lz; a_gdl += dv; }
                else { a_gvr += (double)lz; a_gdr += dv; }
              }
            }
          }
        };
        int i0 = 0;
        if constexpr (GRAD) {
          // ---- two observations of the chain in lockstep, packed float32x2 arithmetic --------
          const float2 S1 = p1(ts.ratio), S3 = p1(3.f * ts.ratio), IDK = p1(ts.inv_dk);
          for (; i0 + 1 < cnt; i0 += 2) {
            const long long gi = base + i0;
            const float2 yv = p2(__shfl_sync(0xffffffffu, yreg, i0), __shfl_sync(0xffffffffu, yreg, i0 + 1));
            const float2 em = padd(p1(beta0), hasq0 ? p2(eb[i0 * 32 + lane], eb[(i0 + 1) * 32 + lane]) : p1(0.f));
            const float2 es = padd(p1(gamma0), hasq1 ? p2(eb[(OPW + i0) * 32 + lane], eb[(OPW + i0 + 1) * 32 + lane]) : p1(0.f));
            // einv = exp(-es): 2^t (1 + e ln 2), t = fl(-es log2 e), e = the recovered rounding error
            const float2 tt = pmul(es, p1(-1.4426950216293335f));
            float2 ee = pfma(es, p1(-1.4426950216293335f), p2(-tt.x, -tt.y));
            ee = pfma(es, p1(-1.9259629911266175e-8f), ee);
            float2 ex;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(tt.x));
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(tt.y));
            const float2 einv = pfma(ex, pmul(ee, p1(0.69314718055994530942f)), ex);
            const float2 r = pmul(psub(yv, em), einv);
            // clamp to the core (NaN -> a); a point is in the core iff the clamp left it alone
            const float2 rc = p2(fminf(fmaxf(r.x, ts.a), ts.b), fminf(fmaxf(r.y, ts.a), ts.b));
            const bool core0 = rc.x == r.x, core1 = rc.y == r.y;
            // core_locate_fast (ptm_math.cuh) for both observations
            const float2 tg2 = pmul(psub(rc, p1(ts.a)), p1(ts.inv_h));
            float2 fm = p2(floorf(tg2.x), floorf(tg2.y));
            float2 frac = psub(tg2, fm);
            if (frac.x < ts.frac_eps || frac.x > 1.f - ts.frac_eps) grid_cell_exact(ts, a.grid, rc.x, fm.x, frac.x);
            if (frac.y < ts.frac_eps || frac.y > 1.f - ts.frac_eps) grid_cell_exact(ts, a.grid, rc.y, fm.y, frac.y);
            const float2 pp = pmul(fm, S1);
            const float kmax = (float)(KK - 1);
            const float2 fk = p2(fminf(floorf(pp.x), kmax), fminf(floorf(pp.y), kmax));
            struct { int kk; } e0{(int)fk.x}, e1{(int)fk.y};
            const float2 u = psub(pp, fk), kap = pmul(frac, p1(ts.kap_scale));
            const float2 over = padd(padd(u, S1), p1(-1.f));
            const float2 mm = p2(fmaxf(over.x, 0.f), fmaxf(over.y, 0.f));
            const float* cp0 = ccl + (e0.kk * 5) * 128;
            const float* cp1 = ccl + (e1.kk * 5) * 128;
            const float2 c0 = p2(cp0[0], cp1[0]), c1 = p2(cp0[128], cp1[128]), c2 = p2(cp0[256], cp1[256]);
            const float2 c3 = p2(cp0[384], cp1[384]), jmp = p2(cp0[512], cp1[512]);
            // core_eval (ptm_math.cuh), two observations per instruction
            const float2 b2 = pfma(u, c3, c2);
            const float2 b1 = pfma(u, b2, c1);
            const float2 P = pfma(u, b1, c0);
            const float2 d1 = pfma(u, c3, b2);
            const float2 Pp = pfma(u, d1, b1);
            const float2 eh = pfma(u, c3, d1);  // P'' / 2
            const float2 mm2 = pmul(mm, mm);
            const float2 j3 = pmul(jmp, p1(3.f));
            const float2 dH = pfma(jmp, pmul(mm2, mm), pmul(S1, pfma(S1, pfma(S1, c3, eh), Pp)));
            float2 z = pfma(kap, dH, P);
            const float2 dD = pfma(j3, mm2, pmul(S1, pfma(S3, c3, padd(eh, eh))));
            float2 d = pmul(pfma(kap, dD, Pp), IDK);
            const float2 dD2 = pfma(padd(j3, j3), mm, pmul(pmul(S3, p1(2.f)), c3));
            float2 dddr = pmul(pmul(pfma(kap, dD2, padd(eh, eh)), IDK), IDK);
            float2 dzdr = d;
            float cf[2] = {0.f, 0.f}, qf[2] = {0.f, 0.f};
            if (!(core0 && core1)) {
              // transitions / tails of either observation: the scalar closed forms (ptm.py:349-410)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h ? core1 : core0) continue;
                const float rr = h ? r.y : r.x;
                const int code = region_code(ts, rr);
                float zz, dd, dzr, ddr, cfac = 0.f, qfac = 0.f;
                if (code == 1) {
                  const float dL = ccl[(KK * 5 + 1) * 128], constL = ccl[(NCC + 0) * 128];
                  const float poly = rr * ts.a - 0.5f * rr * rr;
                  const float t1 = rr - poly * ts.inv_w;
                  zz = (ts.inv_w * poly + dL * t1) + constL;
                  const float q = (ts.a - rr) * ts.inv_w;
                  dd = (1.f - q) * dL + q;
                  dzr = dd; ddr = (dL - 1.f) * ts.inv_w;
                  const float poly0 = ts.a * ts.a - 0.5f * ts.a * ts.a;
                  cfac = t1 - (ts.a - poly0 * ts.inv_w);
                  qfac = 1.f - q;
                } else if (code == 3) {
                  const float dR = ccl[(KK * 5 + 3) * 128], constR = ccl[(NCC + 1) * 128];
                  const float poly = 0.5f * rr * rr - rr * ts.b;
                  const float t1 = rr - poly * ts.inv_w;
                  zz = (ts.inv_w * poly + dR * t1) + constR;
                  const float q = (rr - ts.b) * ts.inv_w;
                  dd = (1.f - q) * dR + q;
                  dzr = dd; ddr = (1.f - dR) * ts.inv_w;
                  const float poly0 = 0.5f * ts.b * ts.b - ts.b * ts.b;
                  cfac = t1 - (ts.b - poly0 * ts.inv_w);
                  qfac = 1.f - q;
                } else if (code == 0) {
                  zz = ccl[(NCC + 2) * 128] - (ts.min_eps - rr);
                  dd = 1.f; dzr = 1.f; ddr = 0.f;
                  cfac = cL_of(ts, ts.min_eps);
                } else {
                  zz = ccl[(NCC + 3) * 128] + (rr - ts.max_eps);
                  dd = 1.f; dzr = 1.f; ddr = 0.f;
                  cfac = cR_of(ts, ts.max_eps);
                }
                if (h) { z.y = zz; d.y = dd; dzdr.y = dzr; dddr.y = ddr; }
                else { z.x = zz; d.x = dd; dzdr.x = dzr; dddr.x = ddr; }
                cf[h] = cfac; qf[h] = qfac;
              }
            }
            const float tiny = tiny_of<float>();
            const bool ab0 = d.x > tiny, ab1 = d.y > tiny;
            const float2 dcl = p2(ab0 ? d.x : tiny, ab1 ? d.y : tiny);
            t_z2 = fmaf(z.x, z.x, t_z2); t_z2 = fmaf(z.y, z.y, t_z2);
            t_ls += es.x + es.y;
            t_ld += mtc_log(dcl.x) + mtc_log(dcl.y);
            const float2 lz = p2(-z.x, -z.y);
            const float2 ld = p2(ab0 ? mtc_rcp(dcl.x) : 0.f, ab1 ? mtc_rcp(dcl.y) : 0.f);
            const float2 gr = pfma(lz, dzdr, pmul(ld, dddr));
            const float2 ngr = p2(-gr.x, -gr.y);
            const float2 gmu = pmul(ngr, einv);
            const float2 gls = pfma(ngr, r, p1(-1.f));
            float* gp2 = Gp + (gi - a.obs_begin) * 256;
            gp2[0] = gmu.x; gp2[128] = gls.x; gp2[256] = gmu.y; gp2[384] = gls.y;
            t_gb0 += gmu.x + gmu.y;
            t_gg0 += gls.x + gls.y;
            // core_backward (ptm_math.cuh) and the piece -> B-spline coefficient transform
            const float2 lzc = p2(core0 ? lz.x : 0.f, core1 ? lz.y : 0.f);
            const float2 ldc = p2(core0 ? ld.x : 0.f, core1 ? ld.y : 0.f);
            const float2 bq = pmul(ldc, IDK);
            const float2 u2 = pmul(u, u);
            const float2 S2 = pmul(S1, S1);
            const float2 phi1 = pfma(kap, S1, u);
            const float2 phi2 = pfma(kap, pfma(u, padd(S1, S1), S2), u2);
            const float2 phi3 = pfma(kap, pfma(u2, S3, pfma(u, pmul(S3, S1), pmul(S2, S1))), pmul(u2, u));
            const float2 w0 = lzc;
            const float2 w1 = pfma(lzc, phi1, bq);
            const float2 w2 = pfma(lzc, phi2, pmul(padd(bq, bq), phi1));
            const float2 bq3 = pmul(bq, p1(3.f));
            const float2 w3 = pfma(lzc, phi3, pmul(bq3, phi2));
            float2 w4 = pmul(kap, pfma(lzc, pmul(mm2, mm), pmul(bq3, mm2)));
            if (e0.kk + 1 >= KK) w4.x = 0.f;
            if (e1.kk + 1 >= KK) w4.y = 0.f;
            const float2 g3 = psub(w3, w4);
            const float2 SX = p1(sixth), HF = p1(0.5f);
            const float2 o0 = pfma(psub(w0, g3), SX, pmul(psub(w2, w1), HF));
            const float2 o1 = padd(psub(pmul(psub(pmul(w0, p1(4.f)), w4), SX), w2), pmul(g3, HF));
            const float2 o2 = pfma(w0, SX, pmul(padd(psub(padd(w1, w2), g3), w4), HF));
            const float2 o3 = psub(pmul(g3, SX), pmul(w4, HF));
            const float2 o4 = pmul(w4, SX);
            {
              float* gp = gvp + e0.kk * 32;
              gp[0] += o0.x; gp[32] += o1.x; gp[64] += o2.x; gp[96] += o3.x; gp[128] += o4.x;
            }
            {
              float* gp = gvp + e1.kk * 32;   // after the first observation's stores: the rows may coincide
              gp[0] += o0.y; gp[32] += o1.y; gp[64] += o2.y; gp[96] += o3.y; gp[128] += o4.y;
            }
            if (!(core0 && core1)) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (h ? core1 : core0) continue;
                const int code = region_code(ts, h ? r.y : r.x);
                const float lzz = h ? lz.y : lz.x, ldd = h ? ld.y : ld.x;
                const double dv = (double)fmaf(lzz, cf[h], ldd * qf[h]);
                if (code <= 1) { a_gvl += (double)lzz; a_gdl += dv; }
                else { a_gvr += (double)lzz; a_gdr += dv; }
              }
            }
          }
        }
        for (; i0 < cnt; ++i0) eval_one(i0);
        yreg = ynext;
        // per-tile float sums -> double (a tile is at most 16 observations per warp)
        a_z2 += (double)t_z2;
        a_ls += (double)t_ls;
        if (GRAD) { a_gb0 += (double)t_gb0; a_gg0 += (double)t_gg0; }
        a_logd += (double)t_ld;
        }  // !ETAONLY
      }
      if constexpr (!ETAONLY) {
      named_bar(1, kMtcEvalWarps * 32);  // every eval warp is done with its eta buffer (aliased below)
      double* rp = s_red + (warp * kRegSlots) * 32 + lane;
      rp[SL_Z2 * 32] = a_z2;
      rp[SL_LS * 32] = a_ls;
      rp[SL_LOGD * 32] = a_logd;
      rp[SL_GB0 * 32] = a_gb0;
      rp[SL_GG0 * 32] = a_gg0;
      rp[SL_GVL * 32] = a_gvl;
      rp[SL_GDL * 32] = a_gdl;
      rp[SL_GVR * 32] = a_gvr;
      rp[SL_GDR * 32] = a_gdr;
      named_bar(1, kMtcEvalWarps * 32);
      // sum over the four eval warps of a chain quarter -> this (chain group, chunk)'s partial
      if (cg < a.CG) {
        double* pout = a.partial + ((long long)(cg * a.M + mchunk) * a.NSLOT) * 32;
        const int nreg = GRAD ? kRegSlots : kLogpSlots;
        for (int idx = sub * 32 + lane; idx < nreg * 32; idx += 128) {
          double v = 0.0;
#pragma unroll
          for (int w = 0; w < 4; ++w) v += s_red[((w * 4 + q4) * kRegSlots) * 32 + idx];
          pout[idx] = v;
        }
        if (GRAD)
          for (int idx = sub * 32 + lane; idx < NGV * 32; idx += 128) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < 4; ++w) v += (double)s_gv[((w * 4 + q4) * NGV) * 32 + idx];
            pout[kRegSlots * 32 + idx] = v;
          }
      }
      named_bar(1, kMtcEvalWarps * 32);
      }
    }
    // every warp leaves the item with the same tile count
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  };
  // (setmaxnreg 96 / 48 on top of this split was measured slower: 17.6 against 16.7 ms at c3 --
  //  the stagers spill at 48 registers and their tile fill sits on the serial chain of the
  //  single-buffered design tile; the register file grants 512-register units per warp, so 104
  //  for the eval warps does not fit and dead-locks in setmaxnreg.inc)
  if (is_eval) run_items(std::integral_constant<int, 0>{});
  else run_items(std::integral_constant<int, 1>{});
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---------------------------------------------------------------------------------------
// backward: dGamma[chain, k] += sum_obs G_q[chain, obs] B[obs, k]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(896, 1) k_grad_tc(const MainTcArgs g) {
  const MainArgs<float>& a = g.a;
  extern __shared__ __align__(128) unsigned char smem_gtc[];
  __shared__ __align__(8) uint64_t s_bar[6];  // [0,1] A tile written, [2,3] B tile written, [4,5] MMAs done
  __shared__ __align__(8) uint64_t s_gbar[2 * kGtcStages];  // [0..S) G tile landed (tx bytes), [S..2S) G tile read
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, NSW = g.NSWb;
  const int Ntot = g.Nq[0] + g.Nq[1];
  const int halfb = Ntot * kGtcKT * 4;   // bytes of one half (hi or lo) of a B buffer, both predictor blocks
  const int bufb = 2 * halfb;
  unsigned char* s_B = smem_gtc;                                   // [2 buffers][hi, lo][Ntot x 32]
  float* s_kn = reinterpret_cast<float*>(s_B + 2 * bufb);          // [T][4][kKnotStride]
  // ring of G tiles [32 obs][2][128 chains], filled by bulk async copies (one per K tile)
  unsigned char* s_G = reinterpret_cast<unsigned char*>(s_kn + T * 4 * kKnotStride);
  s_G += (128 - (smem_u32(s_G) & 127)) & 127;
  constexpr int sboB = (kGtcKT / 4) * 128;

  const bool is_load = warp < kMtcEvalWarps;
  const int sidx = warp - kMtcEvalWarps;
  const bool is_stager = sidx >= 0 && sidx < NSW;
  const bool is_mma = sidx == NSW;
  const bool is_copy = sidx == NSW + 1;
  const int q4 = warp & 3, sub = (warp >> 2) & 3;

  for (int t = warp; t < T; t += blockDim.x >> 5) {
    float* kn = s_kn + t * 4 * kKnotStride;
    const float* gk = a.knots + t * kKnotStride;
    const int nbt = a.nb[t];
    for (int i = lane; i < nbt + 4; i += 32) {
      kn[i] = gk[i];
      kn[kKnotStride + i] = i + 1 < nbt + 4 ? 1.f / (gk[i + 1] - gk[i]) : 0.f;
      kn[2 * kKnotStride + i] = i + 2 < nbt + 4 ? 1.f / (gk[i + 2] - gk[i]) : 0.f;
      kn[3 * kKnotStride + i] = i + 3 < nbt + 4 ? 1.f / (gk[i + 3] - gk[i]) : 0.f;
    }
  }
  for (int i = tid; i < (2 * bufb) >> 4; i += blockDim.x) reinterpret_cast<float4*>(s_B)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid == 0) {
    for (int b = 0; b < 2; ++b) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[0 + b])), "r"(kMtcEvalWarps));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_bar[2 + b])), "r"(NSW));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[4 + b])));
    }
    for (int i = 0; i < kGtcStages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_gbar[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&s_gbar[kGtcStages + i])), "r"(kMtcEvalWarps));
    }
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
  const uint32_t barA = smem_u32(&s_bar[0]), barBf = smem_u32(&s_bar[2]), barM = smem_u32(&s_bar[4]);
  const uint32_t barGf = smem_u32(&s_gbar[0]), barGe = smem_u32(&s_gbar[kGtcStages]);
  const uint32_t colA = 256;  // A tiles: [256 + b * 128 + q * 64 + {0: hi, 32: lo}], D in [0, Ntot)
  uint32_t tg = 0;            // K tiles this CTA has gone through (buffer = tg & 1, use count = tg >> 1)

  int prev_r[4][2];           // stager: D row each (owned term, buffer) wrote last for this lane's observation
#pragma unroll
  for (int i = 0; i < 4; ++i) prev_r[i][0] = prev_r[i][1] = -1;

  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int cb = item / a.M, mchunk = item - cb * a.M;
    const long long o_begin = a.obs_begin + (long long)mchunk * a.chunk_len;
    long long o_end = o_begin + a.chunk_len;
    if (o_end > a.obs_end) o_end = a.obs_end;
    const long long len = o_end > o_begin ? o_end - o_begin : 0;
    const int ntiles = (int)((len + kGtcKT - 1) / kGtcKT);
    const int cg = cb * 4 + q4;

    auto flush_acc = [&](bool first) {
      // loader warps: TMEM accumulators -> double partials of k_post
      if (cg < a.CG) {
        double* pout = a.partial + ((long long)(cg * a.M + mchunk) * a.NSLOT) * 32 + lane;
        // this warp's columns sub * 8 + 32 i: up to four 8-column loads in flight per wait
        for (int c0 = sub * 8; c0 < Ntot; c0 += 128) {
          uint32_t v[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            // (columns past Ntot belong to the A tiles: read and ignored, dslot = -1 there)
            const uint32_t addr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(c0 + 32 * u);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3]), "=r"(v[u][4]), "=r"(v[u][5]),
                           "=r"(v[u][6]), "=r"(v[u][7])
                         : "r"(addr));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int slot = g.dslot[c0 + 32 * u + j];
                if (slot >= 0) {
                  double* pp = pout + (long long)slot * 32;
                  const double x = (double)__uint_as_float(v[u][j]);
                  // single writer per address: a plain store first, then fire-and-forget
                  // reductions (no load-add-store round trip; same-thread same-address order
                  // is program order)
                  if (first) *pp = x;
                  else asm volatile("red.global.add.f64 [%0], %1;" ::"l"(pp), "d"(x) : "memory");
                }
              }
            }
          }
        }
      }
    };

    if (is_load) {
      int since = 0;
      bool first = true;
      // this warp's 8 observations x 2 predictors of a K tile come out of the shared-memory ring
      // the copy warp fills with one bulk async copy per tile (kGtcStages tiles in flight)
      for (int s = 0; s < ntiles; ++s, ++tg) {
        const int b = tg & 1;
        const uint32_t use = tg >> 1;
        const uint32_t stg = tg % kGtcStages, sph = (tg / kGtcStages) & 1;
        mtc_wait<PTM_GTC_HINT, PTM_GTC_SLEEP>(barGf + 8 * stg, sph);
        const float* gs = reinterpret_cast<const float*>(s_G + (size_t)stg * kGtcTileBytes) + q4 * 32 + lane;
        const long long base = o_begin + (long long)s * kGtcKT + sub * 8;
        uint32_t hi[2][8], lo[2][8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v0 = gs[((sub * 8 + j) * 2 + 0) * 128], v1 = gs[((sub * 8 + j) * 2 + 1) * 128];
          if (base + j >= o_end) { v0 = 0.f; v1 = 0.f; }   // rows past the chunk: whatever the copy brought
          const float h0 = tf32_round(v0), h1 = tf32_round(v1);
          hi[0][j] = __float_as_uint(h0); lo[0][j] = __float_as_uint(tf32_round(v0 - h0));
          hi[1][j] = __float_as_uint(h1); lo[1][j] = __float_as_uint(tf32_round(v1 - h1));
        }
        __syncwarp();
        if (lane == 0) mtc_arrive(barGe + 8 * stg);   // this warp has read the ring slot
        if (use > 0) mtc_wait<PTM_GTC_HINT, PTM_GTC_SLEEP>(barM + 8 * b, (use - 1) & 1);  // the MMAs that read A buffer b have completed
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t addr = tmem + ((uint32_t)(q4 * 32) << 16) + colA + (uint32_t)(b * 128 + sub * 8);
        mtc_st8(addr, hi[0]);
        mtc_st8(addr + 32, lo[0]);
        mtc_st8(addr + 64, hi[1]);
        mtc_st8(addr + 96, lo[1]);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mtc_arrive(barA + 8 * b);
        ++since;
        if (since == g.flush || s == ntiles - 1) {
          // every MMA issued so far: wait for the commit of THIS tile
          mtc_wait<PTM_GTC_HINT, PTM_GTC_SLEEP>(barM + 8 * b, use & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          flush_acc(first);
          // (the next tile's MMAs overwrite D only after ALL loader warps have arrived on its
          //  A-tile barrier, i.e. after every warp has finished these reads: no extra barrier)
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          since = 0;
          first = false;
        }
      }
      if (ntiles == 0 && cg < a.CG) {
        double* pout = a.partial + ((long long)(cg * a.M + mchunk) * a.NSLOT) * 32 + lane;
        for (int c0 = sub; c0 < Ntot; c0 += 4) {
          const int slot = g.dslot[c0];
          if (slot >= 0) pout[(long long)slot * 32] = 0.0;
        }
      }
    } else if (is_stager) {
      // covariate values of the next K tile are requested one tile ahead (lane = observation)
      float xn[4] = {0.f, 0.f, 0.f, 0.f};
      auto fetch_x = [&](int s) {
        const long long gi = o_begin + (long long)s * kGtcKT + lane;
        int ti = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int t = sidx + i * NSW;
          if (t < T && gi < o_end) xn[ti] = a.X[(long long)t * a.n + gi];
          ++ti;
        }
      };
      if (ntiles > 0) fetch_x(0);
      for (int s = 0; s < ntiles; ++s, ++tg) {
        const int b = tg & 1;
        const uint32_t use = tg >> 1;
        float xc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xc[i] = xn[i];
        if (s + 1 < ntiles) fetch_x(s + 1);
        const long long gi = o_begin + (long long)s * kGtcKT + lane;  // lane = observation of the K tile
        // the basis windows are evaluated BEFORE waiting for the buffer: only the stores sit on
        // the chain MMA(s) -> refill -> MMA(s + 2)
        float bbv[4][4];
        int nrv[4];
#pragma unroll
        for (int ti = 0; ti < 4; ++ti) {
          const int t = sidx + ti * NSW;
          nrv[ti] = -1;
          if (t < T && gi < o_end && g.ncol[t] >= 0) {
            const float* kn = s_kn + t * 4 * kKnotStride;
            const int jj = basis_window_fast(xc[ti], kn, kn + kKnotStride, kn + 2 * kKnotStride,
                                             kn + 3 * kKnotStride, a.nb[t], a.inv_dk[t], bbv[ti]);
            nrv[ti] = g.ncol[t] + jj;
          }
        }
        if (use > 0) mtc_wait<PTM_GTC_HINT, PTM_GTC_SLEEP>(barM + 8 * b, (use - 1) & 1);
        unsigned char* colp = s_B + b * bufb + (lane >> 2) * 128 + (lane & 3) * 4;
#pragma unroll
        for (int ti = 0; ti < 4; ++ti) {
          const int t = sidx + ti * NSW;
          if (t >= T) break;
          const int pr = prev_r[ti][b];
          if (pr >= 0) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int r = pr + m;
              unsigned char* e = colp + (r >> 3) * sboB + (r & 7) * 16;
              *reinterpret_cast<float*>(e) = 0.f;
              *reinterpret_cast<float*>(e + halfb) = 0.f;
            }
          }
          const int nr = nrv[ti];
          if (nr >= 0) {
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              const int r = nr + m;
              const float vh = tf32_round(bbv[ti][m]);
              unsigned char* e = colp + (r >> 3) * sboB + (r & 7) * 16;
              *reinterpret_cast<float*>(e) = vh;
              *reinterpret_cast<float*>(e + halfb) = tf32_round(bbv[ti][m] - vh);
            }
          }
          prev_r[ti][b] = nr;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mtc_arrive(barBf + 8 * b);
      }
    } else if (is_mma) {
      int since = 0;
      for (int s = 0; s < ntiles; ++s, ++tg) {
        const int b = tg & 1;
        const uint32_t use = tg >> 1;
        if (lane == 0) {
          mtc_wait<PTM_GTC_HINT, PTM_GTC_SLEEP>(barA + 8 * b, use & 1);
          mtc_wait<PTM_GTC_HINT, PTM_GTC_SLEEP>(barBf + 8 * b, use & 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            if (g.Nq[q] == 0) continue;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.Nq[q] >> 3) << 17) | (8u << 24);
            const uint32_t d = tmem + (uint32_t)g.nbase[q];
            const uint32_t a0 = tmem + colA + (uint32_t)(b * 128 + q * 64);
            const uint32_t bq = smem_u32(s_B) + (uint32_t)(b * bufb + (g.nbase[q] >> 3) * sboB);
#pragma unroll
            for (int ks = 0; ks < kGtcKT / 8; ++ks) {
              const uint64_t b_hi = mtc_desc(bq + ks * 256, sboB), b_lo = mtc_desc(bq + halfb + ks * 256, sboB);
              mtc_mma_ts(d, a0 + ks * 8, b_hi, idesc, (since > 0 || ks > 0) ? 1u : 0u);
              mtc_mma_ts(d, a0 + ks * 8, b_lo, idesc, 1u);
              mtc_mma_ts(d, a0 + 32 + ks * 8, b_hi, idesc, 1u);
            }
          }
          mtc_commit(barM + 8 * b);
        }
        __syncwarp();
        ++since;
        if (since == g.flush) since = 0;
      }
    } else if (is_copy) {
      const float* Gc = g.G + ((long long)cb * g.Gld + (o_begin - a.obs_begin)) * 256;
      for (int s = 0; s < ntiles; ++s, ++tg) {
        const uint32_t stg = tg % kGtcStages, round = tg / kGtcStages;
        if (lane == 0) {
          if (round > 0) mtc_wait<PTM_GTC_HINT, PTM_GTC_SLEEP>(barGe + 8 * stg, (round - 1) & 1);   // all loader warps have read the slot
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(barGf + 8 * stg), "r"(kGtcTileBytes)
                       : "memory");
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(s_G + (size_t)stg * kGtcTileBytes)),
              "l"(Gc + (long long)s * kGtcKT * 256), "r"(kGtcTileBytes), "r"(barGf + 8 * stg)
              : "memory");
        }
        __syncwarp();
      }
    } else {
      tg += (uint32_t)ntiles;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

}  // namespace ptm

#pragma nv_diag_default 177
