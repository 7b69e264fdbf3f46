// ptm_xtwx_tc.cuh -- tensor-core route of the chain-batched IWLS information of the
// LOCATION terms (float problems).  Reference: iwls_proposals.py:55-89.
//
// The band  M_c[(r,d)] = sum_i w_ic B[i,r] B[i,r+d]  (see ptm_xtwx.cuh) is a GEMM over the
// observation axis once the weights are taken as one operand and the chain-independent
// band expansion of the design as the other:
//
//     D[chain, slot] = sum_i  W[chain, i] * Q[slot, i],      slot = 4*row + diagonal,
//
//   A = W   [128 chains x K obs]  w_ic = 1/max(sigma_ic, sqrt(eps))^2, produced by the
//                                 weight warps (one value per observation and chain,
//                                 shared by every location term),
//   B = Q_t [4*nb slots x K obs]  10 non-zeros per observation (products of the four
//                                 non-zero basis values), produced once per CTA and shared
//                                 by its 128 chains,
//   D in TMEM, one accumulator block of 4*nb columns per term.
//
// tcgen05.mma kind::tf32 with the 3xTF32 split (hi*hi + hi*lo + lo*hi; hi = cvt.rna.tf32,
// lo = x - hi), M = 128, N = 4*nb rounded to 16, K = 8 per instruction; operands are
// written straight into the K-major no-swizzle core-matrix layout (8 rows x 16 bytes) by
// the producing warps, so no TMA is involved: nothing of A or B ever exists in HBM.  The
// f32 TMEM accumulators are promoted into double partials every `flush_tiles` tiles.
// tools/micro/umma_tf32.cu is the known-answer test of the descriptor encodings used here.
#pragma once
#include <cstdint>

#include "ptm_xtwx.cuh"

namespace ptm {

constexpr int kTcKT = 16;                    // observations per tile = two K=8 MMAs
static_assert(kTcKT == 16, "the MMA issue code is written for two K=8 steps per tile");
constexpr int kTcWW = 16;                    // weight warps: 4 chain quarters x 4 obs quads
constexpr int kTcLBO = 128;                  // bytes between the 16-byte K chunks
constexpr int kTcSBO = (kTcKT / 4) * 128;    // bytes between 8-row groups
constexpr int kTcWTile = 128 * kTcKT * 4;    // bytes of one W tile (hi or lo)

struct XtwxTcArgs {
  XtwxArgs<float> a;
  int npad[kMaxTerms];   // accumulator columns (4*nb rounded up to 16) per requested term
  int col[kMaxTerms];    // first TMEM column of the term
  int qoff[kMaxTerms];   // byte offset of the term's Q tiles (hi then lo) in a Q buffer
  int qbytes;            // bytes of one Q buffer (all terms)
  int flush_tiles;       // tiles between promotions of the accumulators to double
  int CB;                // chain blocks of 128
};

constexpr float kLogSqrtEpsF = -7.9711924f;  // log(sqrt(FLT_EPSILON)) = log(3.4526698e-04)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

__device__ __forceinline__ uint64_t tc_desc(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)(kTcLBO >> 4) << 16) | ((uint64_t)(kTcSBO >> 4) << 32) |
         ((uint64_t)1 << 46);
}

__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}

// Same with the descriptors given as (low word, shared high word): advancing an operand is
// one 32-bit add on the address field, which keeps the single issuing thread short.
__device__ __forceinline__ void tc_mma2(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t}\n" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t addr) { return ((addr >> 4) & 0x3FFF) | ((uint32_t)(kTcLBO >> 4) << 16); }
__device__ __forceinline__ uint32_t tc_desc_hi(int sbo) { return ((uint32_t)(sbo >> 4) & 0x3FFF) | (1u << 14); }

// Bounded wait: a protocol bug traps instead of hanging the device.
__device__ __forceinline__ void tc_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (int it = 0; it < (1 << 24) && !done; ++it)
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  if (!done) __trap();
}

// NSC > 0: number of scale terms known at compile time (predictor loop fully unrolled).
template <int NSC>
__global__ void __launch_bounds__(1024, 1) k_xtwx_tc(const XtwxTcArgs g) {
  const XtwxArgs<float>& a = g.a;
  extern __shared__ __align__(128) unsigned char smem_tc[];
  __shared__ __align__(8) uint64_t s_bar[2];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsc = a.nsc, nbt = a.nbt, NS = nsc + nbt;
  unsigned char* s_W = smem_tc;                                    // [2][hi,lo][128 x KT]
  unsigned char* s_Q = s_W + 4 * kTcWTile;                         // [2][qbytes]
  float* s_gam = reinterpret_cast<float*>(s_Q + 2 * g.qbytes);     // [sum_sc_nb][128]
  float* s_b = s_gam + a.sum_sc_nb * 128;                          // [nsc][2][KT][4]
  float* s_kn = s_b + nsc * 2 * kTcKT * 4;                         // [NS][4][kKnotStride]
  int* s_j = reinterpret_cast<int*>(s_kn + NS * 4 * kKnotStride);  // [nsc][2][KT]

  const bool is_weight = warp < kTcWW;
  const int sidx = warp - kTcWW;
  const bool is_stager = sidx >= 0 && sidx < nsc;
  const bool is_q = sidx >= nsc && sidx < NS;
  const bool is_mma = sidx == NS;
  const int term = is_stager ? a.sc_term[sidx] : (is_q ? a.bt_term[sidx - nsc] : 0);
  int t_nb = 0;
  float t_invdk = 0.f;
  const float* t_x = nullptr;
  if (is_stager || is_q) {
    t_nb = a.nb[term];
    t_invdk = a.inv_dk[term];
    t_x = a.X + (long long)term * a.n;
    float* kn = s_kn + sidx * 4 * kKnotStride;
    const float* gk = a.knots + term * kKnotStride;
    for (int i = lane; i < t_nb + 4; i += 32) {
      kn[i] = gk[i];
      kn[kKnotStride + i] = i + 1 < t_nb + 4 ? 1.f / (gk[i + 1] - gk[i]) : 0.f;
      kn[2 * kKnotStride + i] = i + 2 < t_nb + 4 ? 1.f / (gk[i + 2] - gk[i]) : 0.f;
      kn[3 * kKnotStride + i] = i + 3 < t_nb + 4 ? 1.f / (gk[i + 3] - gk[i]) : 0.f;
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t bar0 = smem_u32(&s_bar[0]);
  uint32_t use0 = 0, use1 = 0;  // commits issued so far on each buffer's barrier (tracked by every thread)

  const int q4 = warp & 3, og = (warp >> 2) & 3;  // weight warp: chain quarter, observation quad
  const int row = q4 * 32 + lane;
  const int woff = (row >> 3) * kTcSBO + og * kTcLBO + (row & 7) * 16;

  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int cb = item / a.M, mchunk = item - cb * a.M;
    const long long o_begin = a.obs_begin + (long long)mchunk * a.chunk_len;
    long long o_end = o_begin + a.chunk_len;
    if (o_end > a.obs_end) o_end = a.obs_end;
    const long long len = o_end > o_begin ? o_end - o_begin : 0;
    const int ntiles = (int)((len + kTcKT - 1) / kTcKT);

    auto stage = [&](int tile) {  // scale stager: interval + basis of tile `tile`
      const int slot = tile & 1;
      const float* kn = s_kn + sidx * 4 * kKnotStride;
      if (lane < kTcKT) {
        const long long gi = o_begin + (long long)tile * kTcKT + lane;
        float bb[4] = {0.f, 0.f, 0.f, 0.f};
        int jj = 0;  // observations past the chunk end: a harmless in-range gather
        if (gi < o_end)
          jj = basis_window_fast(t_x[gi], kn, kn + kKnotStride, kn + 2 * kKnotStride, kn + 3 * kKnotStride, t_nb,
                                 t_invdk, bb);
        store4(s_b + ((sidx * 2 + slot) * kTcKT + lane) * 4, bb);
        s_j[(sidx * 2 + slot) * kTcKT + lane] = (a.sc_goff[sidx] + jj) * 128;
      }
    };
    if (is_stager) {
      for (int j = 0; j < t_nb; ++j)
        for (int q = 0; q < 4; ++q) {
          int ch = cb * 128 + q * 32 + lane;
          if (ch >= a.C) ch = a.C - 1;
          s_gam[(a.sc_goff[sidx] + j) * 128 + q * 32 + lane] = a.Gamma[((long long)term * kMaxNbases + j) * a.C + ch];
        }
      if (ntiles > 0) stage(0);
    }
    float gamma0 = 0.f;
    if (is_weight) {
      int ch = cb * 128 + row;
      if (ch >= a.C) ch = a.C - 1;
      gamma0 = a.cc[(long long)(a.KK * 5 + 5) * a.C + ch];
    }
    cta_bar();

    int since = 0;
    bool first = true;
    for (int s = 0; s < ntiles; ++s) {
      const int b = s & 1;
      const long long base = o_begin + (long long)s * kTcKT;
      const uint32_t ub = b ? use1 : use0;
      if (is_stager && s + 1 < ntiles) stage(s + 1);
      if (is_weight || is_q) {
        if (ub > 0) tc_wait(bar0 + 8 * b, (ub - 1) & 1);  // the MMAs that read buffer b have completed
      }
      if (is_weight) {
        float hi[4], lo[4];
        const float* gl = s_gam + row;
        const int nq = NSC > 0 ? NSC : nsc;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const int o = og * 4 + k4;
          float es = gamma0;
          const float* bq = s_b + (b * kTcKT + o) * 4;
          const int* jq = s_j + b * kTcKT + o;
#pragma unroll
          for (int q = 0; q < nq; ++q) {
            float b0, b1, b2, b3;
            load4(bq + q * (2 * kTcKT * 4), b0, b1, b2, b3);
            const float* gp = gl + jq[q * (2 * kTcKT)];
            float f = b0 * gp[0];
            f = fmaf(b1, gp[128], f);
            f = fmaf(b2, gp[256], f);
            f = fmaf(b3, gp[384], f);
            es += f;
          }
          // w = 1/max(sigma, sqrt(eps))^2 = exp(-2 max(eta_sigma, log sqrt(eps)))
          const float ec = es < kLogSqrtEpsF ? kLogSqrtEpsF : es;  // NaN propagates like jnp.clip
          float wgt;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wgt) : "f"(ec * -2.8853900817779268f));
          if (base + o >= o_end) wgt = 0.f;
          hi[k4] = tf32_round(wgt);
          lo[k4] = wgt - hi[k4];
        }
        unsigned char* wt = s_W + b * 2 * kTcWTile + woff;
        *reinterpret_cast<float4*>(wt) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(wt + kTcWTile) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      } else if (is_q) {
        const int i = sidx - nsc;
        const int npad = g.npad[i];
        unsigned char* qt = s_Q + b * g.qbytes + g.qoff[i];
        float4* q4p = reinterpret_cast<float4*>(qt);
        for (int idx = lane; idx < npad * 8; idx += 32) q4p[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const int k = lane & 15, half = lane >> 4;
        const long long gi = base + k;
        if (gi < o_end) {
          const float* kn = s_kn + sidx * 4 * kKnotStride;
          float bb[4];
          const int jj = basis_window_fast(t_x[gi], kn, kn + kKnotStride, kn + 2 * kKnotStride, kn + 3 * kKnotStride,
                                           t_nb, t_invdk, bb);
          unsigned char* qh = qt + half * npad * (kTcKT * 4) + (k >> 2) * kTcLBO + (k & 3) * 4;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int d = 0; d < 4 - r; ++d) {
              const float v = bb[r] * bb[r + d];
              const float vh = tf32_round(v);
              const int slot = (jj + r) * 4 + d;
              *reinterpret_cast<float*>(qh + (slot >> 3) * kTcSBO + (slot & 7) * 16) = half ? v - vh : vh;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      cta_bar();
      if (is_mma && lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t wbase = smem_u32(s_W) + b * 2 * kTcWTile;
        for (int i = 0; i < nbt; ++i) {
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.npad[i] >> 3) << 17) | (8u << 24);
          const uint32_t qbase = smem_u32(s_Q) + b * g.qbytes + g.qoff[i];
          const uint32_t qlo = qbase + g.npad[i] * (kTcKT * 4);
          const uint32_t d = tmem + (uint32_t)g.col[i];
#pragma unroll
          for (int ks = 0; ks < kTcKT / 8; ++ks) {
            const uint32_t ko = ks * 2 * kTcLBO;
            tc_mma(d, tc_desc(wbase + ko), tc_desc(qbase + ko), idesc, (since > 0 || ks > 0) ? 1u : 0u);
            tc_mma(d, tc_desc(wbase + ko), tc_desc(qlo + ko), idesc, 1u);
            tc_mma(d, tc_desc(wbase + kTcWTile + ko), tc_desc(qbase + ko), idesc, 1u);
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar0 + 8 * b)
                     : "memory");
      }
      if (b) ++use1; else ++use0;
      ++since;
      if (since == g.flush_tiles || s == ntiles - 1) {
        // promote the f32 accumulators: wait for every MMA issued so far, TMEM -> double partials
        tc_wait(bar0 + 8 * b, ((b ? use1 : use0) - 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int cg = cb * 4 + q4;
        if (is_weight && cg < a.CG) {
          double* pout = a.partial + ((long long)(cg * a.M + mchunk) * a.band_slots) * 32 + lane;
          for (int i = 0; i < nbt; ++i) {
            const int nslots = a.nb[a.bt_term[i]] * 4;
            for (int c0 = og * 8; c0 < nslots; c0 += 32) {
              uint32_t v[8];
              const uint32_t addr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(g.col[i] + c0);
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                             "=r"(v[7])
                           : "r"(addr));
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              double* pp = pout + (long long)(a.bt_boff[i] + c0) * 32;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (c0 + j < nslots) {
                  const double x = (double)__uint_as_float(v[j]);
                  pp[j * 32] = first ? x : pp[j * 32] + x;
                }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        since = 0;
        first = false;
        cta_bar();
      }
    }
    if (ntiles == 0) {
      const int cg = cb * 4 + q4;
      if (is_weight && cg < a.CG) {
        double* pout = a.partial + ((long long)(cg * a.M + mchunk) * a.band_slots) * 32 + lane;
        for (int i = og; i < a.band_slots; i += 4) pout[(long long)i * 32] = 0.0;
      }
    }
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// ---------------------------------------------------------------------------------------
// Version 2: the scale predictor is a GEMM too,
//     eta_sigma[chain, obs] = sum_k Gamma[chain, k] * Bs[obs, k],   k = (scale term, coefficient),
// A = Gamma [128 chains x Kp] (written once per work item, hi/lo), B = Bs [16 obs x Kp]
// (4 non-zeros per scale term and observation, written per tile by the stagers, hi/lo),
// D = 16 TMEM columns (double-buffered).  The weight warps then read their four eta values
// with one tcgen05.ld and spend one ex2 + the hi/lo split per value.  Issue order per tile
// boundary: XtWX MMAs of tile s, then the eta MMAs of tile s+2, so eta is ready a full
// step before it is consumed.  The Q tiles are single-buffered when shared memory is short
// (c3: Gamma hi/lo alone is 106 KB); their refill then overlaps the eta MMAs.
// ---------------------------------------------------------------------------------------
struct XtwxTc2Args {
  XtwxTcArgs g;
  int koff[kMaxTerms];   // first K column of each scale term (multiple of 4)
  int Kp;                // K extent (multiple of 8)
  int qbuf;              // 1 or 2 Q buffers
};

// TMEM columns [416, 512): per tile parity three eta accumulators of 16 columns, one per
// 3xTF32 product, so that the K loop is three independent accumulation chains (MMAs that
// accumulate into the same tile serialise on its latency); the weight warps add them.
constexpr int kTcEtaAcc = 3;
constexpr int kTcEtaCol = 512 - 2 * kTcEtaAcc * kTcKT;

__global__ void __launch_bounds__(1024, 1) k_xtwx_tc2(const XtwxTc2Args h) {
  const XtwxTcArgs& g = h.g;
  const XtwxArgs<float>& a = g.a;
  extern __shared__ __align__(128) unsigned char smem_tc[];
  __shared__ __align__(8) uint64_t s_bar[4];  // [0,1] XtWX MMAs of tile parity, [2,3] eta MMAs
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsc = a.nsc, nbt = a.nbt, NS = nsc + nbt;
  const int Kp = h.Kp, sboG = (Kp >> 2) * 128;
  const int gbytes = 128 * Kp * 4, bbytes = kTcKT * Kp * 4;
  unsigned char* s_W = smem_tc;                         // [2][hi,lo][128 x KT]
  unsigned char* s_Q = s_W + 4 * kTcWTile;              // [qbuf][qbytes]
  unsigned char* s_G = s_Q + h.qbuf * g.qbytes;         // [hi,lo][128 x Kp]
  unsigned char* s_B = s_G + 2 * gbytes;                // [2][hi,lo][KT x Kp]
  float* s_kn = reinterpret_cast<float*>(s_B + 4 * bbytes);  // [NS][4][kKnotStride]

  const bool is_weight = warp < kTcWW;
  const int sidx = warp - kTcWW;
  const bool is_stager = sidx >= 0 && sidx < nsc;
  const bool is_q = sidx >= nsc && sidx < NS;
  const bool is_mma = sidx == NS;       // issues the XtWX MMAs
  const bool is_mma_eta = sidx == NS + 1;  // issues the eta MMAs (in parallel with the other issuer)
  const int term = is_stager ? a.sc_term[sidx] : (is_q ? a.bt_term[sidx - nsc] : 0);
  int t_nb = 0;
  float t_invdk = 0.f;
  const float* t_x = nullptr;
  if (is_stager || is_q) {
    t_nb = a.nb[term];
    t_invdk = a.inv_dk[term];
    t_x = a.X + (long long)term * a.n;
    float* kn = s_kn + sidx * 4 * kKnotStride;
    const float* gk = a.knots + term * kKnotStride;
    for (int i = lane; i < t_nb + 4; i += 32) {
      kn[i] = gk[i];
      kn[kKnotStride + i] = i + 1 < t_nb + 4 ? 1.f / (gk[i + 1] - gk[i]) : 0.f;
      kn[2 * kKnotStride + i] = i + 2 < t_nb + 4 ? 1.f / (gk[i + 2] - gk[i]) : 0.f;
      kn[3 * kKnotStride + i] = i + 3 < t_nb + 4 ? 1.f / (gk[i + 3] - gk[i]) : 0.f;
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;
  const uint32_t barM = smem_u32(&s_bar[0]), barE = smem_u32(&s_bar[2]);
  // commits issued before the current work item on each barrier (tracked by every thread)
  uint32_t baseM0 = 0, baseM1 = 0, baseE0 = 0, baseE1 = 0;

  const int q4 = warp & 3, og = (warp >> 2) & 3;
  const int row = q4 * 32 + lane;
  const int woff = (row >> 3) * kTcSBO + og * kTcLBO + (row & 7) * 16;
  const int my_koff = is_stager ? h.koff[sidx] : 0;
  const int nb4 = (t_nb + 3) & ~3;

  const uint32_t ghi = tc_desc_hi(sboG), whi = tc_desc_hi(kTcSBO);
  auto wait_tile = [&](uint32_t bar, uint32_t base0, uint32_t base1, int t) {  // commit of tile t
    tc_wait(bar + 8 * (t & 1), (((t & 1) ? base1 : base0) + (uint32_t)(t >> 1)) & 1);
  };
  const uint32_t idescE = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kTcKT >> 3) << 17) | (8u << 24);
  const uint32_t g_lo0 = tc_desc_lo(smem_u32(s_G)), gl_lo0 = tc_desc_lo(smem_u32(s_G) + gbytes);
  auto issue_eta = [&](int t) {  // eta MMAs of tile t into the accumulators of parity t & 1
    const uint32_t bbase = smem_u32(s_B) + (t & 1) * 2 * bbytes;
    uint32_t g_lo = g_lo0, gl_lo = gl_lo0, b_lo = tc_desc_lo(bbase), bl_lo = tc_desc_lo(bbase + bbytes);
    const uint32_t d = tmem + (uint32_t)(kTcEtaCol + (t & 1) * kTcEtaAcc * kTcKT);
    const int nks = Kp >> 3;
    for (int ks = 0; ks < nks; ++ks) {
      const uint32_t acc = ks > 0 ? 1u : 0u;
      tc_mma2(d, g_lo, b_lo, ghi, idescE, acc);
      tc_mma2(d + kTcKT, g_lo, bl_lo, ghi, idescE, acc);
      tc_mma2(d + 2 * kTcKT, gl_lo, b_lo, ghi, idescE, acc);
      g_lo += (2 * kTcLBO) >> 4; gl_lo += (2 * kTcLBO) >> 4; b_lo += (2 * kTcLBO) >> 4; bl_lo += (2 * kTcLBO) >> 4;
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barE + 8 * (t & 1))
                 : "memory");
  };

  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int cb = item / a.M, mchunk = item - cb * a.M;
    const long long o_begin = a.obs_begin + (long long)mchunk * a.chunk_len;
    long long o_end = o_begin + a.chunk_len;
    if (o_end > a.obs_end) o_end = a.obs_end;
    const long long len = o_end > o_begin ? o_end - o_begin : 0;
    const int ntiles = (int)((len + kTcKT - 1) / kTcKT);

    auto stage = [&](int tile) {  // rows of the sparse basis operand Bs for tile `tile`
      unsigned char* bt = s_B + (tile & 1) * 2 * bbytes;
      // zero this term's K range of both halves: per 8-row group a contiguous run of chunks
      const int run4 = (nb4 >> 2) * 8;  // float4 per run (nb4/4 chunks of 128 B)
#pragma unroll
      for (int blk = 0; blk < 4; ++blk) {  // half (hi/lo) x row group
        float4* p4 = reinterpret_cast<float4*>(bt + (blk >> 1) * bbytes + (blk & 1) * sboG + (my_koff >> 2) * kTcLBO);
        for (int w = lane; w < run4; w += 32) p4[w] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncwarp();
      const int n = lane & 15, half = lane >> 4;
      const long long gi = o_begin + (long long)tile * kTcKT + n;
      if (gi < o_end) {
        const float* kn = s_kn + sidx * 4 * kKnotStride;
        float bb[4];
        const int jj = basis_window_fast(t_x[gi], kn, kn + kKnotStride, kn + 2 * kKnotStride, kn + 3 * kKnotStride, t_nb,
                                         t_invdk, bb);
        unsigned char* rowp = bt + half * bbytes + (n >> 3) * sboG + (n & 7) * 16;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int k = my_koff + jj + m;
          const float vh = tf32_round(bb[m]);
          *reinterpret_cast<float*>(rowp + (k >> 2) * kTcLBO + (k & 3) * 4) = half ? bb[m] - vh : vh;
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    };
    if (is_stager) {
      // Gamma operand: this term's coefficients of the 128 chains (zero rows up to nb4)
      for (int j = 0; j < nb4; ++j)
        for (int q = 0; q < 4; ++q) {
          const int m = q * 32 + lane;
          int ch = cb * 128 + m;
          if (ch >= a.C) ch = a.C - 1;
          const float v = j < t_nb ? a.Gamma[((long long)term * kMaxNbases + j) * a.C + ch] : 0.f;
          const float vh = tf32_round(v);
          const int k = my_koff + j;
          unsigned char* p = s_G + (m >> 3) * sboG + (k >> 2) * kTcLBO + (m & 7) * 16 + (k & 3) * 4;
          *reinterpret_cast<float*>(p) = vh;
          *reinterpret_cast<float*>(p + gbytes) = v - vh;
        }
      if (sidx == nsc - 1) {  // K padding columns behind the last term
        for (int k = my_koff + nb4; k < Kp; ++k)
          for (int q = 0; q < 4; ++q) {
            const int m = q * 32 + lane;
            unsigned char* p = s_G + (m >> 3) * sboG + (k >> 2) * kTcLBO + (m & 7) * 16 + (k & 3) * 4;
            *reinterpret_cast<float*>(p) = 0.f;
            *reinterpret_cast<float*>(p + gbytes) = 0.f;
          }
        for (int idx = lane; idx < 2 * 2 * kTcKT * (Kp - my_koff - nb4); idx += 32) {  // and of both Bs buffers
          const int per = Kp - my_koff - nb4;
          const int k = my_koff + nb4 + idx % per, r = (idx / per) % kTcKT, hb = idx / (per * kTcKT);
          *reinterpret_cast<float*>(s_B + hb * bbytes + (r >> 3) * sboG + (k >> 2) * kTcLBO + (r & 7) * 16 + (k & 3) * 4) = 0.f;
        }
      }
      if (ntiles > 0) stage(0);
      if (ntiles > 1) stage(1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    float gamma0 = 0.f;
    if (is_weight) {
      int ch = cb * 128 + row;
      if (ch >= a.C) ch = a.C - 1;
      gamma0 = a.cc[(long long)(a.KK * 5 + 5) * a.C + ch];
    }
    cta_bar();
    if (is_mma_eta && lane == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (ntiles > 0) issue_eta(0);
      if (ntiles > 1) issue_eta(1);
    }

    int since = 0;
    bool first = true;
    for (int s = 0; s < ntiles; ++s) {
      const int b = s & 1;
      const long long base = o_begin + (long long)s * kTcKT;
      if (is_stager && s + 2 < ntiles) {
        wait_tile(barE, baseE0, baseE1, s);  // eta MMAs of tile s have read Bs buffer b
        stage(s + 2);
      }
      if (is_weight) {
        wait_tile(barE, baseE0, baseE1, s);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v[4], v1[4], v2[4];
        const uint32_t addr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(kTcEtaCol + b * kTcEtaAcc * kTcKT + og * 4);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                     : "r"(addr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v1[0]), "=r"(v1[1]), "=r"(v1[2]), "=r"(v1[3])
                     : "r"(addr + kTcKT));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v2[0]), "=r"(v2[1]), "=r"(v2[2]), "=r"(v2[3])
                     : "r"(addr + 2 * kTcKT));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float hi[4], lo[4];
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const float es = gamma0 + (__uint_as_float(v[k4]) + (__uint_as_float(v1[k4]) + __uint_as_float(v2[k4])));
          const float ec = es < kLogSqrtEpsF ? kLogSqrtEpsF : es;  // NaN propagates like jnp.clip
          float wgt;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(wgt) : "f"(ec * -2.8853900817779268f));
          if (base + og * 4 + k4 >= o_end) wgt = 0.f;
          hi[k4] = tf32_round(wgt);
          lo[k4] = wgt - hi[k4];
        }
        if (s >= 2) wait_tile(barM, baseM0, baseM1, s - 2);  // W buffer b is free
        unsigned char* wt = s_W + b * 2 * kTcWTile + woff;
        *reinterpret_cast<float4*>(wt) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(wt + kTcWTile) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      } else if (is_q) {
        const int qb = h.qbuf == 2 ? b : 0;
        if (s >= h.qbuf) wait_tile(barM, baseM0, baseM1, s - h.qbuf);  // Q buffer is free
        const int i = sidx - nsc;
        const int npad = g.npad[i];
        unsigned char* qt = s_Q + qb * g.qbytes + g.qoff[i];
        float4* q4p = reinterpret_cast<float4*>(qt);
        for (int idx = lane; idx < npad * 8; idx += 32) q4p[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
        __syncwarp();
        const int k = lane & 15, half = lane >> 4;
        const long long gi = base + k;
        if (gi < o_end) {
          const float* kn = s_kn + sidx * 4 * kKnotStride;
          float bb[4];
          const int jj = basis_window_fast(t_x[gi], kn, kn + kKnotStride, kn + 2 * kKnotStride, kn + 3 * kKnotStride,
                                           t_nb, t_invdk, bb);
          unsigned char* qh = qt + half * npad * (kTcKT * 4) + (k >> 2) * kTcLBO + (k & 3) * 4;
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int d = 0; d < 4 - r; ++d) {
              const float v = bb[r] * bb[r + d];
              const float vh = tf32_round(v);
              const int slot = (jj + r) * 4 + d;
              *reinterpret_cast<float*>(qh + (slot >> 3) * kTcSBO + (slot & 7) * 16) = half ? v - vh : vh;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      cta_bar();
      if (is_mma && lane == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t w_lo = tc_desc_lo(smem_u32(s_W) + b * 2 * kTcWTile), wl_lo = w_lo + (kTcWTile >> 4);
        const int qb = h.qbuf == 2 ? b : 0;
        const uint32_t first_acc = since > 0 ? 1u : 0u;
        for (int i = 0; i < nbt; ++i) {
          const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(g.npad[i] >> 3) << 17) | (8u << 24);
          const uint32_t q_lo = tc_desc_lo(smem_u32(s_Q) + qb * g.qbytes + g.qoff[i]);
          const uint32_t ql_lo = q_lo + ((g.npad[i] * (kTcKT * 4)) >> 4);
          const uint32_t d = tmem + (uint32_t)g.col[i];
          constexpr uint32_t kst = (2 * kTcLBO) >> 4;
          tc_mma2(d, w_lo, q_lo, whi, idesc, first_acc);
          tc_mma2(d, w_lo, ql_lo, whi, idesc, 1u);
          tc_mma2(d, wl_lo, q_lo, whi, idesc, 1u);
          tc_mma2(d, w_lo + kst, q_lo + kst, whi, idesc, 1u);
          tc_mma2(d, w_lo + kst, ql_lo + kst, whi, idesc, 1u);
          tc_mma2(d, wl_lo + kst, q_lo + kst, whi, idesc, 1u);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(barM + 8 * b)
                     : "memory");
      }
      if (is_mma_eta && lane == 0 && s + 2 < ntiles) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_eta(s + 2);
      }
      ++since;
      if (since == g.flush_tiles || s == ntiles - 1) {
        wait_tile(barM, baseM0, baseM1, s);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int cg = cb * 4 + q4;
        if (is_weight && cg < a.CG) {
          double* pout = a.partial + ((long long)(cg * a.M + mchunk) * a.band_slots) * 32 + lane;
          for (int i = 0; i < nbt; ++i) {
            const int nslots = a.nb[a.bt_term[i]] * 4;
            for (int c0 = og * 8; c0 < nslots; c0 += 32) {
              uint32_t v[8];
              const uint32_t addr = tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(g.col[i] + c0);
              asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                           : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                             "=r"(v[7])
                           : "r"(addr));
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
              double* pp = pout + (long long)(a.bt_boff[i] + c0) * 32;
#pragma unroll
              for (int j = 0; j < 8; ++j)
                if (c0 + j < nslots) {
                  const double x = (double)__uint_as_float(v[j]);
                  pp[j * 32] = first ? x : pp[j * 32] + x;
                }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        since = 0;
        first = false;
        cta_bar();
      }
    }
    if (ntiles == 0) {
      const int cg = cb * 4 + q4;
      if (is_weight && cg < a.CG) {
        double* pout = a.partial + ((long long)(cg * a.M + mchunk) * a.band_slots) * 32 + lane;
        for (int i = og; i < a.band_slots; i += 4) pout[(long long)i * 32] = 0.0;
      }
    }
    baseM0 += (uint32_t)((ntiles + 1) >> 1); baseM1 += (uint32_t)(ntiles >> 1);
    baseE0 += (uint32_t)((ntiles + 1) >> 1); baseE1 += (uint32_t)(ntiles >> 1);
    __syncthreads();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Plans and launches the tensor-core sweep for the location terms in a.bt_term[0..nbt).
// Returns 0, a cudaError_t, or -1 when the terms do not fit (TMEM columns, shared memory,
// warps) -- the caller then uses the banded CUDA-core sweep.
inline int xtwx_tc_sweep(XtwxArgs<float>& a, double* partial, int num_sms, int flush_tiles, cudaStream_t st,
                         int* launches) {
  if (a.const_w || a.nbt < 1) return -1;
  const int NS = a.nsc + a.nbt;
  if (kTcWW + NS + 1 > 32) return -1;
  XtwxTcArgs g;
  memset(&g, 0, sizeof g);
  int col = 0, qoff = 0;
  for (int i = 0; i < a.nbt; ++i) {
    const int npad = (a.nb[a.bt_term[i]] * 4 + 15) / 16 * 16;
    g.npad[i] = npad; g.col[i] = col; g.qoff[i] = qoff;
    col += npad;
    qoff += 2 * npad * kTcKT * 4;
  }
  if (col > 512) return -1;
  g.qbytes = qoff;
  g.flush_tiles = flush_tiles > 0 ? flush_tiles : 64;
  size_t smem = (size_t)4 * kTcWTile + (size_t)2 * g.qbytes + (size_t)a.sum_sc_nb * 128 * 4 +
                (size_t)a.nsc * 2 * kTcKT * 4 * 4 + (size_t)NS * 4 * kKnotStride * 4 + (size_t)a.nsc * 2 * kTcKT * 4 + 64;
  if (smem > 227 * 1024 - 64) return -1;
  a.partial = partial;
  a.CG = (a.C + 31) / 32;
  g.CB = (a.C + 127) / 128;
  const long long len = a.obs_end > a.obs_begin ? a.obs_end - a.obs_begin : 0;
  long long M = std::max<long long>(1, std::min(num_sms, 160) / g.CB);
  M = std::min<long long>(M, std::max<long long>(1, len / (16LL * kTcKT)));
  long long chunk = std::max<long long>(1, (len + M - 1) / M);
  M = std::max<long long>(1, (len + chunk - 1) / chunk);
  a.M = (int)M;
  a.chunk_len = chunk;
  a.n_items = g.CB * a.M;
  g.a = a;
  const int warps = (kTcWW + NS + 1 + 3) / 4 * 4;
  const int grid = std::min(a.n_items, num_sms);
  cudaError_t e = cudaSuccess;
  auto launch = [&](auto kern) {
    e = ensure_smem(kern, smem);
    if (e == cudaSuccess) kern<<<grid, 32 * warps, smem, st>>>(g);
  };
  switch (a.nsc) {
    case 1: launch(k_xtwx_tc<1>); break;
    case 2: launch(k_xtwx_tc<2>); break;
    case 3: launch(k_xtwx_tc<3>); break;
    case 4: launch(k_xtwx_tc<4>); break;
    case 5: launch(k_xtwx_tc<5>); break;
    default: launch(k_xtwx_tc<0>); break;
  }
  if (e != cudaSuccess) return (int)e;
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++*launches;
  return 0;
}

// Version 2 (eta_sigma on the tensor cores as well).  Same contract as xtwx_tc_sweep.
inline int xtwx_tc2_sweep(XtwxArgs<float>& a, double* partial, int num_sms, int flush_tiles, bool allow_single_q,
                          cudaStream_t st, int* launches) {
  if (a.const_w || a.nbt < 1 || a.nsc < 1) return -1;
  const int NS = a.nsc + a.nbt;
  if (kTcWW + NS + 2 > 32) return -1;
  XtwxTc2Args h;
  memset(&h, 0, sizeof h);
  XtwxTcArgs& g = h.g;
  int col = 0, qoff = 0;
  for (int i = 0; i < a.nbt; ++i) {
    const int npad = (a.nb[a.bt_term[i]] * 4 + 15) / 16 * 16;
    g.npad[i] = npad; g.col[i] = col; g.qoff[i] = qoff;
    col += npad;
    qoff += 2 * npad * kTcKT * 4;
  }
  if (col > kTcEtaCol) return -1;
  g.qbytes = qoff;
  g.flush_tiles = flush_tiles > 0 ? flush_tiles : 64;
  int K = 0;
  for (int i = 0; i < a.nsc; ++i) {
    h.koff[i] = K;
    K += (a.nb[a.sc_term[i]] + 3) & ~3;
  }
  h.Kp = (K + 7) & ~7;
  const size_t fixed = (size_t)4 * kTcWTile + (size_t)2 * 128 * h.Kp * 4 + (size_t)4 * kTcKT * h.Kp * 4 +
                       (size_t)NS * 4 * kKnotStride * 4 + 128;
  const size_t limit = 227 * 1024 - 256;
  h.qbuf = fixed + 2 * (size_t)g.qbytes <= limit ? 2 : 1;
  const size_t smem = fixed + (size_t)h.qbuf * g.qbytes;
  if (smem > limit) return -1;
  // with single-buffered Q tiles the refill serialises behind the MMAs and version 1 is the
  // faster sweep (measured, profiles/r1_xtwx.json)
  if (h.qbuf == 1 && !allow_single_q) return -1;
  a.partial = partial;
  a.CG = (a.C + 31) / 32;
  g.CB = (a.C + 127) / 128;
  const long long len = a.obs_end > a.obs_begin ? a.obs_end - a.obs_begin : 0;
  long long M = std::max<long long>(1, std::min(num_sms, 160) / g.CB);
  M = std::min<long long>(M, std::max<long long>(1, len / (16LL * kTcKT)));
  long long chunk = std::max<long long>(1, (len + M - 1) / M);
  M = std::max<long long>(1, (len + chunk - 1) / chunk);
  a.M = (int)M;
  a.chunk_len = chunk;
  a.n_items = g.CB * a.M;
  g.a = a;
  cudaError_t e = ensure_smem(k_xtwx_tc2, smem);
  if (e != cudaSuccess) return (int)e;
  const int warps = (kTcWW + NS + 2 + 3) / 4 * 4;
  const int grid = std::min(a.n_items, num_sms);
  k_xtwx_tc2<<<grid, 32 * warps, smem, st>>>(h);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++*launches;
  return 0;
}

}  // namespace ptm
