// ptm_xtwx.cuh -- chain-batched IWLS information  X' diag(w) X  (+ penalty, ridge,
// Cholesky) of smooth terms.  Reference: iwls_proposals.py:55-89 (loc terms,
// w_i = 1/max(sigma_i, sqrt(eps))^2) and 118-144 (scale terms, w == 2).
//
// Banded route: X = B Z with B the banded cubic design, so
//     X' W X = Z' (B' W B) Z,
// and B' W B is a 7-diagonal nb x nb matrix that needs only the 10 unique products
// b_a b_b of the four non-zero basis values per observation instead of p^2 (84x less
// arithmetic at p = 29 than the dense contraction).
//
// Same skeleton as k_main: lane = chain, a CTA owns 32 chains and streams a chunk of
// observations in tiles.
//   weight warps  evaluate eta_sigma for the tile (banded gather from the shared
//                 coefficient table of the SCALE terms) and write w = 1/max(sigma,eps)^2
//                 into a double-buffered shared tile [obs][lane];
//   stager warps  (one per scale term) stage knot interval + basis values;
//   band warps    (one per requested term) stage their own term and accumulate the band
//                 rows [row][diagonal][lane] in shared memory -- private per term, so
//                 the read-modify-write needs no atomics.  Several loc terms can be
//                 accumulated in ONE sweep (they share w, which only the scale terms
//                 change).
#pragma once
#include <cuda_runtime.h>

#include "ptm_kernels.cuh"

namespace ptm {

constexpr int kXtMaxItems = 8 * 160;

template <typename R>
size_t xtwx_workspace_bytes(int sum_nb, long long C) {
  const long long CG = (C + 31) / 32;
  return (size_t)(kXtMaxItems + CG) * (size_t)sum_nb * 4 * 32 * sizeof(double);
}

template <typename R>
struct XtwxArgs {
  const R* X;
  const R* knots;
  const R* Gamma;    // [T][kMaxNbases][C] from k_pre
  const R* cc;       // intercept gamma0 at slot KK*5+5
  const R* params;   // [C][P] (tau of the terms)
  const R* Zall;
  const R* Kall;
  double* partial;   // [n_items][band_slots][32]
  long long n, obs_begin, obs_end, chunk_len;
  int C, P, T, KK, CG, M, n_items, To, NW;
  int nsc, nbt, sum_sc_nb, band_slots, const_w, tau_off, add_penalty;
  int nc[kMaxTerms];  // non-centred terms: chol_info = tau * chol (iwls_proposals.py:82-89)
  int sc_term[kMaxTerms], sc_goff[kMaxTerms];  // scale terms feeding sigma
  int bt_term[kMaxTerms], bt_boff[kMaxTerms];  // requested terms, offset of their band block
  int nb[kMaxTerms], p[kMaxTerms], zoff[kMaxTerms], koff[kMaxTerms];
  R inv_dk[kMaxTerms];
};

template <typename R> __device__ __forceinline__ R sqrt_eps();
template <> __device__ __forceinline__ double sqrt_eps<double>() { return 1.4901161193847656e-08; }
template <> __device__ __forceinline__ float sqrt_eps<float>() { return 3.4526698e-04f; }

template <typename R>
__global__ void __launch_bounds__(640, 1) k_xtwx(const XtwxArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int To = a.To, NW = a.NW, nsc = a.nsc, nbt = a.nbt, NS = nsc + nbt;
  double* s_acc = reinterpret_cast<double*>(smem_raw);            // [band_slots][32]
  R* s_gam = reinterpret_cast<R*>(s_acc + a.band_slots * 32);     // [sum_sc_nb][32]
  R* s_w = s_gam + a.sum_sc_nb * 32;                              // [2][To][32]
  R* s_b = s_w + 2 * To * 32;                                     // [NS][3][To][4]
  R* s_kn = s_b + NS * 3 * To * 4;                                // [NS][4][kKnotStride]
  int* s_j = reinterpret_cast<int*>(s_kn + NS * 4 * kKnotStride); // [NS][3][To]

  const bool is_weight = warp < NW;
  const int sidx = warp - NW;                 // staging index: scale terms first, then band terms
  const bool is_stager = !is_weight && sidx < nsc;
  const bool is_band = !is_weight && sidx >= nsc && sidx < NS;
  const int term = is_stager ? a.sc_term[sidx] : (is_band ? a.bt_term[sidx - nsc] : 0);
  int t_nb = 0;
  R t_invdk = R(0);
  const R* t_x = nullptr;
  if (is_stager || is_band) {
    t_nb = a.nb[term];
    t_invdk = a.inv_dk[term];
    t_x = a.X + (long long)term * a.n;
    R* kn = s_kn + sidx * 4 * kKnotStride;
    const R* gk = a.knots + term * kKnotStride;
    for (int i = lane; i < t_nb + 4; i += 32) {
      kn[i] = gk[i];
      kn[kKnotStride + i] = i + 1 < t_nb + 4 ? R(1) / (gk[i + 1] - gk[i]) : R(0);
      kn[2 * kKnotStride + i] = i + 2 < t_nb + 4 ? R(1) / (gk[i + 2] - gk[i]) : R(0);
      kn[3 * kKnotStride + i] = i + 3 < t_nb + 4 ? R(1) / (gk[i + 3] - gk[i]) : R(0);
    }
  }
  __syncthreads();
  const R eps = sqrt_eps<R>();

  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int cg = item / a.M, mchunk = item - cg * a.M;
    const long long o_begin = a.obs_begin + (long long)mchunk * a.chunk_len;
    long long o_end = o_begin + a.chunk_len;
    if (o_end > a.obs_end) o_end = a.obs_end;
    const long long len = o_end > o_begin ? o_end - o_begin : 0;
    const int ntiles = (int)((len + To - 1) / To);
    int chain = cg * 32 + lane;
    if (chain >= a.C) chain = a.C - 1;
    double* pout = a.partial + (long long)item * a.band_slots * 32;

    if (is_stager || is_band) {
      const R* kn = s_kn + sidx * 4 * kKnotStride;
      const int goff32 = is_stager ? a.sc_goff[sidx] * 32 : 0;
      if (is_stager)
        for (int j = 0; j < t_nb; ++j)
          s_gam[(a.sc_goff[sidx] + j) * 32 + lane] = a.Gamma[((long long)term * kMaxNbases + j) * a.C + chain];
      double* acc = s_acc + (is_band ? a.bt_boff[sidx - nsc] : 0) * 32 + lane;
      if (is_band)
        for (int i = 0; i < t_nb * 4; ++i) acc[i * 32] = 0.0;
      auto stage = [&](int tile) {
        const long long base = o_begin + (long long)tile * To;
        const int slot = tile % 3;
        R* bst = s_b + ((sidx * 3 + slot) * To) * 4;
        int* jst = s_j + (sidx * 3 + slot) * To;
        for (int o = lane; o < To; o += 32) {
          const long long gi = base + o;
          if (gi < o_end) {
            R bb[4];
            const int jj = basis_window_fast(t_x[gi], kn, kn + kKnotStride, kn + 2 * kKnotStride,
                                             kn + 3 * kKnotStride, t_nb, t_invdk, bb);
            store4(bst + o * 4, bb);
            jst[o] = is_stager ? goff32 + jj * 32 : jj * 128;  // gamma row offset | band row offset
          }
        }
      };
      if (ntiles > 0) stage(0);
      cta_bar();
      for (int s = 0; s <= ntiles; ++s) {
        if (s + 1 < ntiles) stage(s + 1);
        if (is_band && s >= 1) {
          const int tile = s - 1;
          const long long base = o_begin + (long long)tile * To;
          const int slot = tile % 3;
          const R* bp = s_b + ((sidx * 3 + slot) * To) * 4;
          const int* jp = s_j + (sidx * 3 + slot) * To;
          const R* wp = s_w + ((tile & 1) * To) * 32 + lane;
          const int cnt = (int)((o_end - base) < To ? (o_end - base) : To);
          for (int o = 0; o < cnt; ++o, bp += 4, ++jp, wp += 32) {
            R b0, b1, b2, b3;
            load4(bp, b0, b1, b2, b3);
            const R wgt = *wp;
            const R w0 = wgt * b0, w1 = wgt * b1, w2 = wgt * b2, w3 = wgt * b3;
            double* ap = acc + *jp;  // row jj, [row][diagonal][lane]
            ap[0 * 32] += (double)(w0 * b0); ap[1 * 32] += (double)(w0 * b1);
            ap[2 * 32] += (double)(w0 * b2); ap[3 * 32] += (double)(w0 * b3);
            ap[4 * 32] += (double)(w1 * b1); ap[5 * 32] += (double)(w1 * b2);
            ap[6 * 32] += (double)(w1 * b3);
            ap[8 * 32] += (double)(w2 * b2); ap[9 * 32] += (double)(w2 * b3);
            ap[12 * 32] += (double)(w3 * b3);
          }
        }
        cta_bar();
      }
      if (is_band)
        for (int i = 0; i < t_nb * 4; ++i) pout[(a.bt_boff[sidx - nsc] + i) * 32 + lane] = acc[i * 32];
    } else if (!is_weight) {
      cta_bar();
      for (int s = 0; s <= ntiles; ++s) cta_bar();
    } else {
      // =========================== weight warp =====================================
      const R gamma0 = a.cc[(long long)(a.KK * 5 + 5) * a.C + chain];
      const R* gl = s_gam + lane;
      cta_bar();
      for (int s = 0; s <= ntiles; ++s) {
        if (s < ntiles) {
          const long long base = o_begin + (long long)s * To;
          const int slot = s % 3;
          R* wb = s_w + ((s & 1) * To) * 32 + lane;
          const int cnt = (int)((o_end - base) < To ? (o_end - base) : To);
          for (int o = warp; o < cnt; o += NW) {
            R wgt = R(2);
            if (!a.const_w) {
              R es = gamma0;
              const R* bq = s_b + (slot * To + o) * 4;
              const int* jq = s_j + slot * To + o;
              const int qstep = 3 * To;
#pragma unroll 2
              for (int q = 0; q < nsc; ++q, bq += qstep * 4, jq += qstep) {
                R b0, b1, b2, b3;
                load4(bq, b0, b1, b2, b3);
                const R* gp = gl + *jq;
                R f = b0 * gp[0];
                f = fma(b1, gp[32], f);
                f = fma(b2, gp[64], f);
                f = fma(b3, gp[96], f);
                es += f;
              }
              // w = 1 / clip(sigma, sqrt(eps))^2 with sigma = exp(eta_sigma)
              const R sd = fast_exp(es);
              const R sdc = sd < eps ? eps : sd;  // NaN propagates like jnp.clip
              wgt = fast_rcp(sdc * sdc);
            }
            wb[o * 32] = wgt;
          }
        }
        cta_bar();
      }
    }
    __syncthreads();
  }
}

// Reduce the partial bands of one term over the chunks (fixed order), expand Z' M Z, add
// the penalty and ridge, factorise.  One block (128 threads) per chain.
template <typename R>
__global__ void k_xtwx_post(const XtwxArgs<R> a, int term, int boff, R* __restrict__ xtwx,
                            R* __restrict__ precision, R* __restrict__ chol) {
  typedef double D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int t = term, nbt = a.nb[t], p = a.p[t];
  D* s_band = reinterpret_cast<D*>(smem_raw);  // [nbt*4]
  D* s_mz = s_band + nbt * 4;                  // [nbt][p]  M Z
  D* s_P = s_mz + nbt * p;                     // [p][p]
  __shared__ D s_diag;
  const int cg = c >> 5, lane = c & 31;
  for (int i = tid; i < nbt * 4; i += blockDim.x) {
    D v = 0.0;
    const D* pp = a.partial + ((long long)cg * a.M * a.band_slots + boff + i) * 32 + lane;
    for (int m = 0; m < a.M; ++m) v += pp[(long long)m * a.band_slots * 32];
    s_band[i] = v;
  }
  __syncthreads();
  const R* Zt = a.Zall + a.zoff[t];
  // (M Z)[r][k] = sum_b M[r][b] Z[b][k], M symmetric with band entries M[r][r+d] = band[r*4+d]
  for (int idx = tid; idx < nbt * p; idx += blockDim.x) {
    const int r = idx / p, k = idx - r * p;
    D v = 0.0;
    for (int d = -3; d <= 3; ++d) {
      const int b = r + d;
      if (b < 0 || b >= nbt) continue;
      const D mv = d >= 0 ? s_band[r * 4 + d] : s_band[b * 4 - d];
      v = fma(mv, (D)Zt[b * p + k], v);
    }
    s_mz[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < p * p; idx += blockDim.x) {
    const int i = idx / p, k = idx - i * p;
    D v = 0.0;
    for (int r = 0; r < nbt; ++r) v = fma((D)Zt[r * p + i], s_mz[r * p + k], v);
    s_P[idx] = v;
    if (xtwx) xtwx[((long long)c * p + i) * p + k] = (R)v;
  }
  __syncthreads();
  if (!precision) return;
  const D eps = (D)sqrt_eps<R>();
  const D tau = (D)a.params[(long long)c * a.P + a.tau_off + t];
  const D tc_ = tau < eps ? eps : tau;  // NaN propagates like jnp.clip
  const D is2 = 1.0 / (tc_ * tc_);
  const R* Kt = a.Kall + a.koff[t];
  if (a.add_penalty)
    for (int idx = tid; idx < p * p; idx += blockDim.x) s_P[idx] = fma(is2, (D)Kt[idx], s_P[idx]);
  __syncthreads();
  if (tid == 0) {
    D s = 0.0;
    for (int i = 0; i < p; ++i) s += s_P[i * p + i];
    s_diag = 1e-6 * (s / (D)p);
  }
  __syncthreads();
  if (a.add_penalty)
    for (int i = tid; i < p; i += blockDim.x) s_P[i * p + i] += s_diag;
  __syncthreads();
  for (int idx = tid; idx < p * p; idx += blockDim.x) precision[(long long)c * p * p + idx] = (R)s_P[idx];
  if (!chol) return;
  __syncthreads();
  // right-looking Cholesky in shared memory (p <= 32)
  for (int j = 0; j < p; ++j) {
    if (tid == 0) s_P[j * p + j] = sqrt(s_P[j * p + j]);
    __syncthreads();
    const D dj = s_P[j * p + j];
    for (int i = j + 1 + tid; i < p; i += blockDim.x) s_P[i * p + j] /= dj;
    __syncthreads();
    for (int idx = tid; idx < (p - j - 1) * (p - j - 1); idx += blockDim.x) {
      const int i = j + 1 + idx / (p - j - 1), k = j + 1 + idx % (p - j - 1);
      if (k <= i) s_P[i * p + k] -= s_P[i * p + j] * s_P[k * p + j];
    }
    __syncthreads();
  }
  for (int idx = tid; idx < p * p; idx += blockDim.x) {
    const int i = idx / p, k = idx - i * p;
    chol[(long long)c * p * p + idx] = k <= i ? (R)(a.nc[t] ? tau * s_P[idx] : s_P[idx]) : R(0);
  }
}

// Plans and launches the sweep for the terms listed in a.bt_term[0..nbt); the caller
// runs k_xtwx_post per term afterwards.  Returns a cudaError_t (0 = ok) or -1 when the
// shared-memory budget does not allow this many terms in one sweep.
template <typename R>
int xtwx_sweep(XtwxArgs<R>& a, double* partial, int num_sms, cudaStream_t st, int* launches) {
  a.partial = partial;
  a.CG = (a.C + 31) / 32;
  const int NS = a.nsc + a.nbt;
  a.NW = std::max(4, std::min(12, 20 - NS));
  if (a.NW + NS > 20) return -1;
  a.To = 4 * a.NW;
  auto smem_for = [&](int To) {
    size_t b = 0;
    b += (size_t)a.band_slots * 32 * sizeof(double);
    b += (size_t)a.sum_sc_nb * 32 * sizeof(R);
    b += (size_t)2 * To * 32 * sizeof(R);
    b += (size_t)NS * 3 * To * 4 * sizeof(R);
    b += (size_t)NS * 4 * kKnotStride * sizeof(R);
    b += (size_t)(NS * 3 * To + 8) * sizeof(int);
    return b;
  };
  while (a.To > 8 && smem_for(a.To) > 227 * 1024) a.To -= 4;
  const size_t smem = smem_for(a.To);
  if (smem > 227 * 1024) return -1;
  const long long len = a.obs_end > a.obs_begin ? a.obs_end - a.obs_begin : 0;
  long long M = std::max<long long>(1, (8LL * std::min(num_sms, 160)) / a.CG);
  M = std::min<long long>(M, std::max<long long>(1, len / ((long long)a.To * 16)));
  long long chunk = std::max<long long>(1, (len + M - 1) / M);
  M = std::max<long long>(1, (len + chunk - 1) / chunk);
  a.M = (int)M;
  a.chunk_len = chunk;
  a.n_items = a.CG * a.M;
  cudaError_t e = ensure_smem(k_xtwx<R>, smem);
  if (e != cudaSuccess) return (int)e;
  const int warps = (a.NW + NS + 3) / 4 * 4;
  const int grid = std::min(a.n_items, num_sms);
  k_xtwx<R><<<grid, 32 * warps, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++*launches;
  return 0;
}

template <typename R>
int xtwx_post(const XtwxArgs<R>& a, int term, int boff, R* xtwx, R* precision, R* chol, cudaStream_t st,
              int* launches) {
  const int nbt = a.nb[term], p = a.p[term];
  const size_t smem2 = ((size_t)nbt * 4 + (size_t)nbt * p + (size_t)p * p) * sizeof(double);
  k_xtwx_post<R><<<a.C, 128, smem2, st>>>(a, term, boff, xtwx, precision, chol);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++*launches;
  return 0;
}

}  // namespace ptm
