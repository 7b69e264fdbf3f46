// ptm_xtwx.cuh -- chain-batched IWLS information  X' diag(w) X  (+ penalty, ridge,
// Cholesky) for one smooth term.  Reference: iwls_proposals.py:55-89 (loc terms,
// w_i = 1/max(sigma_i, sqrt(eps))^2) and 118-144 (scale terms, w == 2).
//
// Banded route: X = B Z with B the banded cubic design, so
//     X' W X = Z' (B' W B) Z,
// and B' W B is a 7-diagonal nb x nb matrix that needs only the 10 unique products
// b_a b_b of the four non-zero basis values per observation instead of p^2.
// lane = chain; each warp owns private shared-memory accumulators [nb*4][32]
// (row a, diagonal offset d = b - a) so the scatter is conflict- and atomic-free.
#pragma once
#include <cuda_runtime.h>

#include "ptm_kernels.cuh"

namespace ptm {

constexpr int kXtWarps = 4;
constexpr int kXtMaxItems = 8 * 160;

template <typename R>
size_t xtwx_workspace_bytes(int max_nb, long long C) {
  const long long CG = (C + 31) / 32;
  return (size_t)(kXtMaxItems + CG) * max_nb * 4 * 32 * sizeof(R);
}

template <typename R>
struct XtwxArgs {
  const R* X;
  const R* knots;
  const R* Gamma;   // [T][kMaxNbases][C] from k_pre
  const R* cc;      // intercept gamma0 at slot KK*5+5
  const R* params;  // [C][P] (tau of the term)
  const R* Zall;
  const R* Kall;
  R* partial;       // [n_items][nb*4][32]
  long long n, obs_begin, obs_end, chunk_len;
  int C, P, T, term, KK, CG, M, n_items;
  int tau_off;
  int add_penalty;
  int nb[kMaxTerms], p[kMaxTerms], pred[kMaxTerms], zoff[kMaxTerms], koff[kMaxTerms];
  R inv_dk[kMaxTerms];
};

template <typename R> __device__ __forceinline__ R sqrt_eps();
template <> __device__ __forceinline__ double sqrt_eps<double>() { return 1.4901161193847656e-08; }
template <> __device__ __forceinline__ float sqrt_eps<float>() { return 3.4526698e-04f; }

template <typename R>
__global__ void __launch_bounds__(kXtWarps * 32) k_xtwx(const XtwxArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbt = a.nb[a.term];
  const bool is_loc = a.pred[a.term] == 0;
  // scale terms feed sigma
  int sc_terms[kMaxTerms];
  int nsc = 0;
  if (is_loc)
    for (int t = 0; t < a.T; ++t)
      if (a.pred[t] == 1) sc_terms[nsc++] = t;
  const int nstage = nsc + 1;  // staged terms per observation: scale terms + target
  R* s_gam = reinterpret_cast<R*>(smem_raw);                       // [nsc][kMaxNbases][32]
  R* s_acc = s_gam + nsc * kMaxNbases * 32;                        // [kXtWarps][nbt*4][32]
  R* s_kn = s_acc + kXtWarps * nbt * 4 * 32;                       // [nstage][kKnotStride]
  R* s_b = s_kn + nstage * kKnotStride;                            // [kXtWarps][32][nstage][4]
  int* s_j = reinterpret_cast<int*>(s_b + kXtWarps * 32 * nstage * 4);  // [kXtWarps][32][nstage]
  for (int i = tid; i < nstage * kKnotStride; i += blockDim.x) {
    const int q = i / kKnotStride;
    const int t = q < nsc ? sc_terms[q] : a.term;
    s_kn[i] = a.knots[t * kKnotStride + (i - q * kKnotStride)];
  }
  __syncthreads();
  const R eps = sqrt_eps<R>();
  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int cg = item / a.M, mchunk = item - cg * a.M;
    const long long o_begin = a.obs_begin + (long long)mchunk * a.chunk_len;
    long long o_end = o_begin + a.chunk_len;
    if (o_end > a.obs_end) o_end = a.obs_end;
    int chain = cg * 32 + lane;
    if (chain >= a.C) chain = a.C - 1;
    for (int i = warp; i < nsc * kMaxNbases; i += kXtWarps) {
      const int q = i / kMaxNbases, j = i - q * kMaxNbases;
      s_gam[i * 32 + lane] = a.Gamma[((long long)sc_terms[q] * kMaxNbases + j) * a.C + chain];
    }
    R* acc = s_acc + warp * nbt * 4 * 32 + lane;
    for (int i = 0; i < nbt * 4; ++i) acc[i * 32] = R(0);
    const R gamma0 = a.cc[(long long)(a.KK * 5 + 5) * a.C + chain];
    __syncthreads();
    R* wb = s_b + warp * 32 * nstage * 4;
    int* wj = s_j + warp * 32 * nstage;
    for (long long base = o_begin + warp * 32; base < o_end; base += kXtWarps * 32) {
      // stage 32 observations: lane o evaluates the banded basis of every needed term
      {
        const long long gi = base + lane;
        if (gi < o_end) {
          for (int q = 0; q < nstage; ++q) {
            const int t = q < nsc ? sc_terms[q] : a.term;
            R bb[4];
            const int jj = basis_window(a.X[(long long)t * a.n + gi], s_kn + q * kKnotStride, a.nb[t],
                                        a.inv_dk[t], bb);
            R* bp = wb + (lane * nstage + q) * 4;
            bp[0] = bb[0]; bp[1] = bb[1]; bp[2] = bb[2]; bp[3] = bb[3];
            wj[lane * nstage + q] = jj;
          }
        }
      }
      __syncwarp();
      const int cnt = (int)((o_end - base) < 32 ? (o_end - base) : 32);
      for (int o = 0; o < cnt; ++o) {
        R wgt;
        if (is_loc) {
          R es = gamma0;
          for (int q = 0; q < nsc; ++q) {
            const R* bp = wb + (o * nstage + q) * 4;
            const int jj = wj[o * nstage + q];
            const R* gp = s_gam + (q * kMaxNbases + jj) * 32 + lane;
            es = fma(bp[0], gp[0], es);
            es = fma(bp[1], gp[32], es);
            es = fma(bp[2], gp[64], es);
            es = fma(bp[3], gp[96], es);
          }
          const R sd = exp(es);
          const R sdc = sd > eps ? sd : eps;  // jnp.clip(scale, a_min=eps)
          wgt = R(1) / (sdc * sdc);
        } else {
          wgt = R(2);
        }
        const R* bp = wb + (o * nstage + nsc) * 4;
        const int jj = wj[o * nstage + nsc];
        const R b0 = bp[0], b1 = bp[1], b2 = bp[2], b3 = bp[3];
        const R w0 = wgt * b0, w1 = wgt * b1, w2 = wgt * b2, w3 = wgt * b3;
        R* ap = acc + (jj * 4) * 32;
        ap[0 * 32] += w0 * b0; ap[1 * 32] += w0 * b1; ap[2 * 32] += w0 * b2; ap[3 * 32] += w0 * b3;
        ap[4 * 32] += w1 * b1; ap[5 * 32] += w1 * b2; ap[6 * 32] += w1 * b3;
        ap[8 * 32] += w2 * b2; ap[9 * 32] += w2 * b3;
        ap[12 * 32] += w3 * b3;
      }
      __syncwarp();
    }
    __syncthreads();
    R* pout = a.partial + (long long)item * nbt * 4 * 32;
    for (int idx = tid; idx < nbt * 4 * 32; idx += blockDim.x) {
      R v = R(0);
      for (int w = 0; w < kXtWarps; ++w) v += s_acc[w * nbt * 4 * 32 + idx];
      pout[idx] = v;
    }
    __syncthreads();
  }
}

// Reduce partial bands over the chunks (fixed order), expand Z' M Z, add the
// penalty and ridge, factorise.  One block (128 threads) per chain.
template <typename R>
__global__ void k_xtwx_post(const XtwxArgs<R> a, R* __restrict__ xtwx, R* __restrict__ precision,
                            R* __restrict__ chol) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int t = a.term, nbt = a.nb[t], p = a.p[t];
  R* s_band = reinterpret_cast<R*>(smem_raw);  // [nbt*4]
  R* s_mz = s_band + nbt * 4;                  // [nbt][p]  M Z
  R* s_P = s_mz + nbt * p;                     // [p][p]
  __shared__ R s_diag;
  const int cg = c >> 5, lane = c & 31;
  for (int i = tid; i < nbt * 4; i += blockDim.x) {
    R v = R(0);
    const R* pp = a.partial + ((long long)cg * a.M * nbt * 4 + i) * 32 + lane;
    for (int m = 0; m < a.M; ++m) v += pp[(long long)m * nbt * 4 * 32];
    s_band[i] = v;
  }
  __syncthreads();
  const R* Zt = a.Zall + a.zoff[t];
  // (M Z)[r][k] = sum_b M[r][b] Z[b][k], M symmetric with band entries M[r][r+d] = band[r*4+d]
  for (int idx = tid; idx < nbt * p; idx += blockDim.x) {
    const int r = idx / p, k = idx - r * p;
    R v = R(0);
    for (int d = -3; d <= 3; ++d) {
      const int b = r + d;
      if (b < 0 || b >= nbt) continue;
      const R mv = d >= 0 ? s_band[r * 4 + d] : s_band[b * 4 - d];
      v = fma(mv, Zt[b * p + k], v);
    }
    s_mz[idx] = v;
  }
  __syncthreads();
  for (int idx = tid; idx < p * p; idx += blockDim.x) {
    const int i = idx / p, k = idx - i * p;
    R v = R(0);
    for (int r = 0; r < nbt; ++r) v = fma(Zt[r * p + i], s_mz[r * p + k], v);
    s_P[idx] = v;
    if (xtwx) xtwx[((long long)c * p + i) * p + k] = v;
  }
  __syncthreads();
  if (!precision) return;
  const R eps = sqrt_eps<R>();
  const R tau = a.params[(long long)c * a.P + a.tau_off + t];
  const R tc_ = tau > eps ? tau : eps;
  const R is2 = R(1) / (tc_ * tc_);
  const R* Kt = a.Kall + a.koff[t];
  if (a.add_penalty)
    for (int idx = tid; idx < p * p; idx += blockDim.x) s_P[idx] = fma(is2, Kt[idx], s_P[idx]);
  __syncthreads();
  if (tid == 0) {
    R s = R(0);
    for (int i = 0; i < p; ++i) s += s_P[i * p + i];
    s_diag = R(1e-6) * (s / R(p));
  }
  __syncthreads();
  if (a.add_penalty)
    for (int i = tid; i < p; i += blockDim.x) s_P[i * p + i] += s_diag;
  __syncthreads();
  for (int idx = tid; idx < p * p; idx += blockDim.x) precision[(long long)c * p * p + idx] = s_P[idx];
  if (!chol) return;
  __syncthreads();
  // right-looking Cholesky in shared memory (p <= 32)
  for (int j = 0; j < p; ++j) {
    if (tid == 0) s_P[j * p + j] = sqrt(s_P[j * p + j]);
    __syncthreads();
    const R dj = s_P[j * p + j];
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
    chol[(long long)c * p * p + idx] = k <= i ? s_P[idx] : R(0);
  }
}

template <typename R>
int xtwx_launch(XtwxArgs<R>& a, R* partial, R* xtwx, R* precision, R* chol, int num_sms,
                cudaStream_t st, int* launches) {
  const int t = a.term, nbt = a.nb[t], p = a.p[t];
  a.partial = partial;
  a.CG = (a.C + 31) / 32;
  const long long len = a.obs_end > a.obs_begin ? a.obs_end - a.obs_begin : 0;
  long long M = std::max<long long>(1, (8LL * std::min(num_sms, 160)) / a.CG);
  M = std::min<long long>(M, std::max<long long>(1, len / (kXtWarps * 32 * 8)));
  long long chunk = (len + M - 1) / M;
  chunk = std::max<long long>(1, chunk);
  M = std::max<long long>(1, (len + chunk - 1) / chunk);
  a.M = (int)M;
  a.chunk_len = chunk;
  a.n_items = a.CG * a.M;
  int nsc = 0;
  if (a.pred[t] == 0)
    for (int q = 0; q < a.T; ++q) nsc += a.pred[q] == 1;
  const int nstage = nsc + 1;
  size_t smem = 0;
  smem += (size_t)nsc * kMaxNbases * 32 * sizeof(R);
  smem += (size_t)kXtWarps * nbt * 4 * 32 * sizeof(R);
  smem += (size_t)nstage * kKnotStride * sizeof(R);
  smem += (size_t)kXtWarps * 32 * nstage * 4 * sizeof(R);
  smem += (size_t)kXtWarps * 32 * nstage * sizeof(int);
  cudaError_t e = cudaFuncSetAttribute(k_xtwx<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  const int grid = std::min(a.n_items, 2 * num_sms);
  k_xtwx<R><<<grid, kXtWarps * 32, smem, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++*launches;
  const size_t smem2 = ((size_t)nbt * 4 + (size_t)nbt * p + (size_t)p * p) * sizeof(R);
  k_xtwx_post<R><<<a.C, 128, smem2, st>>>(a, xtwx, precision, chol);
  e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  ++*launches;
  return 0;
}

}  // namespace ptm
