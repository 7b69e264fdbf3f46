// ptm_predict.cuh -- callers of the hot path on NEW data:
//
//   k_predict   z = h((y - mu) / sigma), log-density and CDF for every posterior sample and
//               every new observation (LocScalePTM.init_dist / summarise_dist,
//               model/model.py:698-790 -> dist.py:185-188, 179-181, 339-395, 718-740), or, in
//               quantile mode, y_q = mu + sigma * h^{-1}(Phi^{-1}(u))  (dist.py:203-209,
//               575-650).
//   k_inverse   h^{-1}(z) for raw shape coefficients (TransformationSpline.dot_inverse_n,
//               bspline/util.py:105-122).
//
// Layout: one sample (chain state) per blockIdx.y; the sample's spline coefficients
// gamma_t = Z_t beta_t and the cubic pieces of h come from k_pre's chain-fastest tables and
// are staged in shared memory once per block; lanes run over the observations, so the
// [S][m] outputs are written coalesced.  These sweeps are bound by their output stream
// (up to 3 words written per observation and sample against 1 + T words read, which stay in
// L2 across samples).
//
// The inverse is EXACT for the reference's forward function: in the core that function is
// the lerp of the spline samples H[i] = h(grid[i]) (approx.py:418-433), so the largest i
// with H[i] <= z and kappa = (z - H[i]) / (H[i+1] - H[i]) give r = grid[i] + kappa * step;
// the transitions are quadratics in the distance from the boundary knot (ptm.py:349-394),
// the tails are linear (396-410).  The reference itself interpolates a 1000-point table of
// (h(x), x) pairs with interpax (util/inverse_interpax.py:60-113; third party, absent here)
// and its tests accept |h^{-1}(h(x)) - x| <= 1e-5 .. 1e-4 (tests/test_bspline/test_ptm.py:
// 167-301); the exact inverse is inside that band.
#pragma once
#include "ptm_math.cuh"
#include "ptm_kernels.cuh"  // fast_exp

namespace ptm {

template <typename R>
struct PredictArgs {
  const R* y;      // [m] new responses (value mode) or probabilities u (quantile mode)
  const R* X;      // [T][m] new covariates, SoA
  long long m;
  const R* knots;  // [T][kKnotStride]
  const R* Gamma;  // [T][kMaxNbases][S]  (k_pre)
  const R* cc;     // [KK*5+6][S]         (k_pre)
  const R* grid;   // the reference's 1002-point grid
  int S, T, KK;
  int nb[kMaxTerms], pred[kMaxTerms], goff[kMaxTerms];
  R inv_dk[kMaxTerms];
  int sum_nb;
  TrafoScal<R> ts;
  R* out_z;        // [S][m] each, nullable
  R* out_logpdf;
  R* out_cdf;
  R* out_q;        // quantile mode: [S][m]
  long long u_ld;  // quantile mode: 0 = one probability per row (u [m]); m = one per (sample, row)
                   // (u [S][m]: predictive draws, TransformationDist._sample_n, dist.py:193-201)
};

// h(r), h'(r) with the chain's pieces in shared memory [KK*5] (ptm.py:412-478).
template <typename R>
__device__ __forceinline__ void trafo_forward(const TrafoScal<R>& ts, const ChainBoundary<R>& cb,
                                              const R* pieces, const R* grid, R rv, R& zv, R& dv) {
  const int code = region_code(ts, rv);
  if (code == 2) {
    int mi; R frac;
    grid_cell(ts, grid, rv, mi, frac);
    CoreEval<R> ev;
    core_locate(ts, mi, frac, ev);
    const R* cp = pieces + ev.kk * 5;
    core_eval(ts, cp[0], cp[1], cp[2], cp[3], cp[4], ev);
    zv = ev.z; dv = ev.d;
  } else if (code == 1) {
    const R poly = rv * ts.a - R(0.5) * rv * rv;
    zv = (ts.inv_w * poly + cb.dL * (rv - poly * ts.inv_w)) + cb.constL;
    const R q = (ts.a - rv) * ts.inv_w;
    dv = (R(1) - q) * cb.dL + q;
  } else if (code == 3) {
    const R poly = R(0.5) * rv * rv - rv * ts.b;
    zv = (ts.inv_w * poly + cb.dR * (rv - poly * ts.inv_w)) + cb.constR;
    const R q = (rv - ts.b) * ts.inv_w;
    dv = (R(1) - q) * cb.dR + q;
  } else if (code == 0) {
    zv = cb.ztailL - (ts.min_eps - rv);
    dv = R(1);
  } else {
    zv = cb.ztailR + (rv - ts.max_eps);
    dv = R(1);
  }
}

// Spline samples on the 1000 core grid points: H[i] = h(grid[i+1]) exactly as the forward
// evaluates a point that sits on the grid (kappa = 0).
template <typename R>
__device__ __forceinline__ void fill_grid_samples(const TrafoScal<R>& ts, const R* pieces, R* H) {
  for (int i = threadIdx.x; i < kGridN; i += blockDim.x) {
    CoreEval<R> ev;
    core_locate(ts, i, R(0), ev);
    const R* cp = pieces + ev.kk * 5;
    core_eval(ts, cp[0], cp[1], cp[2], cp[3], cp[4], ev);
    H[i] = ev.z;
  }
}

template <typename R>
__device__ __forceinline__ R trafo_inverse(const TrafoScal<R>& ts, const ChainBoundary<R>& cb,
                                           const R* H, const R* grid, R z) {
  if (!(z == z)) return z;
  if (z < cb.ztailL) return ts.min_eps - (cb.ztailL - z);
  if (z > cb.ztailR) return ts.max_eps + (z - cb.ztailR);
  if (z < cb.vL) {  // left transition: vL - z = dL t + (1 - dL) t^2 / (2 w), t = a - r
    const R c = cb.vL - z;
    const R disc = fma(R(2) * (R(1) - cb.dL) * ts.inv_w, c, cb.dL * cb.dL);
    const R t = R(2) * c / (cb.dL + sqrt(disc > R(0) ? disc : R(0)));
    const R r = ts.a - t;
    return r < ts.min_eps ? ts.min_eps : r;
  }
  if (z > cb.vR) {  // right transition: z - vR = dR t + (1 - dR) t^2 / (2 w), t = r - b
    const R c = z - cb.vR;
    const R disc = fma(R(2) * (R(1) - cb.dR) * ts.inv_w, c, cb.dR * cb.dR);
    const R t = R(2) * c / (cb.dR + sqrt(disc > R(0) ? disc : R(0)));
    const R r = ts.b + t;
    return r > ts.max_eps ? ts.max_eps : r;
  }
  int lo = 0, hi = kGridN - 1;  // H[lo] <= z <= H[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (H[mid] <= z) lo = mid; else hi = mid;
  }
  const R den = H[lo + 1] - H[lo];
  R kappa = den > R(0) ? (z - H[lo]) / den : R(0);
  kappa = kappa < R(0) ? R(0) : (kappa > R(1) ? R(1) : kappa);
  // kappa = (r - grid[i]) / step with step = (b - a) / 1000 = 1 / (inv_h * kap_scale)
  R r = grid[lo + 1] + kappa / (ts.inv_h * ts.kap_scale);
  return r > ts.b ? ts.b : r;
}

__device__ __forceinline__ double std_normal_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }
__device__ __forceinline__ float std_normal_cdf(float z) { return 0.5f * erfcf(-z * 0.70710678f); }
__device__ __forceinline__ double std_normal_quantile(double u) { return normcdfinv(u); }
__device__ __forceinline__ float std_normal_quantile(float u) { return normcdfinvf(u); }

// blockIdx.y = sample; blockIdx.x strides over the observations.
template <typename R, bool QUANTILE>
__global__ void __launch_bounds__(256) k_predict(const PredictArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  R* s_gam = reinterpret_cast<R*>(smem_raw);  // [sum_nb]
  R* s_pc = s_gam + a.sum_nb;                 // [KK*5]
  R* s_H = s_pc + a.KK * 5;                   // [kGridN] (quantile mode)
  __shared__ ChainBoundary<R> s_cb;
  __shared__ R s_icpt[2];
  const int s = blockIdx.y;
  const int KK = a.KK;
  for (int t = 0; t < a.T; ++t)
    for (int j = threadIdx.x; j < a.nb[t]; j += blockDim.x)
      s_gam[a.goff[t] + j] = a.Gamma[((long long)t * kMaxNbases + j) * a.S + s];
  for (int k = threadIdx.x; k < KK * 5; k += blockDim.x) s_pc[k] = a.cc[(long long)k * a.S + s];
  if (threadIdx.x == 0) {
    make_boundary(a.ts, a.cc[(long long)(KK * 5 + 0) * a.S + s], a.cc[(long long)(KK * 5 + 1) * a.S + s],
                  a.cc[(long long)(KK * 5 + 2) * a.S + s], a.cc[(long long)(KK * 5 + 3) * a.S + s], s_cb);
    s_icpt[0] = a.cc[(long long)(KK * 5 + 4) * a.S + s];
    s_icpt[1] = a.cc[(long long)(KK * 5 + 5) * a.S + s];
  }
  __syncthreads();
  if constexpr (QUANTILE) {
    fill_grid_samples(a.ts, s_pc, s_H);
    __syncthreads();
  }
  const TrafoScal<R> ts = a.ts;
  const ChainBoundary<R> cb = s_cb;
  const R beta0 = s_icpt[0], gamma0 = s_icpt[1];
  const R nan = sizeof(R) == 8 ? (R)__longlong_as_double(0x7ff8000000000000LL) : (R)__int_as_float(0x7fc00000);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.m;
       i += (long long)gridDim.x * blockDim.x) {
    R em = beta0, es = gamma0;
    bool inside = true;
    for (int t = 0; t < a.T; ++t) {
      const R x = a.X[(long long)t * a.m + i];
      const R* kn = a.knots + t * kKnotStride;
      const int nb = a.nb[t];
      // Basis.bspline raises outside [knots[3], knots[nb]] (approx.py:197-206); the host
      // wrapper raises the same error up front, NaN is the device-side safety net
      inside = inside && x >= kn[3] && x <= kn[nb];
      R bb[4];
      const R xs = (x >= kn[3] && x <= kn[nb]) ? x : kn[3];
      const int jj = basis_window(xs, kn, nb, a.inv_dk[t], bb);
      const R* g = s_gam + a.goff[t] + jj;
      R f = bb[0] * g[0];
      f = fma(bb[1], g[1], f);
      f = fma(bb[2], g[2], f);
      f = fma(bb[3], g[3], f);
      if (a.pred[t] == 0) em += f; else es += f;
    }
    const long long oi = (long long)s * a.m + i;
    if constexpr (QUANTILE) {
      const R u = a.y[(long long)s * a.u_ld + i];
      const R zq = std_normal_quantile(u);
      const R rq = trafo_inverse(ts, cb, s_H, a.grid, zq);
      a.out_q[oi] = inside ? fma(exp(es), rq, em) : nan;
    } else {
      const R yv = a.y[i];
      const R sd = exp(es);
      const R r = (yv - em) / sd;  // dist.py:735-736
      R z, d;
      trafo_forward(ts, cb, s_pc, a.grid, r, z, d);
      if (!(r == r) || !inside) { z = nan; d = nan; }
      if (a.out_z) a.out_z[oi] = z;
      if (a.out_logpdf) {
        const R tiny = tiny_of<R>();
        const R dcl = d > tiny ? d : (d == d ? tiny : d);
        a.out_logpdf[oi] = (R(-0.5) * z * z - R(0.91893853320467274178)) + (log(dcl) - es);
      }
      if (a.out_cdf) a.out_cdf[oi] = std_normal_cdf(z);
    }
  }
}

// Spline samples H[s][kGridN] of every posterior sample (quantile mode of k_predict_obs).
template <typename R>
__global__ void k_grid_samples(const R* __restrict__ cc, int S, int KK, const TrafoScal<R> ts,
                               R* __restrict__ H) {
  __shared__ R s_pc[kMaxJ * 5];
  const int s = blockIdx.x;
  for (int k = threadIdx.x; k < KK * 5; k += blockDim.x) s_pc[k] = cc[(long long)k * S + s];
  __syncthreads();
  for (int i = threadIdx.x; i < kGridN; i += blockDim.x) {
    CoreEval<R> ev;
    core_locate(ts, i, R(0), ev);
    const R* cp = s_pc + ev.kk * 5;
    core_eval(ts, cp[0], cp[1], cp[2], cp[3], cp[4], ev);
    H[(long long)s * kGridN + i] = ev.z;
  }
}

// Observation-major predictive sweep: thread = new observation, its T basis windows are
// evaluated ONCE and stay in registers; the block then walks over the samples in groups of
// kPredSG whose gamma tables / cubic pieces are staged in shared memory.  blockIdx.y splits
// the samples when there are too few observation blocks to fill the GPU.
constexpr int kPredSG = 4;       // samples per shared-memory stage
constexpr int kPredThreads = 128;

template <typename R, bool QUANTILE, int TT>
__global__ void __launch_bounds__(kPredThreads) k_predict_obs(const PredictArgs<R> a, const R* __restrict__ Hall,
                                                              int s_per_block) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int KK = a.KK;
  const int NCC = KK * 5 + 6;
  R* s_gam = reinterpret_cast<R*>(smem_raw);  // [kPredSG][sum_nb]
  R* s_cc = s_gam + kPredSG * a.sum_nb;       // [kPredSG][NCC]
  __shared__ ChainBoundary<R> s_cb[kPredSG];
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = i < a.m;
  const TrafoScal<R> ts = a.ts;
  const R nan = sizeof(R) == 8 ? (R)__longlong_as_double(0x7ff8000000000000LL) : (R)__int_as_float(0x7fc00000);

  // ---- chain-independent part, once per observation --------------------------------
  R bb[TT][4];
  int joff[TT];
  bool inside = true;
#pragma unroll
  for (int t = 0; t < TT; ++t) {
    joff[t] = 0;
    bb[t][0] = bb[t][1] = bb[t][2] = bb[t][3] = R(0);
    if (t < a.T && live) {
      const R x = a.X[(long long)t * a.m + i];
      const R* kn = a.knots + t * kKnotStride;
      const int nb = a.nb[t];
      const bool ok = x >= kn[3] && x <= kn[nb];
      inside = inside && ok;
      const int jj = basis_window(ok ? x : kn[3], kn, nb, a.inv_dk[t], bb[t]);
      joff[t] = a.goff[t] + jj;
    }
  }
  const R yv = live ? a.y[i] : R(0);
  R zq = R(0);
  if constexpr (QUANTILE) zq = std_normal_quantile(yv);

  const int s_begin = blockIdx.y * s_per_block;
  const int s_end = min(a.S, s_begin + s_per_block);
  for (int s0 = s_begin; s0 < s_end; s0 += kPredSG) {
    const int ns = min(kPredSG, s_end - s0);
    __syncthreads();  // the previous group's tables are no longer read
    for (int idx = threadIdx.x; idx < ns * a.sum_nb; idx += blockDim.x) {
      const int g = idx / a.sum_nb, k = idx - g * a.sum_nb;
      int t = 0;
      while (t + 1 < a.T && k >= a.goff[t + 1]) ++t;
      s_gam[g * a.sum_nb + k] = a.Gamma[((long long)t * kMaxNbases + (k - a.goff[t])) * a.S + s0 + g];
    }
    for (int idx = threadIdx.x; idx < ns * NCC; idx += blockDim.x) {
      const int g = idx / NCC, k = idx - g * NCC;
      s_cc[g * NCC + k] = a.cc[(long long)k * a.S + s0 + g];
    }
    __syncthreads();
    if (threadIdx.x < ns) {
      const R* c = s_cc + threadIdx.x * NCC + KK * 5;
      make_boundary(ts, c[0], c[1], c[2], c[3], s_cb[threadIdx.x]);
    }
    __syncthreads();
    if (!live) continue;
    for (int g = 0; g < ns; ++g) {
      const R* gam = s_gam + g * a.sum_nb;
      const R* pc = s_cc + g * NCC;
      R em = pc[KK * 5 + 4], es = pc[KK * 5 + 5];
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        if (t < a.T) {
          const R* gp = gam + joff[t];
          R f = bb[t][0] * gp[0];
          f = fma(bb[t][1], gp[1], f);
          f = fma(bb[t][2], gp[2], f);
          f = fma(bb[t][3], gp[3], f);
          if (a.pred[t] == 0) em += f; else es += f;
        }
      }
      const long long oi = (long long)(s0 + g) * a.m + i;
      if constexpr (QUANTILE) {
        if (a.u_ld) zq = std_normal_quantile(a.y[(long long)(s0 + g) * a.u_ld + i]);
        const R rq = trafo_inverse(ts, s_cb[g], Hall + (long long)(s0 + g) * kGridN, a.grid, zq);
        a.out_q[oi] = inside ? fma(exp(es), rq, em) : nan;
      } else {
        const R r = (yv - em) * fast_exp(-es);  // dist.py:735-736, as in the fused sweep
        R z, d;
        trafo_forward(ts, s_cb[g], pc, a.grid, r, z, d);
        if (!(r == r) || !inside) { z = nan; d = nan; }
        if (a.out_z) a.out_z[oi] = z;
        if (a.out_logpdf) {
          const R tiny = tiny_of<R>();
          const R dcl = d > tiny ? d : (d == d ? tiny : d);
          a.out_logpdf[oi] = (R(-0.5) * z * z - R(0.91893853320467274178)) + (log(dcl) - es);
        }
        if (a.out_cdf) a.out_cdf[oi] = std_normal_cdf(z);
      }
    }
  }
}

// One block per chain: pieces from raw shape coefficients (as k_transform), then the exact
// inverse for every z.
template <typename R>
__global__ void k_inverse(const R* __restrict__ delta, int C, const R* __restrict__ z, long long m,
                          const TrafoConst<R>* tcp, const R* __restrict__ grid, R* __restrict__ r_out) {
  __shared__ R s_c[(kMaxJ) * 5];
  __shared__ R s_H[kGridN];
  __shared__ ChainBoundary<R> s_cb;
  const int c = blockIdx.x;
  const TrafoConst<R>& tc = *tcp;
  if (threadIdx.x == 0) {
    R e[kMaxNparam], sm[kMaxNparam], vt[kMaxJ];
    coef_map_forward(tc, delta + (long long)c * tc.nparam, e, sm, vt);
    R cl[5] = {0, 0, 0, 0, 0}, cf[5] = {0, 0, 0, 0, 0};
    for (int kk = 0; kk < tc.KK; ++kk) {
      R cq[5];
      piece_coefs(vt, kk, tc.KK, cq);
      for (int q = 0; q < 5; ++q) s_c[kk * 5 + q] = cq[q];
      if (kk == 0) for (int q = 0; q < 5; ++q) cf[q] = cq[q];
      if (kk == tc.KK - 1) for (int q = 0; q < 5; ++q) cl[q] = cq[q];
    }
    make_boundary<R>(tc, cf[0], cf[1] * tc.inv_dk, ((cl[0] + cl[1]) + cl[2]) + cl[3],
                     (cl[1] + R(2) * cl[2] + R(3) * cl[3]) * tc.inv_dk, s_cb);
  }
  __syncthreads();
  const TrafoScal<R> ts = tc;
  fill_grid_samples(ts, s_c, s_H);
  __syncthreads();
  const ChainBoundary<R> cb = s_cb;
  for (long long i = threadIdx.x; i < m; i += blockDim.x)
    r_out[(long long)c * m + i] = trafo_inverse(ts, cb, s_H, grid, z[i]);
}

}  // namespace ptm
