// ptm_hessian.cuh -- chain-batched OBSERVED information of one coefficient block:
//     info = - jacfwd(grad(logp))        (FlatLogProb.hessian, logprob.py:104-113)
// as CholInfo / PTMCholInfo / PTMCholInfoFixed (iwls_proposals.py:168-184, 227-245, 279-298,
// 333-364), IWLSKernelFixed.init_state (iwls_fixed.py:118-143) and kernel.IWLSKernel._chol_info
// (kernel.py:240-258) factorise it.  Blocks: the coefficients beta_t of one smooth term, or the
// latent transformation coefficients delta_z.
//
// Derivative semantics (restated in oracle/ptm_oracle.py: PTMModel.hessian_block, pinned on the
// reference's code run over nested forward-mode values): the inner grad applies the DECLARED
// custom_jvp rule (approx.py:498-520); the outer jacfwd differentiates the rule BODY as ordinary
// code, so the second-order terms carry the true slopes sH, sD, sD2 of the lerped tables:
//     A_i = -d/dr[dl/dr] = sH d + z sD + [d > tiny] (sD d2 / d^2 - sD2 / d)        (core)
//         = d^2 + z s1 + [d > tiny] (s1 / d)^2,  s1 = dd/dr                        (transitions)
//         = 1                                                                      (tails)
//   beta_t, location term:  info = X' diag(e^{-2 eta_sigma} A) X + K / tau^2
//   beta_t, scale term:     info = X' diag(r^2 A - r dl/dr) X + K / tau^2
//   delta_z:  info = - Z_d' [ J' H_theta J + sum_k (dl/dtheta_k) d2theta_k ] Z_d + I / tau_d^2,
//             H_theta = - sum_i ( b_z b_z' + [d > tiny] b_d b_d' / d^2 ),  b_z = dz/dtheta, b_d = dd/dtheta
// X' W X goes through the same banded identity as the IWLS information (Z' (B' W B) Z).
//
// This is a set-up / per-transition-of-an-optional-kernel quantity, not the MCMC inner loop:
// the sweep is kept simple.  lane = chain; a CTA owns one (chain group, observation chunk); its
// warps are autonomous: each stages 32 observations (one per lane: knot intervals + basis
// values of every term), then walks over them with lane = chain, accumulating into its PRIVATE
// shared-memory table [slot][lane] (double) -- no atomics, fixed summation order.
#pragma once
#include <cuda_runtime.h>

#include "ptm_kernels.cuh"

namespace ptm {

constexpr int kHessBeta = 0, kHessDeltaZ = 2;
constexpr int kHessScal = 12;  // scalar sums of the delta_z block (boundary terms)
enum HessScal { HS_N1L = 0, HS_C1L, HS_C2L, HS_Q2L, HS_N1R, HS_C1R, HS_C2R, HS_Q2R, HS_GVL, HS_GDL, HS_GVR, HS_GDR };

// accumulators per lane of a block kind
inline int hess_nacc(int kind, int bnb, int KK) {
  return kind == kHessDeltaZ ? (KK + 4) * 5 + (KK + 4) + kHessScal : bnb * 4;
}

template <typename R>
struct HessArgs {
  const R* y;
  const R* X;
  const R* knots;
  const R* grid;
  const R* Gamma;   // [T][kMaxNbases][C]
  const R* cc;      // [KK*5+6][C]
  double* partial;  // [n_items][NACC][32]
  long long n, obs_begin, obs_end, chunk_len;
  int C, CG, M, n_items, KK, T, sum_nb, NW, NACC;
  int kind;         // kHessBeta | kHessDeltaZ
  int bterm;        // block term (kHessBeta)
  int nb[kMaxTerms], pred[kMaxTerms], goff[kMaxTerms];
  R inv_dk[kMaxTerms];
  TrafoScal<R> ts;
};

// five piece-coefficient weights -> the five B-spline coefficients vt[kk .. kk+4]
// (transpose of piece_coefs, as in k_main's eval warps)
template <typename R>
__device__ __forceinline__ void piece_to_vt(const R (&w)[5], bool lastp, R (&u)[5]) {
  const R sixth = R(1) / R(6);
  const R w4 = lastp ? R(0) : w[4];
  const R g3 = w[3] - w4;
  u[0] = (w[0] - g3) * sixth + (w[2] - w[1]) * R(0.5);
  u[1] = (R(4) * w[0] - w4) * sixth - w[2] + g3 * R(0.5);
  u[2] = w[0] * sixth + (w[1] + w[2] - g3 + w4) * R(0.5);
  u[3] = g3 * sixth - w4 * R(0.5);
  u[4] = w4 * sixth;
}

template <typename R>
__global__ void __launch_bounds__(256, 1) k_hess(const HessArgs<R> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int T = a.T, KK = a.KK, NW = a.NW, NACC = a.NACC;
  const int NCC = KK * 5 + 6, NGV = KK + 4;
  double* s_acc = reinterpret_cast<double*>(smem_raw);             // [NW][NACC][32]
  R* s_gam = reinterpret_cast<R*>(s_acc + (size_t)NW * NACC * 32);  // [sum_nb][32]
  R* s_cc = s_gam + a.sum_nb * 32;                                 // [NCC + 4][32]
  R* s_kn = s_cc + (NCC + 4) * 32;                                 // [T][4][kKnotStride]
  R* s_st = s_kn + T * 4 * kKnotStride;                            // [NW][T][32][4]
  R* s_y = s_st + NW * T * 32 * 4;                                 // [NW][32]

  for (int t = warp; t < T; t += NW) {
    R* kn = s_kn + t * 4 * kKnotStride;
    const R* gk = a.knots + t * kKnotStride;
    const int nbt = a.nb[t];
    for (int i = lane; i < nbt + 4; i += 32) {
      kn[i] = gk[i];
      kn[kKnotStride + i] = i + 1 < nbt + 4 ? R(1) / (gk[i + 1] - gk[i]) : R(0);
      kn[2 * kKnotStride + i] = i + 2 < nbt + 4 ? R(1) / (gk[i + 2] - gk[i]) : R(0);
      kn[3 * kKnotStride + i] = i + 3 < nbt + 4 ? R(1) / (gk[i + 3] - gk[i]) : R(0);
    }
  }
  const TrafoScal<R> ts = a.ts;
  const R tiny = tiny_of<R>();
  double* acc = s_acc + (size_t)warp * NACC * 32 + lane;
  R* st_w = s_st + (size_t)warp * T * 32 * 4;
  const bool dz = a.kind == kHessDeltaZ;
  const int bgo32 = dz ? 0 : a.goff[a.bterm] * 32;
  const bool bscale = !dz && a.pred[a.bterm] != 0;

  for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
    const int cg = item / a.M, mchunk = item - cg * a.M;
    const long long o_begin = a.obs_begin + (long long)mchunk * a.chunk_len;
    long long o_end = o_begin + a.chunk_len;
    if (o_end > a.obs_end) o_end = a.obs_end;
    int chain = cg * 32 + lane;
    if (chain >= a.C) chain = a.C - 1;
    __syncthreads();  // previous item's tables are no longer read
    for (int t = 0; t < T; ++t)
      for (int j = warp; j < a.nb[t]; j += NW)
        s_gam[(a.goff[t] + j) * 32 + lane] = a.Gamma[((long long)t * kMaxNbases + j) * a.C + chain];
    for (int s = warp; s < NCC; s += NW) s_cc[s * 32 + lane] = a.cc[(long long)s * a.C + chain];
    if (warp == 0) {
      ChainBoundary<R> cb;
      make_boundary(ts, a.cc[(long long)(KK * 5 + 0) * a.C + chain], a.cc[(long long)(KK * 5 + 1) * a.C + chain],
                    a.cc[(long long)(KK * 5 + 2) * a.C + chain], a.cc[(long long)(KK * 5 + 3) * a.C + chain], cb);
      s_cc[(NCC + 0) * 32 + lane] = cb.constL;
      s_cc[(NCC + 1) * 32 + lane] = cb.constR;
      s_cc[(NCC + 2) * 32 + lane] = cb.ztailL;
      s_cc[(NCC + 3) * 32 + lane] = cb.ztailR;
    }
    for (int i = 0; i < NACC; ++i) acc[i * 32] = 0.0;
    double sc[kHessScal];
#pragma unroll
    for (int i = 0; i < kHessScal; ++i) sc[i] = 0.0;
    __syncthreads();
    const R beta0 = s_cc[(KK * 5 + 4) * 32 + lane], gamma0 = s_cc[(KK * 5 + 5) * 32 + lane];

    for (long long b0 = o_begin + (long long)warp * 32; b0 < o_end; b0 += (long long)NW * 32) {
      // ---- stage 32 observations: lane = observation -------------------------------------
      const long long gi = b0 + lane;
      if (gi < o_end) {
        for (int t = 0; t < T; ++t) {
          const R* kn = s_kn + t * 4 * kKnotStride;
          R bb[4];
          const int jj = basis_window_fast(a.X[(long long)t * a.n + gi], kn, kn + kKnotStride, kn + 2 * kKnotStride,
                                           kn + 3 * kKnotStride, a.nb[t], a.inv_dk[t], bb);
          store3j(st_w + (t * 32 + lane) * 4, bb, a.goff[t] * 32 + jj * 32);
        }
        s_y[warp * 32 + lane] = a.y[gi];
      }
      __syncwarp();
      const int cnt = (int)((o_end - b0) < 32 ? (o_end - b0) : 32);
      // ---- walk over them: lane = chain ---------------------------------------------------
      for (int o = 0; o < cnt; ++o) {
        const R yv = s_y[warp * 32 + o];
        R em = beta0, es = gamma0;
        R bb0 = R(0), bb1 = R(0), bb2 = R(0), bb3 = R(0);
        int bjj = 0;
        for (int t = 0; t < T; ++t) {
          R b0_, b1_, b2_, b3_;
          int jo;
          load3j(st_w + (t * 32 + o) * 4, b0_, b1_, b2_, b3_, jo);
          const R* gp = s_gam + lane + jo;
          R f = b0_ * gp[0];
          f = fma(b1_, gp[32], f);
          f = fma(b2_, gp[64], f);
          f = fma(b3_, gp[96], f);
          if (a.pred[t] == 0) em += f; else es += f;
          if (!dz && t == a.bterm) { bb0 = b0_; bb1 = b1_; bb2 = b2_; bb3 = b3_; bjj = (jo - bgo32) >> 5; }
        }
        const R einv = fast_exp(-es);
        const R r = (yv - em) * einv;
        const int code = region_code(ts, r);
        const bool core = code == 2;
        const R rc = r < ts.a ? ts.a : (r > ts.b ? ts.b : (r == r ? r : ts.a));
        CoreEval<R> ev;
        core_locate_fast(ts, a.grid, rc, ev);
        const R* cp = s_cc + (ev.kk * 5) * 32 + lane;
        R sH, sD, sD2;
        core_eval_slopes(ts, cp[0], cp[32], cp[64], cp[96], cp[128], ev, sH, sD, sD2);
        R z = ev.z, d = ev.d, dzdr = ev.d, dddr = ev.d2;
        R cfac = R(0), qfac = R(0), s1 = R(0);
        if (!core) {
          if (code == 1) {
            const R dL = s_cc[(KK * 5 + 1) * 32 + lane], constL = s_cc[(NCC + 0) * 32 + lane];
            const R poly = r * ts.a - R(0.5) * r * r;
            const R t1 = r - poly * ts.inv_w;
            z = (ts.inv_w * poly + dL * t1) + constL;
            const R q = (ts.a - r) * ts.inv_w;
            d = (R(1) - q) * dL + q;
            s1 = (dL - R(1)) * ts.inv_w;
            dzdr = d; dddr = s1;
            const R poly0 = ts.a * ts.a - R(0.5) * ts.a * ts.a;
            cfac = t1 - (ts.a - poly0 * ts.inv_w);
            qfac = R(1) - q;
          } else if (code == 3) {
            const R dR = s_cc[(KK * 5 + 3) * 32 + lane], constR = s_cc[(NCC + 1) * 32 + lane];
            const R poly = R(0.5) * r * r - r * ts.b;
            const R t1 = r - poly * ts.inv_w;
            z = (ts.inv_w * poly + dR * t1) + constR;
            const R q = (r - ts.b) * ts.inv_w;
            d = (R(1) - q) * dR + q;
            s1 = (R(1) - dR) * ts.inv_w;
            dzdr = d; dddr = s1;
            const R poly0 = R(0.5) * ts.b * ts.b - ts.b * ts.b;
            cfac = t1 - (ts.b - poly0 * ts.inv_w);
            qfac = R(1) - q;
          } else if (code == 0) {
            z = s_cc[(NCC + 2) * 32 + lane] - (ts.min_eps - r);
            d = R(1); dzdr = R(1); dddr = R(0);
            cfac = cL_of(ts, ts.min_eps);
          } else {
            z = s_cc[(NCC + 3) * 32 + lane] + (r - ts.max_eps);
            d = R(1); dzdr = R(1); dddr = R(0);
            cfac = cR_of(ts, ts.max_eps);
          }
        }
        const bool above = d > tiny;
        const R inv_d = above ? fast_rcp(d) : R(0);
        if (!dz) {
          // ---- beta_t block: weight of the banded contraction --------------------------------
          R A;
          if (core) A = fma(sH, d, z * sD) + inv_d * (sD * dddr * inv_d - sD2);
          else if (code == 1 || code == 3) A = fma(d, d, z * s1) + (s1 * inv_d) * (s1 * inv_d);
          else A = R(1);
          const R g_r = fma(-z, dzdr, inv_d * dddr);
          R wgt = bscale ? fma(r * r, A, -r * g_r) : einv * einv * A;
          if (!(r == r)) wgt = r;  // NaN response -> NaN information, like the reference
          const R w0 = wgt * bb0, w1 = wgt * bb1, w2 = wgt * bb2, w3 = wgt * bb3;
          double* ap = acc + bjj * 128;  // row jj, [row][diagonal][lane]
          ap[0 * 32] += (double)(w0 * bb0); ap[1 * 32] += (double)(w0 * bb1);
          ap[2 * 32] += (double)(w0 * bb2); ap[3 * 32] += (double)(w0 * bb3);
          ap[4 * 32] += (double)(w1 * bb1); ap[5 * 32] += (double)(w1 * bb2);
          ap[6 * 32] += (double)(w1 * bb3);
          ap[8 * 32] += (double)(w2 * bb2); ap[9 * 32] += (double)(w2 * bb3);
          ap[12 * 32] += (double)(w3 * bb3);
        } else {
          // ---- delta_z block: rows dz/dvt, dd/dvt of the piece (core) or of the boundary ------
          const R lz = -z;
          R wz[5], wd[5], uz[5], ud[5];
          const R on = core ? R(1) : R(0);
          core_backward(ts, ev, on, R(0), wz);
          core_backward(ts, ev, R(0), on * inv_d, wd);  // already divided by d
          const bool lastp = ev.kk + 1 >= KK;
          piece_to_vt(wz, lastp, uz);
          piece_to_vt(wd, lastp, ud);
          double* bp = acc + (ev.kk * 5) * 32;
#pragma unroll
          for (int p = 0; p < 5; ++p)
#pragma unroll
            for (int q = p; q < 5; ++q) bp[((p * 5) + (q - p)) * 32] += (double)fma(uz[p], uz[q], ud[p] * ud[q]);
          double* gp = acc + (NGV * 5 + ev.kk) * 32;
#pragma unroll
          for (int p = 0; p < 5; ++p) gp[p * 32] += (double)fma(lz, uz[p], ud[p]);
          if (!core) {
            const R qd = qfac * inv_d;
            const double dv = (double)fma(lz, cfac, inv_d * qfac);
            if (code <= 1) {
              sc[HS_N1L] += 1.0; sc[HS_C1L] += (double)cfac; sc[HS_C2L] += (double)(cfac * cfac);
              sc[HS_Q2L] += (double)(qd * qd); sc[HS_GVL] += (double)lz; sc[HS_GDL] += dv;
            } else {
              sc[HS_N1R] += 1.0; sc[HS_C1R] += (double)cfac; sc[HS_C2R] += (double)(cfac * cfac);
              sc[HS_Q2R] += (double)(qd * qd); sc[HS_GVR] += (double)lz; sc[HS_GDR] += dv;
            }
            if (!(r == r)) sc[HS_N1L] += (double)r;  // NaN response poisons the block
          }
        }
      }
      __syncwarp();
    }
    if (dz) {
#pragma unroll
      for (int i = 0; i < kHessScal; ++i) acc[(NGV * 6 + i) * 32] = sc[i];
    }
    __syncthreads();
    // fixed-order sum over the warps of the CTA -> this item's partial
    double* pout = a.partial + (long long)item * NACC * 32;
    for (int idx = tid; idx < NACC * 32; idx += blockDim.x) {
      double v = 0.0;
      for (int w = 0; w < NW; ++w) v += s_acc[(size_t)w * NACC * 32 + idx];
      pout[idx] = v;
    }
  }
}

// ---- post-processing shared by both block kinds -------------------------------------------
// Smallest eigenvalue of the symmetric d x d matrix S (row-major, destroyed): cyclic Jacobi,
// one thread (d <= 40; only reached when the matrix is not safely positive definite).
__device__ inline double jacobi_min_eig(double* S, int d) {
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int i = 0; i < d; ++i) {
      dia += S[i * d + i] * S[i * d + i];
      for (int j = i + 1; j < d; ++j) off += S[i * d + j] * S[i * d + j];
    }
    if (!(off > 1e-30 * dia)) break;  // also leaves on NaN
    for (int p = 0; p < d - 1; ++p)
      for (int q = p + 1; q < d; ++q) {
        const double apq = S[p * d + q];
        if (apq == 0.0) continue;
        const double theta = (S[q * d + q] - S[p * d + p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < d; ++k) {
          const double akp = S[k * d + p], akq = S[k * d + q];
          S[k * d + p] = c * akp - s * akq;
          S[k * d + q] = s * akp + c * akq;
        }
        for (int k = 0; k < d; ++k) {
          const double apk = S[p * d + k], aqk = S[q * d + k];
          S[p * d + k] = c * apk - s * aqk;
          S[q * d + k] = s * apk + c * aqk;
        }
      }
  }
  double mn = S[0];
  for (int i = 1; i < d; ++i) mn = S[i * d + i] < mn ? S[i * d + i] : mn;
  return mn;
}

// In-place lower Cholesky of the symmetric matrix A (d x d, shared memory), whole block.
// Returns (in *ok, shared) whether every pivot was > 0.
__device__ inline void block_cholesky(double* A, int d, int* ok) {
  const int tid = threadIdx.x;
  if (tid == 0) *ok = 1;
  __syncthreads();
  for (int j = 0; j < d; ++j) {
    if (tid == 0) {
      const double pv = A[j * d + j];
      if (!(pv > 0.0)) *ok = 0;
      A[j * d + j] = sqrt(pv);
    }
    __syncthreads();
    const double dj = A[j * d + j];
    for (int i = j + 1 + tid; i < d; i += blockDim.x) A[i * d + j] /= dj;
    __syncthreads();
    const int m = d - j - 1;
    for (int idx = tid; idx < m * m; idx += blockDim.x) {
      const int i = j + 1 + idx / m, k = j + 1 + idx % m;
      if (k <= i) A[i * d + k] -= A[i * d + j] * A[k * d + j];
    }
    __syncthreads();
  }
}

// info (d x d, shared, row-major) -> the reference's post-processing by `mode`, outputs.
//   0: hess = info as computed;           chol = cholesky(sym(info))
//   1: make_pd(info)                      (iwls_proposals.py:175-181, 236-242; iwls_fixed.py:129-136)
//   2: make_pd(info + 1e-6 I)             (iwls_proposals.py:291-301, 345-355)
//   3: hess = info;                       chol = cholesky(info + I)        (kernel.py:255-257)
//   4: hess = info + 1e-6 I;              chol = cholesky(hess), the identity when that has NaNs
//                                         (ObservedCholInfoOrIdentity.chol_info_or_identity,
//                                          iwls_proposals.py:357-364)
// make_pd: A = (A + A')/2, A += max(0, 1e-6 - lambda_min(A)) I.
template <typename R>
__device__ inline void hess_finish(double* s_A, double* s_B, int d, int mode, long long c, R* hess, R* chol) {
  const int tid = threadIdx.x;
  __shared__ int s_ok;
  __shared__ double s_shift;
  // symmetrise (jacfwd(grad) is symmetric here up to rounding; make_pd and jnp.linalg.cholesky
  // symmetrise anyway)
  for (int idx = tid; idx < d * d; idx += blockDim.x) {
    const int i = idx / d, k = idx - i * d;
    s_B[idx] = 0.5 * (s_A[i * d + k] + s_A[k * d + i]);
  }
  __syncthreads();
  if (mode == 0 || mode == 3 || mode == 4) {
    for (int idx = tid; idx < d * d; idx += blockDim.x)
      hess[c * d * d + idx] = (R)(s_A[idx] + (mode == 4 && idx / d == idx % d ? 1e-6 : 0.0));
  }
  __syncthreads();
  for (int idx = tid; idx < d * d; idx += blockDim.x) s_A[idx] = s_B[idx];
  __syncthreads();
  if (mode == 2 || mode == 4) for (int i = tid; i < d; i += blockDim.x) s_A[i * d + i] += 1e-6;
  if (mode == 3) for (int i = tid; i < d; i += blockDim.x) s_A[i * d + i] += 1.0;
  __syncthreads();
  if (mode == 1 || mode == 2) {
    // is lambda_min >= jitter?  Cholesky of A - jitter I succeeds exactly then (no shift needed)
    for (int idx = tid; idx < d * d; idx += blockDim.x) {
      const int i = idx / d, k = idx - i * d;
      s_B[idx] = s_A[idx] - (i == k ? 1e-6 : 0.0);
    }
    __syncthreads();
    block_cholesky(s_B, d, &s_ok);
    if (!s_ok) {
      for (int idx = tid; idx < d * d; idx += blockDim.x) s_B[idx] = s_A[idx];
      __syncthreads();
      if (tid == 0) {
        const double w_min = jacobi_min_eig(s_B, d);
        const double sh = 1e-6 - w_min;
        s_shift = sh > 0.0 ? sh : (sh == sh ? 0.0 : sh);
      }
      __syncthreads();
      for (int i = tid; i < d; i += blockDim.x) s_A[i * d + i] += s_shift;
    }
    __syncthreads();
    for (int idx = tid; idx < d * d; idx += blockDim.x) hess[c * d * d + idx] = (R)s_A[idx];
  }
  __syncthreads();
  if (!chol) return;
  block_cholesky(s_A, d, &s_ok);
  const double qnan = __longlong_as_double(0x7ff8000000000000LL);
  for (int idx = tid; idx < d * d; idx += blockDim.x) {
    const int i = idx / d, k = idx - i * d;
    if (!s_ok) chol[c * d * d + idx] = mode == 4 ? (i == k ? R(1) : R(0)) : (R)qnan;  // NaN like jnp.linalg.cholesky
    else chol[c * d * d + idx] = k <= i ? (R)s_A[idx] : R(0);
  }
}

template <typename R>
struct HessPostArgs {
  const R* params;
  const double* partial;
  const R* Zall;
  const R* Kall;
  const R* Zd;
  const TrafoConst<double>* tc;
  R* hess;
  R* chol;
  int C, P, M, NACC, KK, mode, add_priors;
  int term, nbt, p, zoff, koff, tau_off, taud_off, dz_off;
  int nc;  // the block is in non-centred form: coef = scale * latent, prior scale 1
};

// beta_t block: one block (128 threads) per chain.
template <typename R>
__global__ void k_hess_post_beta(const HessPostArgs<R> a) {
  typedef double D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int nbt = a.nbt, p = a.p;
  D* s_band = reinterpret_cast<D*>(smem_raw);  // [nbt*4]
  D* s_mz = s_band + nbt * 4;                  // [nbt][p]
  D* s_P = s_mz + nbt * p;                     // [p][p]
  D* s_Q = s_P + p * p;                        // [p][p] scratch
  const int cg = c >> 5, lane = c & 31;
  for (int i = tid; i < nbt * 4; i += blockDim.x) {
    D v = 0.0;
    const D* pp = a.partial + ((long long)cg * a.M * a.NACC + i) * 32 + lane;
    for (int m = 0; m < a.M; ++m) v += pp[(long long)m * a.NACC * 32];
    s_band[i] = v;
  }
  __syncthreads();
  const R* Zt = a.Zall + a.zoff;
  for (int idx = tid; idx < nbt * p; idx += blockDim.x) {
    const int r = idx / p, k = idx - r * p;
    D v = 0.0;
    for (int dd = -3; dd <= 3; ++dd) {
      const int b = r + dd;
      if (b < 0 || b >= nbt) continue;
      const D mv = dd >= 0 ? s_band[r * 4 + dd] : s_band[b * 4 - dd];
      v = fma(mv, (D)Zt[b * p + k], v);
    }
    s_mz[idx] = v;
  }
  __syncthreads();
  const D tau = (D)a.params[(long long)c * a.P + a.tau_off + a.term];
  // non-centred: d gamma / d latent = tau Z, prior latent ~ N(0, inv(K))
  const D is2 = a.add_priors ? (a.nc ? 1.0 : 1.0 / (tau * tau)) : 0.0;
  const D dsc = a.nc ? tau * tau : 1.0;
  const R* Kt = a.Kall + a.koff;
  for (int idx = tid; idx < p * p; idx += blockDim.x) {
    const int i = idx / p, k = idx - i * p;
    D v = 0.0;
    for (int r = 0; r < nbt; ++r) v = fma((D)Zt[r * p + i], s_mz[r * p + k], v);
    v *= dsc;
    s_P[idx] = a.add_priors ? fma(is2, (D)Kt[idx], v) : v;
  }
  __syncthreads();
  hess_finish<R>(s_P, s_Q, p, a.mode, c, a.hess, a.chol);
}

// delta_z block: one block (64 threads) per chain; everything in double.
template <typename R>
__global__ void k_hess_post_dz(const HessPostArgs<R> a) {
  typedef double D;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int c = blockIdx.x, tid = threadIdx.x;
  const TrafoConst<D>& tc = *a.tc;
  const int KK = a.KK, NGV = KK + 4, np_ = tc.nparam, J = tc.J, eo = tc.e_off, dq = np_ - 1;
  D* s_red = reinterpret_cast<D*>(smem_raw);  // [NACC]
  D* s_M = s_red + a.NACC;                    // [J][J]   vt space, then theta space
  D* s_E = s_M + J * J;                       // [np][np] e space, then delta space
  D* s_W = s_E + np_ * np_;                   // [np][np] scratch
  D* s_v = s_W + np_ * np_;                   // vectors: gvt[J+1], gth[J], ge[np], e[np], sm[np], aux 4 x J
  D* gvt = s_v;
  D* gth = gvt + (J + 1);
  D* ge = gth + J;
  D* ev = ge + np_;
  D* sm = ev + np_;
  D* aVL = sm + np_;
  D* aDL = aVL + (J + 1);
  D* aVR = aDL + (J + 1);
  D* aDR = aVR + (J + 1);
  const int cg = c >> 5, lane = c & 31;
  for (int i = tid; i < a.NACC; i += blockDim.x) {
    D v = 0.0;
    const D* pp = a.partial + ((long long)cg * a.M * a.NACC + i) * 32 + lane;
    for (int m = 0; m < a.M; ++m) v += pp[(long long)m * a.NACC * 32];
    s_red[i] = v;
  }
  __syncthreads();
  const D* sc = s_red + NGV * 6;
  const R* par = a.params + (long long)c * a.P;
  if (tid == 0) {
    // coefficient map at this chain's delta (ptm.py:155-172, 210-230) and the boundary functionals
    D delta[kMaxNparam], vt[kMaxJ];
    const R* dzp = par + a.dz_off;
    for (int j = 0; j < np_; ++j) {
      D v = 0.0;
      for (int k = 0; k < dq; ++k) v = fma((D)a.Zd[j * dq + k], (D)dzp[k], v);
      delta[j] = a.nc ? v * (D)par[a.taud_off] : v;
    }
    coef_map_forward(tc, delta, ev, sm, vt);
    for (int j = 0; j <= J; ++j) { gvt[j] = j < NGV ? s_red[NGV * 5 + j] : 0.0; aVL[j] = aDL[j] = aVR[j] = aDR[j] = 0.0; }
    D g1[5] = {1.0, 0.0, 0.0, 0.0, 0.0}, g2[5] = {0.0, tc.inv_dk, 0.0, 0.0, 0.0};
    D g3[5] = {1.0, 1.0, 1.0, 1.0, 0.0}, g4[5] = {0.0, tc.inv_dk, 2.0 * tc.inv_dk, 3.0 * tc.inv_dk, 0.0};
    if (tc.variant == kTrafoPtm) {  // the onion form is the identity off the core (onion.py:133-138)
      piece_coefs_backward(g1, 0, KK, aVL);
      piece_coefs_backward(g2, 0, KK, aDL);
      piece_coefs_backward(g3, KK - 1, KK, aVR);
      piece_coefs_backward(g4, KK - 1, KK, aDR);
    }
    for (int j = 0; j < J; ++j)
      gvt[j] += sc[HS_GVL] * aVL[j] + sc[HS_GDL] * aDL[j] + sc[HS_GVR] * aVR[j] + sc[HS_GDR] * aDR[j];
    D run = 0.0;
    for (int j = J - 1; j >= 0; --j) { run += gvt[j]; gth[j] = run; }  // vt = cumsum(theta)
    for (int j = 0; j < np_; ++j) ge[j] = gth[j + eo] - tc.Bk[j + eo] * gth[0];
  }
  __syncthreads();
  // M_vt [J][J] from the band (5 diagonals) + boundary outer products
  for (int idx = tid; idx < J * J; idx += blockDim.x) {
    const int i = idx / J, k = idx - i * J;
    const int lo = i < k ? i : k, dd = i < k ? k - i : i - k;
    D v = dd < 5 ? s_red[lo * 5 + dd] : 0.0;
    v += sc[HS_N1L] * aVL[i] * aVL[k] + sc[HS_C1L] * (aVL[i] * aDL[k] + aDL[i] * aVL[k]) +
         (sc[HS_C2L] + sc[HS_Q2L]) * aDL[i] * aDL[k];
    v += sc[HS_N1R] * aVR[i] * aVR[k] + sc[HS_C1R] * (aVR[i] * aDR[k] + aDR[i] * aVR[k]) +
         (sc[HS_C2R] + sc[HS_Q2R]) * aDR[i] * aDR[k];
    s_M[idx] = v;
  }
  __syncthreads();
  // theta space: suffix sums over both indices (vt = cumsum(theta)), rows then columns
  for (int k = tid; k < J; k += blockDim.x) {
    D run = 0.0;
    for (int i = J - 1; i >= 0; --i) { run += s_M[i * J + k]; s_M[i * J + k] = run; }
  }
  __syncthreads();
  for (int i = tid; i < J; i += blockDim.x) {
    D run = 0.0;
    for (int k = J - 1; k >= 0; --k) { run += s_M[i * J + k]; s_M[i * J + k] = run; }
  }
  __syncthreads();
  // e space: theta_0 = k1 - Bk[e_off:] . e, theta_{j + e_off} = e_j
  for (int idx = tid; idx < np_ * np_; idx += blockDim.x) {
    const int j = idx / np_, l = idx - j * np_;
    const D bj = tc.Bk[j + eo], bl = tc.Bk[l + eo];
    s_E[idx] = s_M[(j + eo) * J + (l + eo)] - bj * s_M[0 * J + (l + eo)] - bl * s_M[(j + eo) * J + 0] + bj * bl * s_M[0];
  }
  __syncthreads();
  // delta space: de_j/ddelta_l = e_j (1[j=l] - sm_l).  W = info_e Je, then Je' W
  for (int idx = tid; idx < np_ * np_; idx += blockDim.x) {
    const int j = idx / np_, m = idx - j * np_;
    D rs = 0.0;
    for (int l = 0; l < np_; ++l) rs = fma(s_E[j * np_ + l], ev[l], rs);
    s_W[idx] = s_E[idx] * ev[m] - sm[m] * rs;
  }
  __syncthreads();
  {
    D Q = 0.0;
    for (int j = 0; j < np_; ++j) Q = fma(ge[j], ev[j], Q);
    for (int idx = tid; idx < np_ * np_; idx += blockDim.x) {
      const int l = idx / np_, m = idx - l * np_;
      D cs = 0.0;
      for (int j = 0; j < np_; ++j) cs = fma(ev[j], s_W[j * np_ + m], cs);
      const D info = ev[l] * s_W[l * np_ + m] - sm[l] * cs;  // Je' info_e Je  (= -J' H_e J)
      const D ql = ge[l] * ev[l], qm = ge[m] * ev[m];
      const D t2 = (l == m ? ql - Q * sm[l] : 0.0) - ql * sm[m] - sm[l] * qm + 2.0 * Q * sm[l] * sm[m];
      s_E[idx] = info - t2;  // -H_delta
    }
  }
  __syncthreads();
  // delta_z space: Z_d' (-H_delta) Z_d + I / tau_d^2
  for (int idx = tid; idx < np_ * dq; idx += blockDim.x) {
    const int l = idx / dq, k = idx - l * dq;
    D v = 0.0;
    for (int m = 0; m < np_; ++m) v = fma(s_E[l * np_ + m], (D)a.Zd[m * dq + k], v);
    s_W[idx] = v;
  }
  __syncthreads();
  const D taud = (D)par[a.taud_off];
  for (int idx = tid; idx < dq * dq; idx += blockDim.x) {
    const int i = idx / dq, k = idx - i * dq;
    D v = 0.0;
    for (int l = 0; l < np_; ++l) v = fma((D)a.Zd[l * dq + i], s_W[l * dq + k], v);
    if (a.nc) v *= taud * taud;
    if (a.add_priors && i == k) v += a.nc ? 1.0 : 1.0 / (taud * taud);
    s_M[idx] = v;
  }
  __syncthreads();
  hess_finish<R>(s_M, s_E, dq, a.mode, c, a.hess, a.chol);
}

}  // namespace ptm
