// ptm_math.cuh -- scalar building blocks of the PTM hot path, usable from host
// and device.  Everything here is chain-independent arithmetic on plain pointers;
// the kernels in ptm_kernels.cu decide the HBM / shared-memory layout around it.
//
// Reference (paths under /root/reference/liesel_ptm/):
//   find_interval / bspline4      bspline/approx.py:74-107 (order-0 indicator and
//                                 the Cox-de Boor recursion, restricted to the four
//                                 non-zero columns; same operation order, no FMA
//                                 contraction, so values are bit-identical to an
//                                 IEEE evaluation of the reference's formula)
//   coef_map_forward/backward     bspline/ptm.py:92-122,155-172,210-230 + var.py:106-115
//   piece_coefs                   the cubic pieces of sum_j B_j(x) (S theta)_j, i.e. of
//                                 approx.py:385-388's tables G theta, G' theta, G'' theta
//   trafo_eval                    approx.py:418-452 (grid lerp with step=(b-a)/1000 on a
//                                 999-interval grid), ptm.py:349-478 (regions)
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PTM_HD __host__ __device__ __forceinline__
#else
#define PTM_HD inline
#endif

namespace ptm {

constexpr int kMaxNbases = 32;
constexpr int kMaxNparam = 40;   // shape parameters; J = nparam + 1 <= 41 (onion form: nparam + 7 <= 41)
constexpr int kTrafoPtm = 0, kTrafoOnion = 1, kTrafoIdentity = 2;  // model/model.py:98-131
constexpr int kMaxJ = kMaxNparam + 1;
constexpr int kGridN = 1000;     // BSplineApprox ngrid (ptm.py:312)

// ---- round-to-nearest arithmetic without FMA contraction ---------------------
template <typename R> struct RN;
template <> struct RN<double> {
  static PTM_HD double mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
  }
  static PTM_HD double add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
  }
  static PTM_HD double sub(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
  }
  static PTM_HD double div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
  }
};
template <> struct RN<float> {
  static PTM_HD float mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    return a * b;
#endif
  }
  static PTM_HD float add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    return a + b;
#endif
  }
  static PTM_HD float sub(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fsub_rn(a, b);
#else
    return a - b;
#endif
  }
  static PTM_HD float div(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
  }
};

template <typename R> PTM_HD R tiny_of();
template <> PTM_HD double tiny_of<double>() { return 2.2250738585072014e-308; }
template <> PTM_HD float tiny_of<float>() { return 1.17549435e-38f; }

// ---- banded cubic B-spline basis ----------------------------------------------
// j with knots[j] <= x < knots[j+1], 3 <= j <= nb (approx.py:88-91).  The caller
// guarantees knots[3] <= x <= knots[nb] (the reference raises otherwise).
template <typename R>
PTM_HD int find_interval(R x, const R* knots, int nb, R inv_dk) {
  int j = 3 + (int)floor((double)((x - knots[3]) * inv_dk));
  j = j < 3 ? 3 : (j > nb ? nb : j);
  while (j > 3 && x < knots[j]) --j;
  while (j < nb && x >= knots[j + 1]) ++j;
  return j;
}

// Four non-zero cubic values of columns j-3..j at x (approx.py:93-101).
template <typename R>
PTM_HD void bspline4(R x, const R* k, int j, R out[4]) {
  typedef RN<R> A;
  // order 1
  const R d10 = A::sub(k[j + 1], k[j]);
  const R n1a = A::div(A::sub(k[j + 1], x), d10);  // column j-1
  const R n1b = A::div(A::sub(x, k[j]), d10);      // column j
  // order 2
  const R d2a = A::sub(k[j + 1], k[j - 1]);
  const R d2b = A::sub(k[j + 2], k[j]);
  const R n2a = A::mul(A::div(A::sub(k[j + 1], x), d2a), n1a);  // column j-2
  const R n2b = A::add(A::mul(A::div(A::sub(x, k[j - 1]), d2a), n1a),
                       A::mul(A::div(A::sub(k[j + 2], x), d2b), n1b));  // column j-1
  const R n2c = A::mul(A::div(A::sub(x, k[j]), d2b), n1b);  // column j
  // order 3
  const R d3a = A::sub(k[j + 1], k[j - 2]);
  const R d3b = A::sub(k[j + 2], k[j - 1]);
  const R d3c = A::sub(k[j + 3], k[j]);
  out[0] = A::mul(A::div(A::sub(k[j + 1], x), d3a), n2a);
  out[1] = A::add(A::mul(A::div(A::sub(x, k[j - 2]), d3a), n2a),
                  A::mul(A::div(A::sub(k[j + 2], x), d3b), n2b));
  out[2] = A::add(A::mul(A::div(A::sub(x, k[j - 1]), d3b), n2b),
                  A::mul(A::div(A::sub(k[j + 3], x), d3c), n2c));
  out[3] = A::mul(A::div(A::sub(x, k[j]), d3c), n2c);
}

// Same four values with the knot-difference reciprocals tabulated (r1[j] =
// 1/(k[j+1]-k[j]), r2[j] = 1/(k[j+2]-k[j]), r3[j] = 1/(k[j+3]-k[j])): no divisions
// in the fused kernels' staging.  Agrees with bspline4 to a few ulp.
template <typename R>
PTM_HD void bspline4_fast(R x, const R* k, const R* r1, const R* r2, const R* r3, int j, R out[4]) {
  const R xm2 = x - k[j - 2], xm1 = x - k[j - 1], x0 = x - k[j];
  const R p1 = k[j + 1] - x, p2 = k[j + 2] - x, p3 = k[j + 3] - x;
  const R n1a = p1 * r1[j], n1b = x0 * r1[j];
  const R a2 = r2[j - 1] * n1a, b2 = r2[j] * n1b;
  const R n2a = p1 * a2;
  const R n2b = fma(xm1, a2, p2 * b2);
  const R n2c = x0 * b2;
  const R a3 = r3[j - 2] * n2a, b3 = r3[j - 1] * n2b, c3 = r3[j] * n2c;
  out[0] = p1 * a3;
  out[1] = fma(xm2, a3, p2 * b3);
  out[2] = fma(xm1, b3, p3 * c3);
  out[3] = x0 * c3;
}

template <typename R>
PTM_HD int basis_window_fast(R x, const R* knots, const R* r1, const R* r2, const R* r3, int nb,
                             R inv_dk, R out[4]) {
  const int j = find_interval(x, knots, nb, inv_dk);
  bspline4_fast(x, knots, r1, r2, r3, j, out);
  if (j == nb) {
    out[3] = out[2];
    out[2] = out[1];
    out[1] = out[0];
    out[0] = R(0);
    return nb - 4;
  }
  return j - 3;
}

// The fused kernels index gamma[jj .. jj+3] with jj = j - 3 in [0, nb - 4].  At the
// last knot (j == nb) column j does not exist and its value is 0, so shift the
// window one to the left.
template <typename R>
PTM_HD int basis_window(R x, const R* knots, int nb, R inv_dk, R out[4]) {
  const int j = find_interval(x, knots, nb, inv_dk);
  bspline4(x, knots, j, out);
  if (j == nb) {
    out[3] = out[2];
    out[2] = out[1];
    out[1] = out[0];
    out[0] = R(0);
    return nb - 4;
  }
  return j - 3;
}

// ---- transformation constants (chain independent) ---------------------------------
// Scalars every evaluation needs (passed by value to the kernels).
template <typename R>
struct TrafoScal {
  int KK;          // J - 3 cubic pieces on [a, b]
  R a, b;          // knots[3], knots[J]
  R inv_dk;        // 1 / mean knot step
  R w;             // transition width eps*(b-a)  (ptm.py:330)
  R inv_w;
  R min_eps, max_eps;
  R inv_h;         // (ngrid-1)/(b-a): cells of the 1000-point grid
  R kap_scale;     // ngrid/(ngrid-1): kappa = frac * kap_scale (step quirk, approx.py:376)
  R ratio;         // KK/(ngrid-1): grid spacing in knot-step units
  R frac_eps;      // below this distance from a grid point the exact table decides
};

// Scalars + the small arrays of the coefficient map (lives in HBM, read by the
// per-chain prologue / epilogue kernels only).
template <typename R>
struct TrafoConst : TrafoScal<R> {
  int nparam;      // shape parameters
  int J;           // bases of h: nparam + 1 (ptm.py:55-57), nparam + 7 for the onion knots (onion.py:50-59)
  int e_off;       // theta index of the first free increment: 1 (ptm.py:214-218), 4 (onion.py:94, 108)
  int variant;     // kTrafoPtm / kTrafoOnion / kTrafoIdentity
  R dk;            // mean knot step (ptm.py:208)
  R log_dk;
  R k1;            // theta_0 offset: knots[3] (ptm.py:201), knots[2] for the onion form (onion.py:108)
  R log_KK;        // log(J - 3) (ptm.py:122); log(nparam) for the onion form (onion.py:79, 89-90)
  R Bk[kMaxJ];     // (B(k1) S) row, ptm.py:202-205 (zero for the onion form)
  R wts[kMaxNparam];  // log_sfn weights 1/6, 5/6, 1, ..., 5/6, 1/6 (ones for the onion form)
};

// ---- coefficient map delta -> theta -> cumulative B-spline coefficients ---------
// theta[0] = k1 - Bk[e_off:] . e, then e_off - 1 fixed increments dk, e = exp(delta - lsfn + log dk),
// and e_off - 1 fixed increments dk again (none for the PTM form, three for the onion form,
// onion.py:81-108); vt = cumsum(theta) are the B-spline coefficients of h on the core.
template <typename R>
PTM_HD void coef_map_forward(const TrafoConst<R>& tc, const R* delta_in, R* e, R* sm, R* vt) {
  const int np_ = tc.nparam, eo = tc.e_off;
  // Gaussian model (dist.py:766-850): the shape parameters are ignored, h is the spline of delta = 0
  R delta[kMaxNparam];
  for (int j = 0; j < np_; ++j) delta[j] = tc.variant == kTrafoIdentity ? R(0) : delta_in[j];
  R mx = delta[0];
  for (int j = 1; j < np_; ++j) mx = delta[j] > mx ? delta[j] : mx;
  R ssum = R(0);
  for (int j = 0; j < np_; ++j) {
    sm[j] = tc.wts[j] * exp(delta[j] - mx);
    ssum += sm[j];
  }
  const R lsfn = log(ssum) + mx - tc.log_KK;
  const R inv = R(1) / ssum;
  R dot = R(0);
  for (int j = 0; j < np_; ++j) {
    sm[j] *= inv;
    e[j] = exp(delta[j] - lsfn + tc.log_dk);
    dot += tc.Bk[j + eo] * e[j];
  }
  R run = tc.k1 - dot;  // theta[0]
  int i = 0;
  vt[i++] = run;
  for (int j = 1; j < eo; ++j) { run += tc.dk; vt[i++] = run; }
  for (int j = 0; j < np_; ++j) { run += e[j]; vt[i++] = run; }
  for (int j = 1; j < eo; ++j) { run += tc.dk; vt[i++] = run; }
}

// Monomial coefficients (local coordinate u in [0,1)) of piece kk from the uniform
// cubic B-spline coefficients vt[kk .. kk+3]; c[4] = jump of the cubic coefficient
// to the next piece (0 for the last piece).
template <typename R>
PTM_HD void piece_coefs(const R* vt, int kk, int KK, R c[5]) {
  const R p0 = vt[kk], p1 = vt[kk + 1], p2 = vt[kk + 2], p3 = vt[kk + 3];
  const R sixth = R(1) / R(6);
  c[0] = (p0 + R(4) * p1 + p2) * sixth;
  c[1] = (p2 - p0) * R(0.5);
  c[2] = (p0 - R(2) * p1 + p2) * R(0.5);
  c[3] = (-p0 + R(3) * p1 - R(3) * p2 + p3) * sixth;
  if (kk + 1 < KK) {
    const R p4 = vt[kk + 4];
    const R c3n = (-p1 + R(3) * p2 - R(3) * p3 + p4) * sixth;
    c[4] = c3n - c[3];
  } else {
    c[4] = R(0);
  }
}

// Transpose of piece_coefs: accumulate gc[5] of piece kk into gvt.
template <typename R>
PTM_HD void piece_coefs_backward(const R* gc, int kk, int KK, R* gvt) {
  const R sixth = R(1) / R(6);
  // c3 of this piece enters the jump (c[4] = c3_next - c3) with a minus sign
  const R g3 = (kk + 1 < KK) ? gc[3] - gc[4] : gc[3];
  gvt[kk] += gc[0] * sixth - gc[1] * R(0.5) + gc[2] * R(0.5) - g3 * sixth;
  gvt[kk + 1] += gc[0] * R(4) * sixth - gc[2] + g3 * R(0.5);
  gvt[kk + 2] += gc[0] * sixth + gc[1] * R(0.5) + gc[2] * R(0.5) - g3 * R(0.5);
  gvt[kk + 3] += g3 * sixth;
  if (kk + 1 < KK) {
    const R gj = gc[4];  // d c3(next piece)
    gvt[kk + 1] += -gj * sixth;
    gvt[kk + 2] += gj * R(0.5);
    gvt[kk + 3] += -gj * R(0.5);
    gvt[kk + 4] += gj * sixth;
  }
}

// gvt (gradient w.r.t. cumulative coefficients) -> gradient w.r.t. delta.
template <typename R>
PTM_HD void coef_map_backward(const TrafoConst<R>& tc, const R* gvt, const R* e,
                              const R* sm, R* gdelta) {
  const int np_ = tc.nparam, eo = tc.e_off;
  // reverse cumsum: gtheta[j] = sum_{l >= j} gvt[l]
  R run = R(0);
  R gth[kMaxJ];
  for (int j = tc.J - 1; j >= 0; --j) {
    run += gvt[j];
    gth[j] = run;
  }
  R tot = R(0);
  for (int j = 0; j < np_; ++j) {
    const R ge = gth[j + eo] - tc.Bk[j + eo] * gth[0];
    gdelta[j] = e[j] * ge;
    tot += gdelta[j];
  }
  for (int j = 0; j < np_; ++j) gdelta[j] -= sm[j] * tot;
}

// ---- transformation evaluation --------------------------------------------------
// Boundary values of one chain (ptm.py:417-419) and the constants derived from them.
template <typename R>
struct ChainBoundary {
  R vL, dL, vR, dR;
  R constL, constR;    // ptm.py:362-365, 386-389
  R ztailL, ztailR;    // transition value at min_eps / max_eps (ptm.py:435, 440)
};

template <typename R>
PTM_HD void make_boundary(const TrafoScal<R>& tc, R vL, R dL, R vR, R dR,
                          ChainBoundary<R>& cb) {
  cb.vL = vL; cb.dL = dL; cb.vR = vR; cb.dR = dR;
  const R a = tc.a, b = tc.b, iw = tc.inv_w;
  const R poly0L = a * a - R(0.5) * a * a;
  cb.constL = vL - (iw * poly0L + dL * (a - poly0L * iw));
  const R poly0R = R(0.5) * b * b - b * b;
  cb.constR = vR - (iw * poly0R + dR * (b - poly0R * iw));
  const R xe = tc.min_eps;
  const R polyE = xe * a - R(0.5) * xe * xe;
  cb.ztailL = (iw * polyE + dL * (xe - polyE * iw)) + cb.constL;
  const R xf = tc.max_eps;
  const R polyF = R(0.5) * xf * xf - xf * b;
  cb.ztailR = (iw * polyF + dR * (xf - polyF * iw)) + cb.constR;
}

// d(z)/d(dL) in the left transition at r, and its right-hand twin.
template <typename R>
PTM_HD R cL_of(const TrafoScal<R>& tc, R r) {
  const R a = tc.a, iw = tc.inv_w;
  const R poly = r * a - R(0.5) * r * r;
  const R poly0 = a * a - R(0.5) * a * a;
  return (r - poly * iw) - (a - poly0 * iw);
}
template <typename R>
PTM_HD R cR_of(const TrafoScal<R>& tc, R r) {
  const R b = tc.b, iw = tc.inv_w;
  const R poly = R(0.5) * r * r - r * b;
  const R poly0 = R(0.5) * b * b - b * b;
  return (r - poly * iw) - (b - poly0 * iw);
}

// Region code exactly as ptm.py:447-460 (NaN -> 4).
template <typename R>
PTM_HD int region_code(const TrafoScal<R>& tc, R r) {
  if (r >= tc.a && r <= tc.b) return 2;
  if (r < tc.min_eps) return 0;
  if (r < tc.a) return 1;
  if (r < tc.max_eps) return 3;
  return 4;
}

// Grid cell of a core point: m = (number of grid points <= r) - 2 in [0, 999] and
// frac = (r - grid[m+1]) * inv_h in [0, 1).  The multiply-and-floor estimate is
// accepted unless r sits within frac_eps of a grid point; then the exact table
// (BSplineApprox.grid as the reference built it) decides, which keeps the index
// bit-exact with jnp.searchsorted(grid, r, side="right") - 1 == m + 1.
template <typename R>
PTM_HD void grid_cell(const TrafoScal<R>& tc, const R* grid, R r, int& m, R& frac) {
  const R t = (r - tc.a) * tc.inv_h;
  R fm = floor(t);
  fm = fm < R(0) ? R(0) : (fm > R(kGridN - 1) ? R(kGridN - 1) : fm);
  frac = t - fm;
  m = (int)fm;
  if (frac < tc.frac_eps || frac > R(1) - tc.frac_eps) {
    int i = m + 1;  // grid[i] is the candidate left end
    if (r < grid[i]) --i;
    else if (i < kGridN && r >= grid[i + 1]) ++i;
    if (i < 1) i = 1;
    m = i - 1;
    frac = (r - grid[i]) * tc.inv_h;
  }
}

// Core evaluation: lerp between the spline samples at grid points m and m+1 of
// the piece that holds grid point m.  With P the cubic of piece kk in u, s the
// grid spacing in knot units and jmp the jump of the cubic coefficient at the next
// knot, the right sample is P(u+s) + jmp*max(u+s-1,0)^3 (C2 continuity), which is
// expanded around u so that only one set of five coefficients is touched.
// Branch-light variant for the fused sweep: the argument is already clamped to [a, b],
// so floor(t) is in [0, 999] without clamping; everything stays in floating point
// except the piece index.  The exact table is consulted only within frac_eps of a grid
// point (a rare branch), which keeps the cell bit-exact with the reference.
template <typename R> struct CoreEval;
template <typename R>
PTM_HD void core_locate_fast(const TrafoScal<R>& tc, const R* grid, R rc, CoreEval<R>& ev);

template <typename R>
struct CoreEval {
  R z, d, d2;        // lerped G theta, G' theta, G'' theta rows (approx.py:429-452)
  R u, kappa, mm;    // local coordinate, lerp weight, overshoot past the next knot
  int kk;
};

template <typename R>
PTM_HD void core_locate(const TrafoScal<R>& tc, int m, R frac, CoreEval<R>& ev) {
  const R p = R(m) * tc.ratio;
  R fk = floor(p);
  fk = fk > R(tc.KK - 1) ? R(tc.KK - 1) : fk;
  ev.kk = (int)fk;
  ev.u = p - fk;
  ev.kappa = frac * tc.kap_scale;
  const R over = ev.u + tc.ratio - R(1);
  ev.mm = over > R(0) ? over : R(0);
}

// The exact table decides the cell of a point within frac_eps of a grid point (bit-exact with
// jnp.searchsorted(grid, r, side="right") - 1).  The three candidate grid values are loaded side by
// side: one memory latency instead of a chain of three.
template <typename R>
PTM_HD void grid_cell_exact(const TrafoScal<R>& tc, const R* __restrict__ grid, R rc, R& fm, R& frac) {
  int m = (int)fm;
  m = m < 0 ? 0 : (m > kGridN - 1 ? kGridN - 1 : m);
  int i = m + 1;  // grid[i] is the candidate left end; the table has kGridN + 2 entries
  const R gm = grid[i - 1], g0 = grid[i], g1 = grid[i + 1];
  R gl = g0;
  if (rc < g0) {
    if (i > 1) { --i; gl = gm; }
  } else if (i < kGridN && rc >= g1) {
    ++i; gl = g1;
  }
  fm = R(i - 1);
  frac = (rc - gl) * tc.inv_h;
}

template <typename R>
PTM_HD void core_locate_fast(const TrafoScal<R>& tc, const R* grid, R rc, CoreEval<R>& ev) {
  const R t = (rc - tc.a) * tc.inv_h;
  R fm = floor(t);
  R frac = t - fm;
  if (frac < tc.frac_eps || frac > R(1) - tc.frac_eps) grid_cell_exact(tc, grid, rc, fm, frac);
  const R p = fm * tc.ratio;
  R fk = floor(p);
  fk = fk > R(tc.KK - 1) ? R(tc.KK - 1) : fk;
  ev.kk = (int)fk;
  ev.u = p - fk;
  ev.kappa = frac * tc.kap_scale;
  const R over = ev.u + tc.ratio - R(1);
  ev.mm = over > R(0) ? over : R(0);
}

template <typename R>
PTM_HD void core_eval(const TrafoScal<R>& tc, R c0, R c1, R c2, R c3, R jmp,
                      CoreEval<R>& ev) {
  const R u = ev.u, s = tc.ratio, kap = ev.kappa, mm = ev.mm;
  // Horner with synthetic division: P, P', P''/2
  const R b2 = fma(u, c3, c2);
  const R b1 = fma(u, b2, c1);
  const R P = fma(u, b1, c0);
  const R d1 = fma(u, c3, b2);
  const R Pp = fma(u, d1, b1);
  const R e1 = fma(u, c3, d1);  // P''/2
  const R mm2 = mm * mm;
  const R j3 = R(3) * jmp;
  // value: H[m+1]-H[m] = s(P' + s(P''/2 + s c3)) + jmp mm^3
  const R dH = fma(jmp, mm2 * mm, s * fma(s, fma(s, c3, e1), Pp));
  ev.z = fma(kap, dH, P);
  // first derivative (times dk): D[m+1]-D[m] = s(2 e1 + 3 s c3) + 3 jmp mm^2
  const R dD = fma(j3, mm2, s * fma(R(3) * s, c3, R(2) * e1));
  ev.d = fma(kap, dD, Pp) * tc.inv_dk;
  // second derivative (times dk^2): D2[m+1]-D2[m] = 6 s c3 + 6 jmp mm
  const R dD2 = fma(R(2) * j3, mm, R(6) * s * c3);
  ev.d2 = fma(kap, dD2, R(2) * e1) * (tc.inv_dk * tc.inv_dk);
}

// core_eval plus the TRUE slopes in r of the three lerps (what an ordinary derivative of the
// lerp sees: (T[m+1] - T[m]) / step with step = (b-a)/ngrid, i.e. d kappa / d r = inv_h *
// kap_scale) -- the second-order terms of jacfwd(grad(logp)), see ptm_hessian.cuh.
template <typename R>
PTM_HD void core_eval_slopes(const TrafoScal<R>& tc, R c0, R c1, R c2, R c3, R jmp, CoreEval<R>& ev,
                             R& sH, R& sD, R& sD2) {
  const R u = ev.u, s = tc.ratio, kap = ev.kappa, mm = ev.mm;
  const R b2 = fma(u, c3, c2);
  const R b1 = fma(u, b2, c1);
  const R P = fma(u, b1, c0);
  const R d1 = fma(u, c3, b2);
  const R Pp = fma(u, d1, b1);
  const R e1 = fma(u, c3, d1);
  const R mm2 = mm * mm;
  const R j3 = R(3) * jmp;
  const R dH = fma(jmp, mm2 * mm, s * fma(s, fma(s, c3, e1), Pp));
  ev.z = fma(kap, dH, P);
  const R dD = fma(j3, mm2, s * fma(R(3) * s, c3, R(2) * e1));
  ev.d = fma(kap, dD, Pp) * tc.inv_dk;
  const R dD2 = fma(R(2) * j3, mm, R(6) * s * c3);
  ev.d2 = fma(kap, dD2, R(2) * e1) * (tc.inv_dk * tc.inv_dk);
  const R dk_dr = tc.inv_h * tc.kap_scale;
  sH = dH * dk_dr;
  sD = dD * tc.inv_dk * dk_dr;
  sD2 = dD2 * (tc.inv_dk * tc.inv_dk) * dk_dr;
}

// Weights of dl/dz = lz and dl/dd = ld on the five coefficients of the piece.
template <typename R>
PTM_HD void core_backward(const TrafoScal<R>& tc, const CoreEval<R>& ev, R lz, R ld,
                          R w[5]) {
  const R u = ev.u, s = tc.ratio, kap = ev.kappa, mm = ev.mm;
  const R bq = ld * tc.inv_dk;
  const R u2 = u * u;
  const R phi1 = fma(kap, s, u);
  const R phi2 = fma(kap, fma(u, R(2) * s, s * s), u2);
  const R phi3 = fma(kap, fma(u2, R(3) * s, fma(u, R(3) * s * s, s * s * s)), u2 * u);
  const R mm2 = mm * mm;
  w[0] = lz;
  w[1] = fma(lz, phi1, bq);
  w[2] = fma(lz, phi2, R(2) * bq * phi1);
  w[3] = fma(lz, phi3, R(3) * bq * phi2);
  w[4] = kap * fma(lz, mm2 * mm, R(3) * bq * mm2);
}

}  // namespace ptm
