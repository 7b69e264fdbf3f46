// ptm_capi.cu -- C ABI (include/ptm_b200.h) over the kernels in ptm_kernels.cuh.
// Host side only: problem set-up (copies the static data into HBM once), launch
// planning, and the extern "C" entry points.  No torch types, no exceptions across
// the boundary, no allocation or synchronisation inside the per-call entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/ptm_b200.h"
#include "ptm_kernels.cuh"
#include "ptm_xtwx.cuh"
#include "ptm_xtwx_tc.cuh"
#include "ptm_predict.cuh"
#include "ptm_hessian.cuh"
#include "ptm_main_tc.cuh"

namespace {

thread_local std::string g_err;
thread_local int g_launches = 0;
thread_local int g_profile = 0;
thread_local cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define PTM_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t e_ = (call);                                                        \
    if (e_ != cudaSuccess)                                                          \
      return fail(PTM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
  } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Plan {
  int CG, M, n_items, grid, NSLOT, To, NU;
  int TPW, NTW;  // smooth terms per term warp, term warps
  long long chunk_len;
  size_t smem_main;
  size_t off_gamma, off_cc, off_partial, off_xt, off_scratch, off_G, off_red, total;
  // tensor-core route of the float sweep (ptm_main_tc.cuh): its own chunking of the observations
  bool tc = false;
  // launches of the tensor-core sweep: term masks of the eta-only groups (one predictor each, <= 224
  // coefficients) followed by the main group; ext_mask = predictors with an external part
  int tc_npass = 0;
  unsigned tc_terms[8] = {0};
  int tc_ext_mask = 0;
  int tc_NT = 0, tc_NSW = 0, tc_M = 0, tc_CB = 0;
  size_t off_E = 0;
  long long tc_chunk = 0;
  size_t tc_smem_fwd = 0, tc_smem_bwd = 0;
};

}  // namespace

struct PtmProblem {
  int dtype = PTM_F64;
  long long n = 0;
  int T = 0, T_loc = 0, T_scale = 0;
  int nparam = 0, J = 0, KK = 0, P = 0;
  int variant = 0;  // PTM_TRAFO_*
  int nb[ptm::kMaxTerms] = {0}, p[ptm::kMaxTerms] = {0}, pred[ptm::kMaxTerms] = {0};
  int rank[ptm::kMaxTerms] = {0}, qidx[ptm::kMaxTerms] = {0};
  int nc[ptm::kMaxTerms] = {0}, nc_dz = 0;  // non-centred parameterisations
  int zoff[ptm::kMaxTerms] = {0}, koff[ptm::kMaxTerms] = {0}, boff[ptm::kMaxTerms] = {0};
  int goff[ptm::kMaxTerms] = {0};
  double inv_dk[ptm::kMaxTerms] = {0};
  int sum_nb = 0, max_nb = 0, max_p = 0;
  int dz_off = 0, tau_off = 0, taud_off = 0;
  long long obs_begin = 0, obs_end = 0;
  int add_priors = 1;
  // every scale term k evaluates the same design (covariate, knots) as location term k:
  // twin[q] = scale term paired with the q-th location term (staged and dispatched once)
  int shared = 0;
  int twin[ptm::kMaxTerms] = {0};
  int num_sms = 148;
  int device = 0;
  // tuning knobs, read from the environment ONCE at create (0 / -1 = planner default)
  struct Tune {
    int tpw = 0, nu = 0, to = 0, items_per_sm = 0;
    int xtwx_tc = 3;   // PTM_XTWX_TC: 0 banded, 1 / 2 tensor-core versions, 3 = pick
    int tc_flush = 0;  // PTM_TC_FLUSH
    int main_tc = 1;   // PTM_MAIN_TC: 0 = CUDA-core sweep only, 1 = tensor-core sweep when the model fits
    int mtc_items = 0; // PTM_MTC_ITEMS: work items per CTA of the tensor-core sweep (0 = default)
    int mtc_nsw = 0;   // PTM_MTC_NSW: stager warps of k_main_tc (0 = default)
  } tune;
  // shard views share the device arrays of the handle they were made from
  const PtmProblem* base = nullptr;
  mutable int views = 0;
  // device buffers (element type by dtype)
  void* d_y = nullptr;
  void* d_X = nullptr;
  void* d_knots = nullptr;
  void* d_grid = nullptr;
  void* d_Zall = nullptr;
  void* d_Kall = nullptr;
  void* d_Zd = nullptr;
  void* d_tc = nullptr;
  void* d_tc64 = nullptr;  // TrafoConst<double> for the epilogue
  void* d_meta = nullptr;  // int [T][3]: p, koff, boff
  ptm::TrafoConst<double> tc64;  // host copy (double); cast for f32
};

namespace {

template <typename R>
R get(const void* p, long long i) { return reinterpret_cast<const R*>(p)[i]; }

// B-spline basis row at x for the transformation knots, in double.
void host_basis_row(const std::vector<double>& k, int nbases, double x, std::vector<double>& row) {
  row.assign(nbases, 0.0);
  const double inv_dk = (double)(k.size() - 1) / (k.back() - k.front());
  int j = ptm::find_interval<double>(x, k.data(), nbases, inv_dk);
  double b[4];
  ptm::bspline4<double>(x, k.data(), j, b);
  for (int m = 0; m < 4; ++m)
    if (j - 3 + m < nbases) row[j - 3 + m] = b[m];
}

template <typename R>
void fill_tc(const ptm::TrafoConst<double>& s, ptm::TrafoConst<R>& d) {
  d.KK = s.KK; d.a = (R)s.a; d.b = (R)s.b; d.inv_dk = (R)s.inv_dk; d.w = (R)s.w;
  d.inv_w = (R)s.inv_w; d.min_eps = (R)s.min_eps; d.max_eps = (R)s.max_eps;
  d.inv_h = (R)s.inv_h; d.kap_scale = (R)s.kap_scale; d.ratio = (R)s.ratio;
  // distance (in cells) from a grid point below which the exact table decides the cell: the
  // multiply-and-floor estimate carries ~1e-13 (double) / ~1.2e-4 (float: rounding of r - a and of
  // the product) of a cell, so 1e-6 / 1e-3 keep the cell index bit-exact with searchsorted while
  // a warp of 32 chains takes the table path in ~6 % of the float rows (it was 73 % at 2e-2)
  d.frac_eps = sizeof(R) == 8 ? (R)1e-6 : (R)1e-3;
  d.nparam = s.nparam; d.J = s.J; d.e_off = s.e_off; d.variant = s.variant; d.dk = (R)s.dk; d.log_dk = (R)s.log_dk; d.k1 = (R)s.k1;
  d.log_KK = (R)s.log_KK;
  for (int i = 0; i < ptm::kMaxJ; ++i) d.Bk[i] = (R)s.Bk[i];
  for (int i = 0; i < ptm::kMaxNparam; ++i) d.wts[i] = (R)s.wts[i];
}

template <typename R>
int create_impl(const PtmProblemDesc* d, PtmProblem* pr) {
  const long long n = d->n;
  const int T = d->n_terms;
  // ---- validate + offsets -------------------------------------------------------
  int zo = 0, ko = 0, bo = 2, go = 0;
  for (int t = 0; t < T; ++t) {
    const int nb = d->nbases[t], p = d->ncoef[t];
    if (nb < 4 || nb > PTM_MAX_NBASES) return fail(PTM_ERR_SHAPE, "nbases must be in [4, 32]");
    if (p < 1 || p > nb) return fail(PTM_ERR_SHAPE, "ncoef must be in [1, nbases]");
    if (d->predictor[t] != PTM_LOC && d->predictor[t] != PTM_SCALE)
      return fail(PTM_ERR_SHAPE, "predictor must be PTM_LOC or PTM_SCALE");
    if (!d->x[t] || !d->knots[t] || !d->Z[t] || !d->K[t]) return fail(PTM_ERR_NULL, "null term array");
    pr->nb[t] = nb; pr->p[t] = p; pr->pred[t] = d->predictor[t]; pr->rank[t] = d->rank[t];
    pr->nc[t] = d->noncentered ? (d->noncentered[t] != 0) : 0;
    pr->qidx[t] = d->predictor[t] == PTM_LOC ? pr->T_loc++ : pr->T_scale++;
    pr->zoff[t] = zo; zo += nb * p;
    pr->koff[t] = ko; ko += p * p;
    pr->boff[t] = bo; bo += p;
    pr->goff[t] = go; go += nb;
    pr->max_nb = std::max(pr->max_nb, nb);
    pr->max_p = std::max(pr->max_p, p);
    const R* kn = reinterpret_cast<const R*>(d->knots[t]);
    for (int i = 0; i + 1 < nb + 4; ++i)
      if (!(kn[i] < kn[i + 1])) return fail(PTM_ERR_SHAPE, "knots must be strictly increasing");
    pr->inv_dk[t] = (double)(nb + 3) / ((double)kn[nb + 3] - (double)kn[0]);
    // the reference raises when a covariate leaves [knots[3], knots[nb]]
    // (bspline/approx.py:197-206); NaN fails the comparison like jnp.min would
    const R* xv = reinterpret_cast<const R*>(d->x[t]);
    const R lo = kn[3], hi = kn[nb];
    for (long long i = 0; i < n; ++i)
      if (!(xv[i] >= lo && xv[i] <= hi)) {
        char buf[200];
        snprintf(buf, sizeof buf,
                 "Values of x are not inside the range of interior knots, [%g, %g] (term %d, obs %lld)",
                 (double)lo, (double)hi, t, i);
        return fail(PTM_ERR_RANGE, buf);
      }
  }
  pr->sum_nb = go;
  pr->T = T;
  // shared designs: every scale term has the same covariate values, knots and basis count as
  // one location term (the usual "same covariates in both predictors" model)
  pr->shared = 0;
  {
    auto env_int = [](const char* name, int dflt) {
      const char* e = getenv(name);
      return e ? atoi(e) : dflt;
    };
    pr->tune.tpw = env_int("PTM_TPW", 0);
    pr->tune.nu = env_int("PTM_NU", 0);
    pr->tune.to = env_int("PTM_TO", 0);
    pr->tune.items_per_sm = env_int("PTM_ITEMS_PER_SM", 0);
    pr->tune.xtwx_tc = env_int("PTM_XTWX_TC", 3);
    pr->tune.tc_flush = env_int("PTM_TC_FLUSH", 0);
    pr->tune.main_tc = env_int("PTM_MAIN_TC", 1);
    pr->tune.mtc_items = env_int("PTM_MTC_ITEMS", 0);
    pr->tune.mtc_nsw = env_int("PTM_MTC_NSW", 0);
  }
  if (pr->T_loc == pr->T_scale && pr->T_loc > 0 && !getenv("PTM_NO_SHARED")) {
    std::vector<int> loc, sc;
    for (int t = 0; t < T; ++t) (d->predictor[t] == PTM_LOC ? loc : sc).push_back(t);
    // one-to-one matching in any order: the first unmatched scale term with the same design
    bool ok = true;
    std::vector<char> used(sc.size(), 0);
    for (size_t k = 0; k < loc.size() && ok; ++k) {
      const int a = loc[k];
      int found = -1;
      for (size_t m = 0; m < sc.size() && found < 0; ++m) {
        const int b = sc[m];
        if (used[m] || pr->nb[a] != pr->nb[b]) continue;
        if (memcmp(d->knots[a], d->knots[b], (size_t)(pr->nb[a] + 4) * sizeof(R)) != 0) continue;
        if (d->x[a] != d->x[b] && memcmp(d->x[a], d->x[b], (size_t)n * sizeof(R)) != 0) continue;
        found = (int)m;
      }
      ok = found >= 0;
      if (ok) { used[found] = 1; pr->twin[k] = sc[found]; }
    }
    pr->shared = ok ? 1 : 0;
  }
  pr->n = n;
  pr->nparam = d->nparam;
  pr->nc_dz = d->trafo_noncentered != 0;
  pr->variant = d->trafo_variant;
  pr->J = d->nparam + (d->trafo_variant == PTM_TRAFO_ONION ? 7 : 1);  // ptm.py:55-57, onion.py:50-59
  pr->KK = pr->J - 3;
  pr->dz_off = bo;
  pr->tau_off = bo + d->nparam - 1;
  pr->taud_off = pr->tau_off + T;
  pr->P = pr->taud_off + 1;
  pr->obs_begin = 0;
  pr->obs_end = n;

  // ---- transformation constants (double, then cast) -----------------------------
  const int J = pr->J, L = J + 4;
  std::vector<double> tk(L);
  for (int i = 0; i < L; ++i) tk[i] = (double)get<R>(d->trafo_knots, i);
  for (int i = 0; i + 1 < L; ++i)
    if (!(tk[i] < tk[i + 1])) return fail(PTM_ERR_SHAPE, "trafo knots must be strictly increasing");
  ptm::TrafoConst<double>& tc = pr->tc64;
  memset(&tc, 0, sizeof tc);
  tc.nparam = d->nparam; tc.J = J; tc.KK = J - 3;
  const bool onion = d->trafo_variant == PTM_TRAFO_ONION;
  tc.variant = d->trafo_variant;
  tc.e_off = onion ? 4 : 1;
  tc.a = tk[3]; tc.b = tk[J];
  double dsum = 0;
  for (int i = 0; i + 1 < L; ++i) dsum += tk[i + 1] - tk[i];
  tc.dk = dsum / (L - 1);
  tc.log_dk = std::log(tc.dk);
  tc.inv_dk = 1.0 / tc.dk;
  tc.k1 = onion ? tk[2] : tk[3];
  // the onion form has boundary value a and slope 1 by construction, so any transition is the
  // identity off the core (onion.py:133-138); the width only has to be finite
  double eps = onion ? 0.1 : d->trafo_lambda;
  if (eps < 1e-6) return fail(PTM_ERR_SHAPE, "eps is < 1e-6; that is numerically unstable.");
  tc.w = eps * (tc.b - tc.a);
  tc.inv_w = 1.0 / tc.w;
  tc.min_eps = tc.a - tc.w;
  tc.max_eps = tc.b + tc.w;
  tc.inv_h = (double)(ptm::kGridN - 1) / (tc.b - tc.a);
  tc.kap_scale = (double)ptm::kGridN / (double)(ptm::kGridN - 1);
  // Gaussian model: the sweep evaluates the PTM spline at delta = 0, whose samples are the identity;
  // with the exact lerp weight (instead of the reference's range/1000 step, approx.py:375-380) the
  // lerp of those samples is the identity too, h' = 1 (dist.py:846-850: value, zeros)
  if (d->trafo_variant == PTM_TRAFO_IDENTITY) tc.kap_scale = 1.0;
  tc.ratio = (double)tc.KK / (double)(ptm::kGridN - 1);
  tc.log_KK = std::log((double)(J - 3));
  if (onion) {  // onion.py:74-108: sum of the free increments = nparam * dk, no weights, theta_0 = knots[2]
    tc.log_KK = std::log((double)d->nparam);
    for (int j = 0; j < d->nparam; ++j) tc.wts[j] = 1.0;
  } else {
    std::vector<double> row;
    host_basis_row(tk, J, tk[3], row);
    double run = 0;
    for (int j = J - 1; j >= 0; --j) { run += row[j]; tc.Bk[j] = run; }
    for (int j = 0; j < d->nparam; ++j) tc.wts[j] = 1.0;
    tc.wts[0] = 1.0 / 6.0; tc.wts[1] = 1.0 / 6.0 + 2.0 / 3.0;
    tc.wts[d->nparam - 1] = 1.0 / 6.0; tc.wts[d->nparam - 2] = 2.0 / 3.0 + 1.0 / 6.0;
  }

  // ---- device copies ---------------------------------------------------------------
  PTM_CUDA(cudaGetDevice(&pr->device));
  PTM_CUDA(cudaDeviceGetAttribute(&pr->num_sms, cudaDevAttrMultiProcessorCount, pr->device));
  const size_t es = sizeof(R);
  PTM_CUDA(cudaMalloc(&pr->d_y, std::max<size_t>(n, 1) * es));
  PTM_CUDA(cudaMemcpy(pr->d_y, d->y, n * es, cudaMemcpyHostToDevice));
  PTM_CUDA(cudaMalloc(&pr->d_X, std::max<size_t>((size_t)T * n, 1) * es));
  for (int t = 0; t < T; ++t)
    PTM_CUDA(cudaMemcpy((char*)pr->d_X + (size_t)t * n * es, d->x[t], n * es, cudaMemcpyHostToDevice));
  {
    std::vector<R> kn((size_t)std::max(T, 1) * ptm::kKnotStride, R(0));
    for (int t = 0; t < T; ++t)
      for (int i = 0; i < pr->nb[t] + 4; ++i) kn[(size_t)t * ptm::kKnotStride + i] = get<R>(d->knots[t], i);
    PTM_CUDA(cudaMalloc(&pr->d_knots, kn.size() * es));
    PTM_CUDA(cudaMemcpy(pr->d_knots, kn.data(), kn.size() * es, cudaMemcpyHostToDevice));
    std::vector<R> zall(std::max(zo, 1)), kall(std::max(ko, 1));
    for (int t = 0; t < T; ++t) {
      for (int i = 0; i < pr->nb[t] * pr->p[t]; ++i) zall[pr->zoff[t] + i] = get<R>(d->Z[t], i);
      for (int i = 0; i < pr->p[t] * pr->p[t]; ++i) kall[pr->koff[t] + i] = get<R>(d->K[t], i);
    }
    PTM_CUDA(cudaMalloc(&pr->d_Zall, zall.size() * es));
    PTM_CUDA(cudaMemcpy(pr->d_Zall, zall.data(), zall.size() * es, cudaMemcpyHostToDevice));
    PTM_CUDA(cudaMalloc(&pr->d_Kall, kall.size() * es));
    PTM_CUDA(cudaMemcpy(pr->d_Kall, kall.data(), kall.size() * es, cudaMemcpyHostToDevice));
    const size_t nzd = (size_t)d->nparam * (d->nparam - 1);
    PTM_CUDA(cudaMalloc(&pr->d_Zd, nzd * es));
    PTM_CUDA(cudaMemcpy(pr->d_Zd, d->trafo_Z, nzd * es, cudaMemcpyHostToDevice));
    // grid: as the reference built it (data), or the restated jnp.linspace formula
    std::vector<R> grid(ptm::kGridN + 2);
    if (d->trafo_grid) {
      for (int i = 0; i < ptm::kGridN + 2; ++i) grid[i] = get<R>(d->trafo_grid, i);
    } else {
      const R a = (R)tk[3], b = (R)tk[J];
      const R step = (b - a) / R(ptm::kGridN);
      grid[0] = a - step;
      for (int i = 0; i < ptm::kGridN - 1; ++i) {
        const R s = R(i) / R(ptm::kGridN - 1);
        grid[1 + i] = a * (R(1) - s) + b * s;
      }
      grid[ptm::kGridN] = b;
      grid[ptm::kGridN + 1] = b + step;
    }
    PTM_CUDA(cudaMalloc(&pr->d_grid, grid.size() * es));
    PTM_CUDA(cudaMemcpy(pr->d_grid, grid.data(), grid.size() * es, cudaMemcpyHostToDevice));
    ptm::TrafoConst<R> tcr;
    memset(&tcr, 0, sizeof tcr);
    fill_tc<R>(tc, tcr);
    PTM_CUDA(cudaMalloc(&pr->d_tc, sizeof tcr));
    PTM_CUDA(cudaMemcpy(pr->d_tc, &tcr, sizeof tcr, cudaMemcpyHostToDevice));
    std::vector<int> meta(std::max(T, 1) * 3, 0);
    for (int t = 0; t < T; ++t) { meta[t * 3] = pr->p[t]; meta[t * 3 + 1] = pr->koff[t]; meta[t * 3 + 2] = pr->boff[t]; }
    PTM_CUDA(cudaMalloc(&pr->d_meta, meta.size() * sizeof(int)));
    PTM_CUDA(cudaMemcpy(pr->d_meta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice));
    ptm::TrafoConst<double> tcd;
    memset(&tcd, 0, sizeof tcd);
    fill_tc<double>(tc, tcd);
    PTM_CUDA(cudaMalloc(&pr->d_tc64, sizeof tcd));
    PTM_CUDA(cudaMemcpy(pr->d_tc64, &tcd, sizeof tcd, cudaMemcpyHostToDevice));
  }
  return PTM_OK;
}

template <typename R>
Plan make_plan(const PtmProblem* pr, long long C) {
  Plan pl;
  const int T = pr->T;
  pl.CG = (int)((C + 31) / 32);
  const int NGV = pr->KK + 4;
  pl.NSLOT = ptm::kRegSlots + NGV + pr->sum_nb;
  const int NCC = pr->KK * 5 + 6;
  // eval warps: fill the CTA up to 20 warps (96 registers per thread; registers are
  // granted per 4 warps), measured best at c3 (T = 10 -> NU = 10); beyond 12 terms up to
  // 28 warps (72 registers; T = 20 -> NU = 8 measured 5 % faster than 4)
  // beyond 12 terms a term warp owns two terms (two accumulator sets in registers): T = 20
  // -> 10 term warps + 6 eval warps = 512 threads at 128 registers, instead of 20 + 8 warps
  // at 72 (where the eval warps spill)
  pl.TPW = T > 12 ? 2 : 1;  // (shared designs keep one term per warp: see k_main)
  if (pr->tune.tpw) pl.TPW = pr->tune.tpw == 2 ? 2 : 1;  // tuning knob (PTM_TPW at create)
  pl.NTW = (T + pl.TPW - 1) / pl.TPW;
  if (pl.TPW == 1)
    pl.NU = T <= 12 ? std::min(12, std::max(8, 20 - T)) : std::max(4, std::min(8, 28 - T));
  else  // 16 warps (128 registers per thread) while that leaves at least 4 eval warps
    pl.NU = pl.NTW <= 12 ? std::max(4, 16 - pl.NTW) : 4;
  if (pr->tune.nu) pl.NU = std::max(1, std::min(32 - pl.NTW, pr->tune.nu));  // tuning knob (PTM_NU)
  auto smem_for = [&](int To) {
    size_t b = 0;
    b += (size_t)pr->sum_nb * 32 * sizeof(R);                 // gamma [coef][lane]
    b += (size_t)(NCC + 4) * 32 * sizeof(R);                  // cubic pieces of h + 4 derived boundary rows
    b += (size_t)pl.NU * NGV * 32 * sizeof(R);                // private grad tables
    b += std::max((size_t)2 * To * 2 * 32 * sizeof(R),        // dl/deta hand-off, reused for the
                  (size_t)pl.NU * ptm::kRegSlots * 32 * sizeof(double));  // scalar partials at item end
    b += (size_t)T * 3 * To * 4 * sizeof(R);                  // staged records: 3 basis values + row offset
    b += (size_t)T * 4 * ptm::kKnotStride * sizeof(R);        // knots + reciprocals
    b += 64;                                                  // tail pad
    return b;
  };
  pl.To = 4 * pl.NU;  // four observations per eval warp and tile
  if (pr->tune.to) pl.To = std::max(8, pr->tune.to);  // tuning knob (PTM_TO)
  while (pl.To > 8 && smem_for(pl.To) > 227 * 1024) pl.To -= 4;
  // large coefficient tables (many terms x many bases): give up eval warps (their private
  // gradient tables) before giving up
  while (pl.NU > 4 && smem_for(pl.To) > 227 * 1024) {
    --pl.NU;
    pl.To = std::max(8, std::min(pl.To, 4 * pl.NU));
  }
  pl.smem_main = smem_for(pl.To);
  const long long len = std::max<long long>(pr->obs_end - pr->obs_begin, 0);
  long long per_sm = 8;  // work items per CTA: amortises the per-item table loads, keeps the tail short
  if (pr->tune.items_per_sm) per_sm = std::max(1, pr->tune.items_per_sm);  // tuning knob (PTM_ITEMS_PER_SM)
  const long long target_items = per_sm * pr->num_sms;
  long long M = std::max<long long>(1, target_items / pl.CG);
  const long long max_m = std::max<long long>(1, len / ((long long)pl.To * 16));  // >= 16 tiles per item
  M = std::min(M, max_m);
  {
    // whole waves: make CG * M a multiple of the number of CTAs when the cap allows it
    long long g = pl.CG, h = pr->num_sms;
    while (h) { const long long t = g % h; g = h; h = t; }
    const long long unit = pr->num_sms / g;  // smallest M with CG * M % num_sms == 0
    if (M >= unit) M = M / unit * unit;
  }
  long long chunk = (len + M - 1) / M;
  chunk = std::max<long long>(pl.To, (chunk + pl.To - 1) / pl.To * pl.To);
  M = std::max<long long>(1, (len + chunk - 1) / chunk);
  pl.M = (int)M;
  pl.chunk_len = chunk;
  pl.n_items = pl.CG * pl.M;
  pl.grid = std::min(pl.n_items, pr->num_sms);
  // float problems whose coefficient count fits tensor memory: eta and the coefficient
  // gradients on tcgen05 (ptm_main_tc.cuh)
  size_t g_bytes = 0, e_bytes = 0;
  long long n_part_items = pl.n_items;
  if (sizeof(R) == 4 && pr->tune.main_tc && T >= 1 && T <= 24) {
    int K8[2] = {0, 0}, K16[2] = {0, 0};
    for (int t = 0; t < T; ++t) {
      const int q = pr->pred[t] == PTM_LOC ? 0 : 1;
      K8[q] += (pr->nb[t] + 3) & ~3;
      K16[q] += (pr->nb[t] + 3) & ~3;
    }
    for (int q = 0; q < 2; ++q) { K8[q] = (K8[q] + 7) & ~7; K16[q] = (K16[q] + 15) & ~15; }
    // Groups of terms per launch.  Everything in one launch when the coefficients fit tensor
    // memory (2 K columns for Gamma hi / lo); otherwise eta-only launches peel off groups of
    // <= 224 coefficients of one predictor (largest predictor first) until the rest fits.
    unsigned groups[8] = {0};
    int ngroups = 0, ext_mask = 0, Kt = 0, K16t = 0;
    bool groups_ok = true;
    {
      int nb4[ptm::kMaxTerms];
      bool left[ptm::kMaxTerms];
      int rest[2] = {0, 0};
      for (int t = 0; t < T; ++t) {
        nb4[t] = (pr->nb[t] + 3) & ~3;
        left[t] = true;
        rest[pr->pred[t] == PTM_LOC ? 0 : 1] += nb4[t];
      }
      auto pad8 = [](int v) { return (v + 7) & ~7; };
      while (pad8(rest[0]) + pad8(rest[1]) > ptm::kMtcMaxK) {
        if (ngroups >= 7) { groups_ok = false; break; }
        const int q = rest[0] >= rest[1] ? 0 : 1;
        unsigned m = 0;
        int k = 0;
        for (int t = 0; t < T; ++t)
          if (left[t] && (pr->pred[t] == PTM_LOC ? 0 : 1) == q && k + nb4[t] <= ptm::kMtcMaxK) {
            m |= 1u << t; k += nb4[t]; left[t] = false;
          }
        if (!m) { groups_ok = false; break; }
        rest[q] -= k;
        groups[ngroups++] = m;
        ext_mask |= 1 << q;
        Kt = std::max(Kt, pad8(k));
        K16t = std::max(K16t, (k + 15) & ~15);
      }
      unsigned m = 0;
      for (int t = 0; t < T; ++t) if (left[t]) m |= 1u << t;
      groups[ngroups++] = m;  // the main launch (may be empty of terms: every term went external)
      Kt = std::max(Kt, pad8(rest[0]) + pad8(rest[1]));
      K16t = std::max(K16t, ((rest[0] + 15) & ~15) + ((rest[1] + 15) & ~15));
    }
    const bool two_pass = ngroups > 1;
    const int avail = (512 - 2 * Kt) / 2;
    int NT = avail >= 64 ? 64 : (avail >= 48 ? 48 : (avail >= 32 ? 32 : 0));
    // stager warps: the terms of a predictor block of a launch are dealt round-robin (term index
    // modulo NSW) and the two blocks of a tile are filled one after the other, so the count that
    // gives every warp the same number of terms per block is the one to take (c3: 5 + 5 terms -> 5,
    // 14.2 -> 14.05 ms; c4: 7 location terms in the eta-only launch -> 7; tools/mtc_c5_nsw.py)
    int nsw_auto = std::max(1, (T + 7) / 8);
    for (int gi = 0; gi < ngroups; ++gi)
      for (int q = 0; q < 2; ++q) {
        int tq = 0;
        for (int t = 0; t < T; ++t)
          if ((groups[gi] >> t & 1u) && (pr->pred[t] == PTM_LOC ? 0 : 1) == q) ++tq;
        if (tq > 0) {
          const int rounds = (tq + 6) / 7;
          nsw_auto = std::max(nsw_auto, (tq + rounds - 1) / rounds);
        }
      }
    const int NSW = std::min(T, pr->tune.mtc_nsw > 0 ? std::min(pr->tune.mtc_nsw, 7) : nsw_auto);
    auto fwd_smem = [&](int nt) {
      return (size_t)2 * nt * Kt * 4 + (size_t)(NCC + 4) * 128 * 4 + (size_t)ptm::kMtcEvalWarps * NGV * 32 * 4 +
             (size_t)T * 4 * ptm::kKnotStride * 4 +
             std::max((size_t)ptm::kMtcEvalWarps * ptm::kRegSlots * 32 * 8,            // reduction scratch, aliased
                      (size_t)ptm::kMtcEvalWarps * (2 * (nt / 4) * 32) * 4) + 128;  // onto the eta buffers
    };
    while (NT > 32 && fwd_smem(NT) > 227 * 1024 - 1280) NT -= 16;
    if (groups_ok && Kt <= ptm::kMtcMaxK && NT > 0 && K16t <= 256 && len > 0) {
      pl.tc_smem_fwd = fwd_smem(NT);
      pl.tc_smem_bwd = (size_t)2 * 2 * K16t * ptm::kGtcKT * 4 + (size_t)T * 4 * ptm::kKnotStride * 4 +
                       (size_t)ptm::kGtcStages * ptm::kGtcTileBytes + 256;
      if (pl.tc_smem_fwd <= 227 * 1024 - 1280 && pl.tc_smem_bwd <= 227 * 1024 - 1280) {
        pl.tc = true;
        pl.tc_npass = ngroups;
        for (int i = 0; i < ngroups; ++i) pl.tc_terms[i] = groups[i];
        pl.tc_ext_mask = ext_mask;
        pl.tc_NT = NT; pl.tc_NSW = NSW;
        pl.tc_CB = (int)((C + 127) / 128);
        // work items per CTA: every item re-loads Gamma / the piece table and drains the pipeline, so
        // one wave when there are few chain blocks, two when eight blocks share the SMs (measured:
        // tools/mtc_chains.py)
        const long long items_per_sm = pr->tune.mtc_items > 0 ? pr->tune.mtc_items : (pl.tc_CB >= 8 ? 2 : 1);
        long long Mt = std::max<long long>(1, items_per_sm * pr->num_sms / pl.tc_CB);
        Mt = std::min(Mt, std::max<long long>(1, len / (8LL * 192)));
        long long ch = (len + Mt - 1) / Mt;
        ch = std::max<long long>(192, (ch + 191) / 192 * 192);  // whole forward (32/48/64) and backward (32) tiles
        Mt = std::max<long long>(1, (len + ch - 1) / ch);
        pl.tc_M = (int)Mt; pl.tc_chunk = ch;
        n_part_items = std::max<long long>(n_part_items, (long long)pl.CG * Mt);
        g_bytes = (size_t)(len + ptm::kGtcKT) * 2 * (size_t)pl.tc_CB * 128 * 4;  // one backward tile of padding per block
        if (two_pass) e_bytes = 2 * (size_t)(len + ptm::kGtcKT) * (size_t)pl.tc_CB * 128 * 4;  // E[0], E[1]
      }
    }
  }
  size_t off = 0;
  pl.off_gamma = off; off = align_up(off + (size_t)std::max(T, 1) * ptm::kMaxNbases * C * sizeof(R), 256);
  pl.off_cc = off;    off = align_up(off + (size_t)NCC * C * sizeof(R), 256);
  pl.off_partial = off; off = align_up(off + (size_t)n_part_items * pl.NSLOT * 32 * sizeof(double), 256);
  pl.off_xt = off;    off = align_up(off + ptm::xtwx_workspace_bytes<R>(std::max(pr->sum_nb, 1), C), 256);
  pl.off_scratch = off; off = align_up(off + (size_t)C * sizeof(R), 256);
  pl.off_G = off;     off = align_up(off + g_bytes, 256);
  pl.off_red = off;   off = align_up(off + (size_t)pl.CG * pl.NSLOT * 32 * sizeof(double), 256);
  pl.off_E = off;     off = align_up(off + e_bytes, 256);
  pl.total = off;
  return pl;
}

template <typename R, int NB, int MODE, int MAXT, int TPW, bool SHARED = false>
int launch_main_t(const ptm::MainArgs<R>& a, const Plan& pl, int T, cudaStream_t st) {
  auto kern = ptm::k_main<R, NB, MODE, MAXT, TPW, SHARED>;
  PTM_CUDA(ptm::ensure_smem(kern, pl.smem_main));
  kern<<<pl.grid, 32 * (pl.NTW + pl.NU), pl.smem_main, st>>>(a);
  PTM_CUDA(cudaGetLastError());
  return PTM_OK;
}

// CTA = term warps + NU eval warps (c3: 10 + 10 warps, 96 registers per thread); with two
// terms per term warp (T > 12, or shared designs) 512 threads at 128 registers cover up to
// 24 terms.
template <typename R, int NB, int MODE>
int launch_main(const ptm::MainArgs<R>& a, const Plan& pl, int T, cudaStream_t st) {
  // registers are granted per 4 warps: <= 16 warps -> 128/thread, 18 -> 112, 20 -> 96, 24 -> 80, 32 -> 64
  const int threads = 32 * (pl.NTW + pl.NU);
  if (pl.TPW == 2) {
    if (a.shared) return launch_main_t<R, NB, MODE, 512, 2, true>(a, pl, T, st);
    if (threads <= 512) return launch_main_t<R, NB, MODE, 512, 2>(a, pl, T, st);
    if (threads <= 640) return launch_main_t<R, NB, MODE, 640, 2>(a, pl, T, st);
    return launch_main_t<R, NB, MODE, 768, 2>(a, pl, T, st);
  }
  if (a.shared && threads <= 640) return launch_main_t<R, NB, MODE, 640, 1, true>(a, pl, T, st);
  if (threads <= 512) return launch_main_t<R, NB, MODE, 512, 1>(a, pl, T, st);
  if (threads <= 640) return launch_main_t<R, NB, MODE, 640, 1>(a, pl, T, st);
  if (threads <= 768) return launch_main_t<R, NB, MODE, 768, 1>(a, pl, T, st);
  return launch_main_t<R, NB, MODE, 1024, 1>(a, pl, T, st);
}

template <typename R>
void fill_main_args(const PtmProblem* pr, const Plan& pl, long long C, R* Gamma, R* cc, double* partial,
                    ptm::MainArgs<R>& ma) {
  memset(&ma, 0, sizeof ma);
  ma.y = reinterpret_cast<const R*>(pr->d_y);
  ma.X = reinterpret_cast<const R*>(pr->d_X);
  ma.knots = reinterpret_cast<const R*>(pr->d_knots);
  ma.grid = reinterpret_cast<const R*>(pr->d_grid);
  ma.Gamma = Gamma; ma.cc = cc; ma.partial = partial;
  ma.n = pr->n; ma.obs_begin = pr->obs_begin; ma.obs_end = pr->obs_end;
  ma.chunk_len = pl.chunk_len;
  ma.C = (int)C; ma.CG = pl.CG; ma.M = pl.M; ma.n_items = pl.n_items;
  ma.NSLOT = pl.NSLOT;
  ma.KK = pr->KK;
  ma.T = pr->T; ma.T_loc = pr->T_loc; ma.T_scale = pr->T_scale;
  ma.To = pl.To; ma.NU = pl.NU; ma.sum_nb = pr->sum_nb;
  {
    int q = 0;
    for (int t = 0; t < pr->T; ++t) if (pr->pred[t] == PTM_LOC) ma.order[q++] = t;
    for (int t = 0; t < pr->T; ++t) if (pr->pred[t] == PTM_SCALE) ma.order[q++] = t;
  }
  // the shared-design kernels exist for the CTA sizes the planner picks (<= 640 threads with
  // one term per warp, <= 512 with two); anything else runs the general kernel
  const int cta_threads = 32 * (pl.NTW + pl.NU);
  ma.shared = pr->shared && cta_threads <= (pl.TPW == 2 ? 512 : 640);
  if (ma.shared) {
    // slots 2k / 2k+1 = location term k and its twin: one term warp, one staged design
    int k = 0;
    for (int t = 0; t < pr->T; ++t)
      if (pr->pred[t] == PTM_LOC) {
        ma.order[2 * k] = t;
        ma.order[2 * k + 1] = pr->twin[k];
        ma.dgo[k] = (pr->goff[pr->twin[k]] - pr->goff[t]) * 32;
        ++k;
      }
  }
  for (int t = 0; t < pr->T; ++t) {
    ma.nb[t] = pr->nb[t]; ma.pred[t] = pr->pred[t];
    ma.goff[t] = pr->goff[t]; ma.inv_dk[t] = (R)pr->inv_dk[t];
  }
  ptm::TrafoConst<R> tcr;
  fill_tc<R>(pr->tc64, tcr);
  ma.ts = tcr;
}

template <typename R>
void fill_pre_args(const PtmProblem* pr, const R* params, long long C, R* Gamma, R* cc, ptm::PreArgs<R>& pa) {
  memset(&pa, 0, sizeof pa);
  pa.params = params; pa.Gamma = Gamma; pa.cc = cc;
  pa.Zall = reinterpret_cast<const R*>(pr->d_Zall);
  pa.Zd = reinterpret_cast<const R*>(pr->d_Zd);
  pa.tc = reinterpret_cast<const ptm::TrafoConst<double>*>(pr->d_tc64);
  pa.C = (int)C; pa.P = pr->P; pa.T = pr->T; pa.dz_off = pr->dz_off;
  pa.tau_off = pr->tau_off; pa.taud_off = pr->taud_off; pa.nc_dz = pr->nc_dz;
  for (int t = 0; t < pr->T; ++t) {
    pa.nb[t] = pr->nb[t]; pa.p[t] = pr->p[t]; pa.zoff[t] = pr->zoff[t]; pa.boff[t] = pr->boff[t];
    pa.nc[t] = pr->nc[t];
  }
}

// The handle's device must be the calling thread's current device: the kernels launch on the
// current device with pointers that live on pr->device.
int check_device(const PtmProblem* pr) {
  int dev = -1;
  PTM_CUDA(cudaGetDevice(&dev));
  if (dev != pr->device) {
    char buf[160];
    snprintf(buf, sizeof buf, "problem lives on device %d but the current device is %d (cudaSetDevice first)",
             pr->device, dev);
    return fail(PTM_ERR_CUDA, buf);
  }
  return PTM_OK;
}
#define PTM_CHECK_DEVICE(pr)                                \
  do {                                                      \
    const int rc_dev_ = check_device(pr);                   \
    if (rc_dev_ != PTM_OK) return rc_dev_;                  \
  } while (0)

bool is_sharded(const PtmProblem* pr) { return pr->obs_begin != 0 || pr->obs_end != pr->n || !pr->add_priors; }

// K-column / D-column tables of the terms in `term_mask`: forward K columns padded to 8 per
// predictor block, backward D columns to 16.
void fill_tc_tables(const PtmProblem* pr, unsigned term_mask, ptm::MainTcArgs& g) {
  const int NGV = pr->KK + 4;
  memset(g.kterm, 255, sizeof g.kterm);
  for (size_t i = 0; i < sizeof g.dslot / sizeof g.dslot[0]; ++i) g.dslot[i] = -1;
  for (int t = 0; t < ptm::kMaxTerms; ++t) g.kcol[t] = g.ncol[t] = -1;  // -1: the term is not in this launch
  int k = 0, n = 0;
  for (int q = 0; q < 2; ++q) {
    g.kbase[q] = k; g.nbase[q] = n;
    {
      for (int t = 0; t < pr->T; ++t) {
        if ((pr->pred[t] == PTM_LOC ? 0 : 1) != q || !(term_mask & (1u << t))) continue;
        g.kcol[t] = k; g.ncol[t] = n;
        for (int j = 0; j < pr->nb[t]; ++j) {
          g.kterm[k + j] = (unsigned char)t; g.kj[k + j] = (unsigned char)j;
          g.dslot[n + j] = (short)(ptm::kRegSlots + NGV + pr->goff[t] + j);
        }
        k += (pr->nb[t] + 3) & ~3; n += (pr->nb[t] + 3) & ~3;
      }
    }
    k = (k + 7) & ~7; n = (n + 15) & ~15;
    g.Kq[q] = k - g.kbase[q]; g.Nq[q] = n - g.nbase[q];
  }
  g.Kt = k;
}

// Tensor-core route of the float sweep: k_main_tc (+ k_grad_tc when the gradient is wanted).
// Returns PTM_OK or an error; the caller has checked pl.tc.
int launch_main_tc(const PtmProblem* pr, const Plan& pl, long long C, float* Gamma, float* cc, double* partial,
                   float* G, float* E, bool want_grad, cudaStream_t st) {
  static thread_local ptm::MainTcArgs g;  // 3 KB of launch arguments: keep it off the stack hot path
  memset(&g, 0, sizeof g);
  Plan plc = pl;
  plc.M = pl.tc_M; plc.chunk_len = pl.tc_chunk; plc.n_items = pl.tc_CB * pl.tc_M;
  fill_main_args<float>(pr, plc, C, Gamma, cc, partial, g.a);
  g.a.shared = 0;
  if (!want_grad) g.a.NSLOT = ptm::kLogpSlots;
  g.Gld = std::max<long long>(pr->obs_end - pr->obs_begin, 0) + ptm::kGtcKT;
  g.G = G; g.E[0] = g.E[1] = E; g.CB = pl.tc_CB; g.Cpad = pl.tc_CB * 128; g.NT = pl.tc_NT; g.NSW = pl.tc_NSW;
  g.mode = want_grad ? ptm::kModeGrad : ptm::kModeLogp;
  g.flush = pr->tune.tc_flush > 0 ? pr->tune.tc_flush : 8;
  g.NSWb = std::min(pr->T, 10);  // 16 loader + 10 stager + MMA + copy warps = 896 threads
  const int threads = 32 * ((ptm::kMtcEvalWarps + pl.tc_NSW + 1 + 3) / 4 * 4);
  const int grid = std::min(plc.n_items, pr->num_sms);
#define PTM_MTC_LAUNCH2(NTV, GR, MT, EO)                                               \
  do {                                                                                 \
    auto kern = ptm::k_main_tc<NTV, GR, MT, EO>;                                       \
    PTM_CUDA(ptm::ensure_smem(kern, pl.tc_smem_fwd));                                  \
    kern<<<grid, threads, pl.tc_smem_fwd, st>>>(g);                                    \
  } while (0)
#define PTM_MTC_LAUNCH(NTV, EO)                                                        \
  do {                                                                                 \
    if constexpr (EO) {  /* the eta-only launch has no gradient form */                \
      if (threads <= 640) PTM_MTC_LAUNCH2(NTV, false, 640, true); else PTM_MTC_LAUNCH2(NTV, false, 768, true); \
    } else if (threads <= 640) {                                                       \
      if (want_grad) PTM_MTC_LAUNCH2(NTV, true, 640, false); else PTM_MTC_LAUNCH2(NTV, false, 640, false); \
    } else {                                                                           \
      if (want_grad) PTM_MTC_LAUNCH2(NTV, true, 768, false); else PTM_MTC_LAUNCH2(NTV, false, 768, false); \
    }                                                                                  \
  } while (0)
#define PTM_MTC_LAUNCH_NT(EO)                                                          \
  do {                                                                                 \
    if (pl.tc_NT == 64) PTM_MTC_LAUNCH(64, EO);                                        \
    else if (pl.tc_NT == 48) PTM_MTC_LAUNCH(48, EO);                                   \
    else PTM_MTC_LAUNCH(32, EO);                                                       \
    PTM_CUDA(cudaGetLastError());                                                      \
    ++g_launches;                                                                      \
  } while (0)
  size_t e_stride = 0;
  {
    const long long len = std::max<long long>(pr->obs_end - pr->obs_begin, 0);
    e_stride = (size_t)(len + ptm::kGtcKT) * (size_t)pl.tc_CB * 128;
  }
  g.E[0] = E; g.E[1] = E + e_stride;
  bool seen[2] = {false, false};
  for (int i = 0; i + 1 < pl.tc_npass; ++i) {   // eta-only launches: one predictor's group each
    fill_tc_tables(pr, pl.tc_terms[i], g);
    const int q = g.Kq[0] > 0 ? 0 : 1;
    g.e_acc = seen[q] ? 1 : 0;
    g.ext_mask = 0;
    seen[q] = true;
    PTM_MTC_LAUNCH_NT(true);
  }
  fill_tc_tables(pr, pl.tc_terms[pl.tc_npass - 1], g);
  g.e_acc = 0;
  g.ext_mask = pl.tc_ext_mask;
  PTM_MTC_LAUNCH_NT(false);
#undef PTM_MTC_LAUNCH_NT
#undef PTM_MTC_LAUNCH
#undef PTM_MTC_LAUNCH2
  if (want_grad) {
    auto kern = ptm::k_grad_tc;
    PTM_CUDA(ptm::ensure_smem(kern, pl.tc_smem_bwd));
    const int threads_b = 32 * ((ptm::kMtcEvalWarps + g.NSWb + 2 + 3) / 4 * 4);  // loaders, stagers, MMA, copy
    for (int i = 0; i < pl.tc_npass; ++i) {   // one launch per group: its blocks of dGamma
      if (pl.tc_npass > 1) {
        if (!pl.tc_terms[i]) continue;          // (a main launch without terms)
        fill_tc_tables(pr, pl.tc_terms[i], g);
      }
      kern<<<grid, threads_b, pl.tc_smem_bwd, st>>>(g);
      PTM_CUDA(cudaGetLastError());
      ++g_launches;
    }
  }
  return PTM_OK;
}

static thread_local int g_last_main_route = 0;  // 0 = CUDA-core sweep (k_main), 1 = tcgen05 sweep

// ld_logp / ld_grad: element strides of the outputs (1 / P for separate arrays, 1 + P for the
// packed [C, 1 + P] block of ptm_logp_grad_block_*).
template <typename R>
int eval_impl(const PtmProblem* pr, const R* params, long long C, R* logp, R* grad, void* ws,
              size_t ws_bytes, void* stream, R* stats = nullptr, long long ld_logp = 1, long long ld_grad = 0) {
  if (!pr || !params || !logp || !ws) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  if (C < 1) return fail(PTM_ERR_SHAPE, "n_chains must be >= 1");
  PTM_CHECK_DEVICE(pr);
  const Plan pl = make_plan<R>(pr, C);
  if (ws_bytes < pl.total) return fail(PTM_ERR_WORKSPACE, "workspace too small");
  if (pl.smem_main > 227 * 1024) return fail(PTM_ERR_SHAPE, "shared-memory budget exceeded");
  if (pl.NTW + pl.NU > 32) return fail(PTM_ERR_SHAPE, "too many smooth terms for one CTA");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* wsb = reinterpret_cast<char*>(ws);
  R* Gamma = reinterpret_cast<R*>(wsb + pl.off_gamma);
  R* cc = reinterpret_cast<R*>(wsb + pl.off_cc);
  double* partial = reinterpret_cast<double*>(wsb + pl.off_partial);
  const bool want_grad = grad != nullptr;
  g_launches = 0;

  ptm::PreArgs<R> pa;
  fill_pre_args<R>(pr, params, C, Gamma, cc, pa);
  if (g_profile) PTM_CUDA(cudaEventRecord(g_ev[0], st));
  ptm::k_pre<R><<<(unsigned)C, 128, pr->P * sizeof(R), st>>>(pa);
  PTM_CUDA(cudaGetLastError());
  ++g_launches;
  if (g_profile) PTM_CUDA(cudaEventRecord(g_ev[1], st));

  ptm::MainArgs<R> ma;
  fill_main_args<R>(pr, pl, C, Gamma, cc, partial, ma);
  if (!want_grad) ma.NSLOT = ptm::kLogpSlots;
  int rc;
  int post_M = pl.M;
  bool took_tc = false;
  if constexpr (sizeof(R) == 4) {
    if (pl.tc) {
      rc = launch_main_tc(pr, pl, C, reinterpret_cast<float*>(Gamma), reinterpret_cast<float*>(cc), partial,
                          reinterpret_cast<float*>(wsb + pl.off_G), reinterpret_cast<float*>(wsb + pl.off_E),
                          want_grad, st);
      if (rc != PTM_OK) return rc;
      post_M = pl.tc_M;
      took_tc = true;
    }
  }
  g_last_main_route = took_tc ? 1 : 0;
  if (!took_tc) {
    if (pr->max_nb <= 20)
      rc = want_grad ? launch_main<R, 20, ptm::kModeGrad>(ma, pl, pr->T, st)
                     : launch_main<R, 20, ptm::kModeLogp>(ma, pl, pr->T, st);
    else
      rc = want_grad ? launch_main<R, 32, ptm::kModeGrad>(ma, pl, pr->T, st)
                     : launch_main<R, 32, ptm::kModeLogp>(ma, pl, pr->T, st);
    if (rc != PTM_OK) return rc;
    ++g_launches;
  }
  if (g_profile) PTM_CUDA(cudaEventRecord(g_ev[2], st));

  // chunk partials -> one row per chain group (many CTAs), then the per-chain epilogue
  double* red = reinterpret_cast<double*>(wsb + pl.off_red);
  {
    const int per = ma.NSLOT * 32;
    const dim3 rgrid((unsigned)pl.CG, (unsigned)std::max(1, std::min(32, (2 * pr->num_sms + pl.CG - 1) / pl.CG)));
    ptm::k_reduce_partials<<<rgrid, 256, 0, st>>>(partial, post_M, per, red);
    PTM_CUDA(cudaGetLastError());
    ++g_launches;
  }
  ptm::PostArgs<R> po;
  memset(&po, 0, sizeof po);
  po.params = params; po.partial = red; po.logp = logp; po.grad = grad; po.stats = stats;
  po.ld_logp = ld_logp; po.ld_grad = ld_grad ? ld_grad : pr->P;
  po.Zall = reinterpret_cast<const R*>(pr->d_Zall);
  po.Kall = reinterpret_cast<const R*>(pr->d_Kall);
  po.Zd = reinterpret_cast<const R*>(pr->d_Zd);
  po.tc = reinterpret_cast<const ptm::TrafoConst<double>*>(pr->d_tc64);
  po.C = (int)C; po.P = pr->P; po.T = pr->T; po.M = 1; po.NSLOT = ma.NSLOT; po.KK = pr->KK;
  for (int t = 0; t < pr->T; ++t) {
    po.nb[t] = pr->nb[t]; po.p[t] = pr->p[t]; po.zoff[t] = pr->zoff[t]; po.koff[t] = pr->koff[t];
    po.boff[t] = pr->boff[t]; po.goff[t] = pr->goff[t]; po.rank[t] = pr->rank[t];
  }
  po.dz_off = pr->dz_off; po.tau_off = pr->tau_off; po.taud_off = pr->taud_off;
  po.n_obs = std::max<long long>(pr->obs_end - pr->obs_begin, 0);
  po.add_priors = pr->add_priors;
  po.nc_dz = pr->nc_dz;
  for (int t = 0; t < pr->T; ++t) po.nc[t] = pr->nc[t];
  size_t smem_post = ((size_t)ma.NSLOT * 32 + (size_t)std::max(pr->T, 1) * 32) * sizeof(double);
  // the 32 parameter rows of a chain group next to the reduced partials when they fit
  po.stage_par = smem_post + (size_t)32 * pr->P * sizeof(R) <= 200 * 1024 ? 1 : 0;
  if (po.stage_par) smem_post += (size_t)32 * pr->P * sizeof(R);
  if (want_grad) {
    auto kp = ptm::k_post<R, true>;
    PTM_CUDA(ptm::ensure_smem(kp, smem_post));
    kp<<<pl.CG, 256, smem_post, st>>>(po);
  } else {
    auto kp = ptm::k_post<R, false>;
    PTM_CUDA(ptm::ensure_smem(kp, smem_post));
    kp<<<pl.CG, 256, smem_post, st>>>(po);
  }
  PTM_CUDA(cudaGetLastError());
  ++g_launches;
  if (g_profile) PTM_CUDA(cudaEventRecord(g_ev[3], st));
  return PTM_OK;
}

// Pointwise predictive records over S posterior samples (evaluate.py:120-133, 177-236):
// lppd_i and p_waic_i for every observation of the shard.  The sweep is the logp-only
// kernel in its pointwise mode with one CTA per observation chunk visiting all sample
// groups in turn; the records [n][4] live behind the ordinary workspace.
template <typename R>
int pointwise_impl(const PtmProblem* pr, const R* params, long long S, R* lppd, R* pwaic, void* ws,
                   size_t ws_bytes, void* stream) {
  if (!pr || !params || !lppd || !pwaic || !ws) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (S < 1) return fail(PTM_ERR_SHAPE, "n_samples must be >= 1");
  Plan pl = make_plan<R>(pr, S);
  const size_t need = pl.total + (size_t)pr->n * 4 * sizeof(double);
  if (ws_bytes < need) return fail(PTM_ERR_WORKSPACE, "workspace too small (see ptm_pointwise_workspace_bytes)");
  if (pl.smem_main > 227 * 1024) return fail(PTM_ERR_SHAPE, "shared-memory budget exceeded");
  if (pl.NTW + pl.NU > 32) return fail(PTM_ERR_SHAPE, "too many smooth terms for one CTA");
  const long long len = std::max<long long>(pr->obs_end - pr->obs_begin, 0);
  if (len == 0) return PTM_OK;
  // one CTA per observation chunk, all sample groups of a chunk on the same CTA:
  // item = cg * M + mchunk with M == gridDim.x
  long long M = std::min<long long>(pr->num_sms, std::max<long long>(1, len / pl.To));
  long long chunk = (len + M - 1) / M;
  chunk = std::max<long long>(pl.To, (chunk + pl.To - 1) / pl.To * pl.To);
  M = std::max<long long>(1, (len + chunk - 1) / chunk);
  pl.M = (int)M; pl.chunk_len = chunk; pl.n_items = pl.CG * pl.M; pl.grid = pl.M;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* wsb = reinterpret_cast<char*>(ws);
  R* Gamma = reinterpret_cast<R*>(wsb + pl.off_gamma);
  R* cc = reinterpret_cast<R*>(wsb + pl.off_cc);
  double* point = reinterpret_cast<double*>(wsb + pl.total);
  g_launches = 0;
  ptm::PreArgs<R> pa;
  fill_pre_args<R>(pr, params, S, Gamma, cc, pa);
  ptm::k_pre<R><<<(unsigned)S, 128, pr->P * sizeof(R), st>>>(pa);
  PTM_CUDA(cudaGetLastError());
  ++g_launches;
  ptm::MainArgs<R> ma;
  fill_main_args<R>(pr, pl, S, Gamma, cc, nullptr, ma);
  ma.point = point;
  int rc = launch_main<R, 20, ptm::kModePoint>(ma, pl, pr->T, st);
  if (rc != PTM_OK) return rc;
  ++g_launches;
  const int blocks = (int)std::min<long long>((len + 255) / 256, 148LL * 8);
  ptm::k_point_final<R><<<blocks, 256, 0, st>>>(point, pr->obs_begin, pr->obs_end, S, lppd, pwaic);
  PTM_CUDA(cudaGetLastError());
  ++g_launches;
  return PTM_OK;
}

template <typename R>
int basis_impl(const PtmProblem* pr, int term, int32_t* interval, R* values, void* stream) {
  if (!pr || !interval || !values) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (term < 0 || term >= pr->T) return fail(PTM_ERR_SHAPE, "term index out of range");
  if (pr->n == 0) return PTM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = (int)std::min<long long>((pr->n + 255) / 256, 148LL * 16);
  ptm::k_basis<R><<<blocks, 256, 0, st>>>(
      reinterpret_cast<const R*>(pr->d_X) + (size_t)term * pr->n, pr->n,
      reinterpret_cast<const R*>(pr->d_knots) + (size_t)term * ptm::kKnotStride, pr->nb[term],
      (R)pr->inv_dk[term], interval, values);
  PTM_CUDA(cudaGetLastError());
  return PTM_OK;
}

template <typename R>
int transform_impl(const PtmProblem* pr, const R* delta, long long C, const R* r, long long m, R* z,
                   R* dz, int32_t* cell, void* stream) {
  if (!pr || !delta || !r || !z || !dz) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (C < 1 || m < 0) return fail(PTM_ERR_SHAPE, "bad shape");
  if (m == 0) return PTM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ptm::k_transform<R><<<(unsigned)C, 256, 0, st>>>(
      delta, (int)C, r, m, reinterpret_cast<const ptm::TrafoConst<R>*>(pr->d_tc),
      reinterpret_cast<const R*>(pr->d_grid), z, dz, cell);
  PTM_CUDA(cudaGetLastError());
  return PTM_OK;
}

// ---- callers of the path on new data: predictive z / log-density / CDF, quantiles, inverse --
size_t predict_ws_bytes(const PtmProblem* pr, long long S) {
  const size_t w = pr->dtype == PTM_F64 ? 8 : 4;
  const size_t gam = ((size_t)pr->T * ptm::kMaxNbases * S * w + 255) / 256 * 256;
  const size_t cc = ((size_t)(pr->KK * 5 + 6) * S * w + 255) / 256 * 256;
  const size_t H = (size_t)S * ptm::kGridN * w;  // spline samples of every sample (quantile mode)
  return gam + cc + H;
}

template <typename R>
int predict_impl(const PtmProblem* pr, const R* params, long long S, const R* y_or_u, const R* x_new,
                 long long m, R* z, R* logpdf, R* cdf, R* yq, void* ws, size_t ws_bytes, void* stream,
                 bool u_per_sample = false) {
  if (!pr || !params || !y_or_u || !ws) return fail(PTM_ERR_NULL, "null argument");
  if (pr->T > 0 && !x_new) return fail(PTM_ERR_NULL, "x_new is null");
  if (!z && !logpdf && !cdf && !yq) return fail(PTM_ERR_NULL, "no output requested");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (S < 1 || m < 0) return fail(PTM_ERR_SHAPE, "bad shape");
  if (S > 65535 && pr->T > 12) return fail(PTM_ERR_SHAPE, "at most 65535 samples per call beyond 12 terms");
  if (ws_bytes < predict_ws_bytes(pr, S)) return fail(PTM_ERR_WORKSPACE, "workspace too small (see ptm_predict_workspace_bytes)");
  if (m == 0) return PTM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* wsb = reinterpret_cast<char*>(ws);
  R* Gamma = reinterpret_cast<R*>(wsb);
  R* cc = reinterpret_cast<R*>(wsb + ((size_t)pr->T * ptm::kMaxNbases * S * sizeof(R) + 255) / 256 * 256);
  ptm::PreArgs<R> pa;
  fill_pre_args<R>(pr, params, S, Gamma, cc, pa);
  ptm::k_pre<R><<<(unsigned)S, 128, pr->P * sizeof(R), st>>>(pa);
  PTM_CUDA(cudaGetLastError());
  ptm::PredictArgs<R> a;
  memset(&a, 0, sizeof a);
  a.y = y_or_u; a.X = x_new; a.m = m;
  a.u_ld = u_per_sample ? m : 0;
  a.knots = reinterpret_cast<const R*>(pr->d_knots);
  a.Gamma = Gamma; a.cc = cc; a.grid = reinterpret_cast<const R*>(pr->d_grid);
  a.S = (int)S; a.T = pr->T; a.KK = pr->KK; a.sum_nb = pr->sum_nb;
  for (int t = 0; t < pr->T; ++t) {
    a.nb[t] = pr->nb[t]; a.pred[t] = pr->pred[t] == PTM_LOC ? 0 : 1;
    a.goff[t] = pr->goff[t]; a.inv_dk[t] = (R)pr->inv_dk[t];
  }
  ptm::TrafoConst<R> tcr;
  fill_tc<R>(pr->tc64, tcr);
  a.ts = tcr;
  a.out_z = z; a.out_logpdf = logpdf; a.out_cdf = cdf; a.out_q = yq;
  if (pr->T <= 12) {
    // observation-major sweep: basis windows once per observation, samples staged in groups
    const long long bx = (m + ptm::kPredThreads - 1) / ptm::kPredThreads;
    const long long groups = (S + ptm::kPredSG - 1) / ptm::kPredSG;
    long long by = std::max<long long>(1, std::min<long long>(groups, (2LL * pr->num_sms + bx - 1) / bx));
    by = std::min<long long>(by, 65535);
    const int s_per_block = (int)((groups + by - 1) / by) * ptm::kPredSG;
    by = (S + s_per_block - 1) / s_per_block;
    R* H = nullptr;
    if (yq) {
      H = reinterpret_cast<R*>(wsb + predict_ws_bytes(pr, S) - (size_t)S * ptm::kGridN * sizeof(R));
      ptm::k_grid_samples<R><<<(unsigned)S, 256, 0, st>>>(cc, (int)S, pr->KK, a.ts, H);
      PTM_CUDA(cudaGetLastError());
    }
    const dim3 grid((unsigned)bx, (unsigned)by);
    const size_t smem = (size_t)ptm::kPredSG * (pr->sum_nb + pr->KK * 5 + 6) * sizeof(R);
#define PTM_PREDICT_LAUNCH(TT)                                                                         \
    do {                                                                                               \
      if (yq) ptm::k_predict_obs<R, true, TT><<<grid, ptm::kPredThreads, smem, st>>>(a, H, s_per_block); \
      else ptm::k_predict_obs<R, false, TT><<<grid, ptm::kPredThreads, smem, st>>>(a, H, s_per_block);   \
    } while (0)
    if (pr->T <= 4) PTM_PREDICT_LAUNCH(4);
    else if (pr->T <= 8) PTM_PREDICT_LAUNCH(8);
    else PTM_PREDICT_LAUNCH(12);
#undef PTM_PREDICT_LAUNCH
    PTM_CUDA(cudaGetLastError());
    return PTM_OK;
  }
  // more than 12 terms: sample-major sweep (basis windows recomputed per sample)
  const long long per_sample = std::max<long long>(1, ((long long)pr->num_sms * 8 + S - 1) / S);
  const long long gx = std::max<long long>(1, std::min<long long>((m + 255) / 256, per_sample));
  const dim3 grid((unsigned)gx, (unsigned)S);
  const size_t smem = (size_t)(pr->sum_nb + pr->KK * 5 + (yq ? ptm::kGridN : 0)) * sizeof(R);
  if (yq) ptm::k_predict<R, true><<<grid, 256, smem, st>>>(a);
  else ptm::k_predict<R, false><<<grid, 256, smem, st>>>(a);
  PTM_CUDA(cudaGetLastError());
  return PTM_OK;
}

template <typename R>
int inverse_impl(const PtmProblem* pr, const R* delta, long long C, const R* z, long long m, R* r,
                 void* stream) {
  if (!pr || !delta || !z || !r) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (C < 1 || m < 0) return fail(PTM_ERR_SHAPE, "bad shape");
  if (m == 0) return PTM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  ptm::k_inverse<R><<<(unsigned)C, 256, 0, st>>>(
      delta, (int)C, z, m, reinterpret_cast<const ptm::TrafoConst<R>*>(pr->d_tc),
      reinterpret_cast<const R*>(pr->d_grid), r);
  PTM_CUDA(cudaGetLastError());
  return PTM_OK;
}

static thread_local int g_last_xtwx_route = 0;  // 0 = banded CUDA-core sweep, 1 = tcgen05 sweep
// PtmProblem::tune.xtwx_tc (PTM_XTWX_TC at create): 0 banded sweep, 1 tensor-core contraction
// with the weights on CUDA cores, 2 scale predictor on the tensor cores too; unset (3) = pick 2
// when its Q tiles can be double-buffered, else 1.

// Information sweep for the terms in `terms` (all LOC, or one SCALE term): shared
// prologue (gamma of the scale terms), one banded sweep, one epilogue per term.
// Outputs are per term: out_x / out_p / out_l[i] may be null.
template <typename R>
int xtwx_terms(const PtmProblem* pr, const int* terms, int nterms, const R* params, long long C,
               R* const* out_x, R* const* out_p, R* const* out_l, void* ws, size_t ws_bytes, void* stream) {
  if (!pr || !params || !ws) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (C < 1) return fail(PTM_ERR_SHAPE, "n_chains must be >= 1");
  if (nterms < 1) return PTM_OK;
  for (int i = 0; i < nterms; ++i)
    if (terms[i] < 0 || terms[i] >= pr->T) return fail(PTM_ERR_SHAPE, "term index out of range");
  const bool loc = pr->pred[terms[0]] == PTM_LOC;
  for (int i = 1; i < nterms; ++i)
    if ((pr->pred[terms[i]] == PTM_LOC) != loc) return fail(PTM_ERR_SHAPE, "terms must share a predictor");
  if (!loc && nterms != 1) return fail(PTM_ERR_SHAPE, "one scale term per call");
  // a shard view holds part of the observations: X'WX is additive over shards, the precision
  // (penalty, 1e-6 mean(diag) ridge) and its Cholesky factor are not -- all-reduce ptm_xtwx and
  // finish on the full sum (parallel.finish_precision) instead of getting shard-local factors
  bool wants_factor = false;
  for (int i = 0; i < nterms; ++i) wants_factor |= (out_p && out_p[i]) || (out_l && out_l[i]);
  if (is_sharded(pr) && wants_factor)
    return fail(PTM_ERR_SHAPE,
                "precision / Cholesky of an observation-shard view are not additive over shards: "
                "use ptm_xtwx_* per shard, all-reduce, then add the penalty and factorise");
  const Plan pl = make_plan<R>(pr, C);
  if (ws_bytes < pl.total) return fail(PTM_ERR_WORKSPACE, "workspace too small");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* wsb = reinterpret_cast<char*>(ws);
  R* Gamma = reinterpret_cast<R*>(wsb + pl.off_gamma);
  R* cc = reinterpret_cast<R*>(wsb + pl.off_cc);
  double* xt = reinterpret_cast<double*>(wsb + pl.off_xt);
  ptm::PreArgs<R> pa;
  fill_pre_args<R>(pr, params, C, Gamma, cc, pa);
  g_launches = 0;
  ptm::k_pre<R><<<(unsigned)C, 128, pr->P * sizeof(R), st>>>(pa);
  PTM_CUDA(cudaGetLastError());
  ++g_launches;
  ptm::XtwxArgs<R> xa;
  memset(&xa, 0, sizeof xa);
  xa.X = reinterpret_cast<const R*>(pr->d_X);
  xa.knots = reinterpret_cast<const R*>(pr->d_knots);
  xa.Gamma = Gamma; xa.cc = cc; xa.params = params;
  xa.Zall = reinterpret_cast<const R*>(pr->d_Zall);
  xa.Kall = reinterpret_cast<const R*>(pr->d_Kall);
  xa.n = pr->n; xa.obs_begin = pr->obs_begin; xa.obs_end = pr->obs_end;
  xa.C = (int)C; xa.P = pr->P; xa.T = pr->T; xa.KK = pr->KK;
  xa.tau_off = pr->tau_off;
  xa.add_penalty = pr->add_priors;
  for (int t = 0; t < pr->T; ++t) {
    xa.nb[t] = pr->nb[t]; xa.p[t] = pr->p[t];
    xa.zoff[t] = pr->zoff[t]; xa.koff[t] = pr->koff[t]; xa.inv_dk[t] = (R)pr->inv_dk[t];
    xa.nc[t] = pr->nc[t];
  }
  xa.const_w = loc ? 0 : 1;  // scale terms: W == 2 (iwls_proposals.py:118-119)
  if (loc)
    for (int t = 0; t < pr->T; ++t)
      if (pr->pred[t] == PTM_SCALE) {
        xa.sc_term[xa.nsc] = t; xa.sc_goff[xa.nsc] = xa.sum_sc_nb; xa.sum_sc_nb += pr->nb[t]; ++xa.nsc;
      }
  // as many requested terms per sweep as the 20-warp / shared-memory budget allows
  int done = 0;
  while (done < nterms) {
    int rc = -1, take = 0;
    auto select = [&](int cnt) {
      xa.nbt = cnt; xa.band_slots = 0;
      for (int i = 0; i < cnt; ++i) {
        xa.bt_term[i] = terms[done + i]; xa.bt_boff[i] = xa.band_slots; xa.band_slots += pr->nb[terms[done + i]] * 4;
      }
    };
    if constexpr (sizeof(R) == 4) {
      // float problems, location terms: tensor-core route (ptm_xtwx_tc.cuh) with as many
      // terms per sweep as TMEM / shared memory take
      if (loc && pr->tune.xtwx_tc != 0)
        for (take = std::min(nterms - done, 8); take >= 1 && rc == -1; --take) {
          select(take);
          rc = pr->tune.xtwx_tc >= 2 ? ptm::xtwx_tc2_sweep(xa, xt, pr->num_sms, pr->tune.tc_flush,
                                                           pr->tune.xtwx_tc == 2, st, &g_launches)
                                     : -1;
          if (rc != -1) { g_last_xtwx_route = 2; break; }
          rc = ptm::xtwx_tc_sweep(xa, xt, pr->num_sms, pr->tune.tc_flush, st, &g_launches);
          if (rc != -1) { g_last_xtwx_route = 1; break; }
        }
    }
    if (rc == -1)
      for (take = std::min(nterms - done, std::max(1, 16 - xa.nsc)); take >= 1; --take) {
        select(take);
        rc = ptm::xtwx_sweep<R>(xa, xt, pr->num_sms, st, &g_launches);
        if (rc != -1) { g_last_xtwx_route = 0; break; }
      }
    if (rc == -1) return fail(PTM_ERR_SHAPE, "information sweep does not fit in shared memory");
    if (rc != 0) return fail(PTM_ERR_CUDA, std::string("xtwx sweep: ") + cudaGetErrorString((cudaError_t)rc));
    for (int i = 0; i < take; ++i) {
      const int k = done + i;
      rc = ptm::xtwx_post<R>(xa, terms[k], xa.bt_boff[i], out_x ? out_x[k] : nullptr, out_p ? out_p[k] : nullptr,
                             out_l ? out_l[k] : nullptr, st, &g_launches);
      if (rc != 0) return fail(PTM_ERR_CUDA, std::string("xtwx post: ") + cudaGetErrorString((cudaError_t)rc));
    }
    done += take;
  }
  return PTM_OK;
}

template <typename R>
int xtwx_impl(const PtmProblem* pr, int term, const R* params, long long C, R* xtwx, R* precision,
              R* chol, void* ws, size_t ws_bytes, void* stream) {
  R* ox[1] = {xtwx}; R* op[1] = {precision}; R* ol[1] = {chol};
  return xtwx_terms<R>(pr, &term, 1, params, C, ox, op, ol, ws, ws_bytes, stream);
}

// Precision + Cholesky of ALL location terms from one sweep (they share the weights).
template <typename R>
int precision_all_impl(const PtmProblem* pr, const R* params, long long C, R* precision, R* chol, void* ws,
                       size_t ws_bytes, void* stream) {
  if (!pr || !precision) return fail(PTM_ERR_NULL, "null argument");
  int terms[PTM_MAX_TERMS]; R* op[PTM_MAX_TERMS]; R* ol[PTM_MAX_TERMS];
  int nt = 0; size_t off = 0;
  for (int t = 0; t < pr->T; ++t)
    if (pr->pred[t] == PTM_LOC) {
      terms[nt] = t; op[nt] = precision + off; ol[nt] = chol ? chol + off : nullptr;
      off += (size_t)C * pr->p[t] * pr->p[t]; ++nt;
    }
  return xtwx_terms<R>(pr, terms, nt, params, C, nullptr, op, ol, ws, ws_bytes, stream);
}

template <typename R>
int quadform_impl(const PtmProblem* pr, const R* params, long long C, R* quad, void* stream) {
  if (!pr || !params || !quad) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (C < 1) return fail(PTM_ERR_SHAPE, "n_chains must be >= 1");
  if (pr->T == 0) return PTM_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long total = C * pr->T;
  ptm::k_quadform<R><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
      params, (int)C, pr->P, pr->T, reinterpret_cast<const R*>(pr->d_Kall),
      reinterpret_cast<const int*>(pr->d_meta), quad);
  PTM_CUDA(cudaGetLastError());
  return PTM_OK;
}

// ---- observed information of one coefficient block (ptm_hessian.cuh) -------------------------
struct HessPlan {
  int kind, d, NACC, NW, CG, M, n_items, grid;
  long long chunk_len;
  size_t smem, off_gamma, off_cc, off_partial, total;
};

template <typename R>
int make_hess_plan(const PtmProblem* pr, int block, long long C, HessPlan& hp) {
  if (block != PTM_BLOCK_DELTA_Z && (block < 0 || block >= pr->T)) return fail(PTM_ERR_SHAPE, "block: term index or PTM_BLOCK_DELTA_Z");
  const bool dz = block == PTM_BLOCK_DELTA_Z;
  if (dz && pr->variant == PTM_TRAFO_IDENTITY)
    return fail(PTM_ERR_SHAPE, "the Gaussian model (PTM_TRAFO_IDENTITY) has no shape-parameter block");
  hp.kind = dz ? ptm::kHessDeltaZ : ptm::kHessBeta;
  hp.d = dz ? pr->nparam - 1 : pr->p[block];
  hp.NACC = ptm::hess_nacc(hp.kind, dz ? 0 : pr->nb[block], pr->KK);
  const int NCC = pr->KK * 5 + 6, T = pr->T;
  auto smem_for = [&](int NW) {
    size_t b = (size_t)NW * hp.NACC * 32 * sizeof(double);
    b += (size_t)pr->sum_nb * 32 * sizeof(R) + (size_t)(NCC + 4) * 32 * sizeof(R);
    b += (size_t)T * 4 * ptm::kKnotStride * sizeof(R);
    b += (size_t)NW * T * 32 * 4 * sizeof(R) + (size_t)NW * 32 * sizeof(R) + 64;
    return b;
  };
  hp.NW = 8;
  while (hp.NW > 1 && smem_for(hp.NW) > 227 * 1024) --hp.NW;
  hp.smem = smem_for(hp.NW);
  if (hp.smem > 227 * 1024) return fail(PTM_ERR_SHAPE, "observed-information sweep does not fit in shared memory");
  hp.CG = (int)((C + 31) / 32);
  const long long len = std::max<long long>(pr->obs_end - pr->obs_begin, 0);
  long long M = std::max<long long>(1, (4LL * pr->num_sms + hp.CG - 1) / hp.CG);
  M = std::min(M, std::max<long long>(1, len / (32LL * hp.NW * 4)));
  long long chunk = std::max<long long>(32, (len + M - 1) / M);
  chunk = (chunk + 31) / 32 * 32;
  M = std::max<long long>(1, (len + chunk - 1) / chunk);
  hp.M = (int)M; hp.chunk_len = chunk; hp.n_items = hp.CG * hp.M;
  hp.grid = std::min(hp.n_items, pr->num_sms);
  size_t off = 0;
  hp.off_gamma = off; off = align_up(off + (size_t)std::max(T, 1) * ptm::kMaxNbases * C * sizeof(R), 256);
  hp.off_cc = off;    off = align_up(off + (size_t)NCC * C * sizeof(R), 256);
  hp.off_partial = off; off = align_up(off + (size_t)hp.n_items * hp.NACC * 32 * sizeof(double), 256);
  hp.total = off;
  return PTM_OK;
}

template <typename R>
int hessian_impl(const PtmProblem* pr, int block, int mode, const R* params, long long C, R* info, R* chol,
                 void* ws, size_t ws_bytes, void* stream) {
  if (!pr || !params || !info || !ws) return fail(PTM_ERR_NULL, "null argument");
  if (pr->dtype != (sizeof(R) == 8 ? PTM_F64 : PTM_F32)) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  PTM_CHECK_DEVICE(pr);
  if (C < 1) return fail(PTM_ERR_SHAPE, "n_chains must be >= 1");
  if (mode < 0 || mode > 4) return fail(PTM_ERR_SHAPE, "mode must be 0 .. 4");
  // like the precision: the eigen-shift and the factor are not additive over observation shards
  if (is_sharded(pr) && (mode != 0 || chol))
    return fail(PTM_ERR_SHAPE, "observed information of a shard view: mode 0 without Cholesky only (additive), finish on the sum");
  HessPlan hp;
  int rc = make_hess_plan<R>(pr, block, C, hp);
  if (rc != PTM_OK) return rc;
  if (ws_bytes < hp.total) return fail(PTM_ERR_WORKSPACE, "workspace too small (see ptm_hessian_workspace_bytes)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  char* wsb = reinterpret_cast<char*>(ws);
  R* Gamma = reinterpret_cast<R*>(wsb + hp.off_gamma);
  R* cc = reinterpret_cast<R*>(wsb + hp.off_cc);
  double* partial = reinterpret_cast<double*>(wsb + hp.off_partial);
  g_launches = 0;
  ptm::PreArgs<R> pa;
  fill_pre_args<R>(pr, params, C, Gamma, cc, pa);
  ptm::k_pre<R><<<(unsigned)C, 128, pr->P * sizeof(R), st>>>(pa);
  PTM_CUDA(cudaGetLastError());
  ++g_launches;
  ptm::HessArgs<R> ha;
  memset(&ha, 0, sizeof ha);
  ha.y = reinterpret_cast<const R*>(pr->d_y);
  ha.X = reinterpret_cast<const R*>(pr->d_X);
  ha.knots = reinterpret_cast<const R*>(pr->d_knots);
  ha.grid = reinterpret_cast<const R*>(pr->d_grid);
  ha.Gamma = Gamma; ha.cc = cc; ha.partial = partial;
  ha.n = pr->n; ha.obs_begin = pr->obs_begin; ha.obs_end = pr->obs_end; ha.chunk_len = hp.chunk_len;
  ha.C = (int)C; ha.CG = hp.CG; ha.M = hp.M; ha.n_items = hp.n_items; ha.KK = pr->KK; ha.T = pr->T;
  ha.sum_nb = pr->sum_nb; ha.NW = hp.NW; ha.NACC = hp.NACC; ha.kind = hp.kind;
  ha.bterm = hp.kind == ptm::kHessBeta ? block : 0;
  for (int t = 0; t < pr->T; ++t) {
    ha.nb[t] = pr->nb[t]; ha.pred[t] = pr->pred[t] == PTM_LOC ? 0 : 1;
    ha.goff[t] = pr->goff[t]; ha.inv_dk[t] = (R)pr->inv_dk[t];
  }
  ptm::TrafoConst<R> tcr;
  fill_tc<R>(pr->tc64, tcr);
  ha.ts = tcr;
  {
    auto kern = ptm::k_hess<R>;
    PTM_CUDA(ptm::ensure_smem(kern, hp.smem));
    kern<<<hp.grid, 32 * hp.NW, hp.smem, st>>>(ha);
    PTM_CUDA(cudaGetLastError());
    ++g_launches;
  }
  ptm::HessPostArgs<R> po;
  memset(&po, 0, sizeof po);
  po.params = params; po.partial = partial;
  po.Zall = reinterpret_cast<const R*>(pr->d_Zall);
  po.Kall = reinterpret_cast<const R*>(pr->d_Kall);
  po.Zd = reinterpret_cast<const R*>(pr->d_Zd);
  po.tc = reinterpret_cast<const ptm::TrafoConst<double>*>(pr->d_tc64);
  po.hess = info; po.chol = chol;
  po.C = (int)C; po.P = pr->P; po.M = hp.M; po.NACC = hp.NACC; po.KK = pr->KK; po.mode = mode;
  po.add_priors = pr->add_priors;
  po.tau_off = pr->tau_off; po.taud_off = pr->taud_off; po.dz_off = pr->dz_off;
  po.nc = hp.kind == ptm::kHessBeta ? pr->nc[block] : pr->nc_dz;
  if (hp.kind == ptm::kHessBeta) {
    po.term = block; po.nbt = pr->nb[block]; po.p = pr->p[block]; po.zoff = pr->zoff[block]; po.koff = pr->koff[block];
    const size_t sm = ((size_t)po.nbt * 4 + (size_t)po.nbt * po.p + 2 * (size_t)po.p * po.p) * sizeof(double);
    auto kp = ptm::k_hess_post_beta<R>;
    PTM_CUDA(ptm::ensure_smem(kp, sm));
    kp<<<(unsigned)C, 128, sm, st>>>(po);
  } else {
    const int J = pr->J, np_ = pr->nparam;
    const size_t sm = ((size_t)hp.NACC + (size_t)J * J + 2 * (size_t)np_ * np_ + (size_t)(J + 1) + J + 3 * (size_t)np_ +
                       4 * (size_t)(J + 1) + 8) * sizeof(double);
    auto kp = ptm::k_hess_post_dz<R>;
    PTM_CUDA(ptm::ensure_smem(kp, sm));
    kp<<<(unsigned)C, 64, sm, st>>>(po);
  }
  PTM_CUDA(cudaGetLastError());
  ++g_launches;
  return PTM_OK;
}

void free_all(PtmProblem* p) {
  if (p->base) return;  // a shard view borrows the arrays of its base handle
  cudaFree(p->d_y); cudaFree(p->d_X); cudaFree(p->d_knots); cudaFree(p->d_grid);
  cudaFree(p->d_Zall); cudaFree(p->d_Kall); cudaFree(p->d_Zd); cudaFree(p->d_tc); cudaFree(p->d_tc64); cudaFree(p->d_meta);
}

}  // namespace

extern "C" {

int ptm_problem_create(const PtmProblemDesc* d, PtmProblem** out) {
  if (!d || !out) return fail(PTM_ERR_NULL, "null argument");
  *out = nullptr;
  if (d->dtype != PTM_F32 && d->dtype != PTM_F64) return fail(PTM_ERR_DTYPE, "dtype must be PTM_F32 or PTM_F64");
  if (d->n < 0) return fail(PTM_ERR_SHAPE, "n must be >= 0");
  if (d->n_terms < 0 || d->n_terms > PTM_MAX_TERMS) return fail(PTM_ERR_SHAPE, "n_terms out of range");
  if (d->nparam < 4 || d->nparam > PTM_MAX_NPARAM) return fail(PTM_ERR_SHAPE, "nparam must be in [4, 40]");
  if (d->trafo_variant < PTM_TRAFO_PTM || d->trafo_variant > PTM_TRAFO_IDENTITY)
    return fail(PTM_ERR_SHAPE, "trafo_variant must be PTM_TRAFO_PTM, PTM_TRAFO_ONION or PTM_TRAFO_IDENTITY");
  if (d->trafo_variant == PTM_TRAFO_ONION && d->nparam + 7 > PTM_MAX_NPARAM + 1)
    return fail(PTM_ERR_SHAPE, "onion form: nparam must be <= 34 (nparam + 7 bases)");
  if (!d->y && d->n > 0) return fail(PTM_ERR_NULL, "y is null");
  if (!d->trafo_knots || !d->trafo_Z) return fail(PTM_ERR_NULL, "trafo arrays are null");
  if (d->n_terms > 0 && (!d->x || !d->nbases || !d->ncoef || !d->predictor || !d->knots || !d->Z || !d->K || !d->rank))
    return fail(PTM_ERR_NULL, "term arrays are null");
  PtmProblem* pr = new (std::nothrow) PtmProblem();
  if (!pr) return fail(PTM_ERR_CUDA, "out of host memory");
  pr->dtype = d->dtype;
  const int rc = d->dtype == PTM_F64 ? create_impl<double>(d, pr) : create_impl<float>(d, pr);
  if (rc != PTM_OK) {
    free_all(pr);
    delete pr;
    return rc;
  }
  *out = pr;
  return PTM_OK;
}

int ptm_problem_destroy(PtmProblem* p) {
  if (!p) return PTM_OK;
  free_all(p);
  delete p;
  return PTM_OK;
}

int64_t ptm_num_params(const PtmProblem* p) { return p ? p->P : -1; }

int ptm_param_layout(const PtmProblem* p, int64_t* beta0_off, int64_t* gamma0_off, int64_t* beta_off,
                     int64_t* delta_z_off, int64_t* tau_off, int64_t* tau_delta_off) {
  if (!p) return fail(PTM_ERR_NULL, "null problem");
  if (beta0_off) *beta0_off = 0;
  if (gamma0_off) *gamma0_off = 1;
  if (beta_off) for (int t = 0; t < p->T; ++t) beta_off[t] = p->boff[t];
  if (delta_z_off) *delta_z_off = p->dz_off;
  if (tau_off) for (int t = 0; t < p->T; ++t) tau_off[t] = p->tau_off + t;
  if (tau_delta_off) *tau_delta_off = p->taud_off;
  return PTM_OK;
}

size_t ptm_workspace_bytes(const PtmProblem* p, int64_t C) {
  if (!p || C < 1) return 0;
  return p->dtype == PTM_F64 ? make_plan<double>(p, C).total : make_plan<float>(p, C).total;
}

int ptm_problem_shard_view(const PtmProblem* p, int64_t b, int64_t e, int32_t add_priors, PtmProblem** out) {
  if (!p || !out) return fail(PTM_ERR_NULL, "null argument");
  *out = nullptr;
  if (b < 0 || e > p->n || b > e) return fail(PTM_ERR_SHAPE, "shard must satisfy 0 <= begin <= end <= n");
  PtmProblem* v = new (std::nothrow) PtmProblem(*p);
  if (!v) return fail(PTM_ERR_CUDA, "out of host memory");
  v->base = p->base ? p->base : p;
  v->views = 0;
  v->obs_begin = b;
  v->obs_end = e;
  v->add_priors = add_priors ? 1 : 0;
  *out = v;
  return PTM_OK;
}
int ptm_problem_shard(const PtmProblem* p, int64_t* obs_begin, int64_t* obs_end, int32_t* add_priors) {
  if (!p) return fail(PTM_ERR_NULL, "null problem");
  if (obs_begin) *obs_begin = p->obs_begin;
  if (obs_end) *obs_end = p->obs_end;
  if (add_priors) *add_priors = p->add_priors;
  return PTM_OK;
}

int ptm_logp_f64(const PtmProblem* p, const double* params, int64_t C, double* logp, void* ws, size_t wsb, void* st) {
  return eval_impl<double>(p, params, C, logp, nullptr, ws, wsb, st);
}
int ptm_logp_f32(const PtmProblem* p, const float* params, int64_t C, float* logp, void* ws, size_t wsb, void* st) {
  return eval_impl<float>(p, params, C, logp, nullptr, ws, wsb, st);
}
int ptm_logp_grad_f64(const PtmProblem* p, const double* params, int64_t C, double* logp, double* grad, void* ws,
                      size_t wsb, void* st) {
  if (!grad) return fail(PTM_ERR_NULL, "grad is null");
  return eval_impl<double>(p, params, C, logp, grad, ws, wsb, st);
}
int ptm_logp_grad_f32(const PtmProblem* p, const float* params, int64_t C, float* logp, float* grad, void* ws,
                      size_t wsb, void* st) {
  if (!grad) return fail(PTM_ERR_NULL, "grad is null");
  return eval_impl<float>(p, params, C, logp, grad, ws, wsb, st);
}

int ptm_logp_grad_block_f64(const PtmProblem* p, const double* params, int64_t C, double* block, void* ws,
                            size_t wsb, void* st) {
  if (!block || !p) return fail(PTM_ERR_NULL, "null argument");
  return eval_impl<double>(p, params, C, block, block + 1, ws, wsb, st, nullptr, p->P + 1, p->P + 1);
}
int ptm_logp_grad_block_f32(const PtmProblem* p, const float* params, int64_t C, float* block, void* ws,
                            size_t wsb, void* st) {
  if (!block || !p) return fail(PTM_ERR_NULL, "null argument");
  return eval_impl<float>(p, params, C, block, block + 1, ws, wsb, st, nullptr, p->P + 1, p->P + 1);
}

int ptm_basis_f64(const PtmProblem* p, int32_t term, int32_t* interval, double* values, void* st) {
  return basis_impl<double>(p, term, interval, values, st);
}
int ptm_basis_f32(const PtmProblem* p, int32_t term, int32_t* interval, float* values, void* st) {
  return basis_impl<float>(p, term, interval, values, st);
}

int ptm_transform_f64(const PtmProblem* p, const double* delta, int64_t C, const double* r, int64_t m, double* z,
                      double* dz, int32_t* cell, void* st) {
  return transform_impl<double>(p, delta, C, r, m, z, dz, cell, st);
}
int ptm_transform_f32(const PtmProblem* p, const float* delta, int64_t C, const float* r, int64_t m, float* z,
                      float* dz, int32_t* cell, void* st) {
  return transform_impl<float>(p, delta, C, r, m, z, dz, cell, st);
}

int ptm_xtwx_f64(const PtmProblem* p, int32_t term, const double* params, int64_t C, double* xtwx, void* ws,
                 size_t wsb, void* st) {
  if (!xtwx) return fail(PTM_ERR_NULL, "xtwx is null");
  return xtwx_impl<double>(p, term, params, C, xtwx, nullptr, nullptr, ws, wsb, st);
}
int ptm_xtwx_f32(const PtmProblem* p, int32_t term, const float* params, int64_t C, float* xtwx, void* ws,
                 size_t wsb, void* st) {
  if (!xtwx) return fail(PTM_ERR_NULL, "xtwx is null");
  return xtwx_impl<float>(p, term, params, C, xtwx, nullptr, nullptr, ws, wsb, st);
}
int ptm_loc_precision_f64(const PtmProblem* p, int32_t term, const double* params, int64_t C, double* precision,
                          double* chol, void* ws, size_t wsb, void* st) {
  if (!precision) return fail(PTM_ERR_NULL, "precision is null");
  return xtwx_impl<double>(p, term, params, C, nullptr, precision, chol, ws, wsb, st);
}
int ptm_loc_precision_f32(const PtmProblem* p, int32_t term, const float* params, int64_t C, float* precision,
                          float* chol, void* ws, size_t wsb, void* st) {
  if (!precision) return fail(PTM_ERR_NULL, "precision is null");
  return xtwx_impl<float>(p, term, params, C, nullptr, precision, chol, ws, wsb, st);
}

int ptm_intercept_stats_f64(const PtmProblem* p, const double* params, int64_t C, double* stats, void* ws,
                            size_t wsb, void* st) {
  if (!p || !stats || !ws || C < 1) return fail(PTM_ERR_NULL, "null argument");
  if (p->dtype != PTM_F64) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  const Plan pl = make_plan<double>(p, C);
  if (wsb < pl.total) return fail(PTM_ERR_WORKSPACE, "workspace too small");
  double* scratch = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + pl.off_scratch);
  return eval_impl<double>(p, params, C, scratch, nullptr, ws, wsb, st, stats);
}
int ptm_intercept_stats_f32(const PtmProblem* p, const float* params, int64_t C, float* stats, void* ws,
                            size_t wsb, void* st) {
  if (!p || !stats || !ws || C < 1) return fail(PTM_ERR_NULL, "null argument");
  if (p->dtype != PTM_F32) return fail(PTM_ERR_DTYPE, "dtype mismatch");
  const Plan pl = make_plan<float>(p, C);
  if (wsb < pl.total) return fail(PTM_ERR_WORKSPACE, "workspace too small");
  float* scratch = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + pl.off_scratch);
  return eval_impl<float>(p, params, C, scratch, nullptr, ws, wsb, st, stats);
}

int ptm_quadform_f64(const PtmProblem* p, const double* params, int64_t C, double* quad, void* st) {
  return quadform_impl<double>(p, params, C, quad, st);
}
int ptm_quadform_f32(const PtmProblem* p, const float* params, int64_t C, float* quad, void* st) {
  return quadform_impl<float>(p, params, C, quad, st);
}

int ptm_loc_precision_all_f64(const PtmProblem* p, const double* params, int64_t C, double* precision,
                              double* chol, void* ws, size_t wsb, void* st) {
  return precision_all_impl<double>(p, params, C, precision, chol, ws, wsb, st);
}
int ptm_loc_precision_all_f32(const PtmProblem* p, const float* params, int64_t C, float* precision,
                              float* chol, void* ws, size_t wsb, void* st) {
  return precision_all_impl<float>(p, params, C, precision, chol, ws, wsb, st);
}

int ptm_problem_shared_designs(const PtmProblem* p) { return (p && p->shared) ? p->T_loc : 0; }

size_t ptm_predict_workspace_bytes(const PtmProblem* p, int64_t S) {
  if (!p || S < 1) return 0;
  return predict_ws_bytes(p, S);
}
int ptm_predict_f64(const PtmProblem* p, const double* params, int64_t S, const double* y_new,
                    const double* x_new, int64_t m, double* z, double* logpdf, double* cdf, void* ws,
                    size_t wsb, void* st) {
  return predict_impl<double>(p, params, S, y_new, x_new, m, z, logpdf, cdf, nullptr, ws, wsb, st);
}
int ptm_predict_f32(const PtmProblem* p, const float* params, int64_t S, const float* y_new,
                    const float* x_new, int64_t m, float* z, float* logpdf, float* cdf, void* ws,
                    size_t wsb, void* st) {
  return predict_impl<float>(p, params, S, y_new, x_new, m, z, logpdf, cdf, nullptr, ws, wsb, st);
}
int ptm_predict_quantile_f64(const PtmProblem* p, const double* params, int64_t S, const double* u,
                             const double* x_new, int64_t m, double* yq, void* ws, size_t wsb, void* st) {
  if (!yq) return fail(PTM_ERR_NULL, "yq is null");
  return predict_impl<double>(p, params, S, u, x_new, m, nullptr, nullptr, nullptr, yq, ws, wsb, st);
}
int ptm_predict_quantile_f32(const PtmProblem* p, const float* params, int64_t S, const float* u,
                             const float* x_new, int64_t m, float* yq, void* ws, size_t wsb, void* st) {
  if (!yq) return fail(PTM_ERR_NULL, "yq is null");
  return predict_impl<float>(p, params, S, u, x_new, m, nullptr, nullptr, nullptr, yq, ws, wsb, st);
}
int ptm_predict_sample_f64(const PtmProblem* p, const double* params, int64_t S, const double* u,
                           const double* x_new, int64_t m, double* ydraw, void* ws, size_t wsb, void* st) {
  return predict_impl<double>(p, params, S, u, x_new, m, nullptr, nullptr, nullptr, ydraw, ws, wsb, st, true);
}
int ptm_predict_sample_f32(const PtmProblem* p, const float* params, int64_t S, const float* u,
                           const float* x_new, int64_t m, float* ydraw, void* ws, size_t wsb, void* st) {
  return predict_impl<float>(p, params, S, u, x_new, m, nullptr, nullptr, nullptr, ydraw, ws, wsb, st, true);
}
int ptm_inverse_f64(const PtmProblem* p, const double* delta, int64_t C, const double* z, int64_t m,
                    double* r, void* st) {
  return inverse_impl<double>(p, delta, C, z, m, r, st);
}
int ptm_inverse_f32(const PtmProblem* p, const float* delta, int64_t C, const float* z, int64_t m,
                    float* r, void* st) {
  return inverse_impl<float>(p, delta, C, z, m, r, st);
}

size_t ptm_pointwise_workspace_bytes(const PtmProblem* p, int64_t S) {
  if (!p || S < 1) return 0;
  const size_t base = ptm_workspace_bytes(p, S);
  return base ? base + (size_t)p->n * 4 * sizeof(double) : 0;
}
int ptm_pointwise_f64(const PtmProblem* p, const double* params, int64_t S, double* lppd, double* pwaic, void* ws,
                      size_t wsb, void* st) {
  return pointwise_impl<double>(p, params, S, lppd, pwaic, ws, wsb, st);
}
int ptm_pointwise_f32(const PtmProblem* p, const float* params, int64_t S, float* lppd, float* pwaic, void* ws,
                      size_t wsb, void* st) {
  return pointwise_impl<float>(p, params, S, lppd, pwaic, ws, wsb, st);
}

size_t ptm_hessian_workspace_bytes(const PtmProblem* p, int64_t C, int32_t block) {
  if (!p || C < 1) return 0;
  HessPlan hp;
  const int rc = p->dtype == PTM_F64 ? make_hess_plan<double>(p, block, C, hp) : make_hess_plan<float>(p, block, C, hp);
  return rc == PTM_OK ? hp.total : 0;
}
int ptm_hessian_block_f64(const PtmProblem* p, int32_t block, int32_t mode, const double* params, int64_t C,
                          double* info, double* chol, void* ws, size_t wsb, void* st) {
  return hessian_impl<double>(p, block, mode, params, C, info, chol, ws, wsb, st);
}
int ptm_hessian_block_f32(const PtmProblem* p, int32_t block, int32_t mode, const float* params, int64_t C,
                          float* info, float* chol, void* ws, size_t wsb, void* st) {
  return hessian_impl<float>(p, block, mode, params, C, info, chol, ws, wsb, st);
}

int ptm_last_launch_count(void) { return g_launches; }
int ptm_last_main_route(void) { return g_last_main_route; }
int ptm_last_xtwx_route(void) { return g_last_xtwx_route; }

int ptm_profile_enable(int on) {
  if (on && !g_ev[0])
    for (int i = 0; i < 4; ++i) PTM_CUDA(cudaEventCreate(&g_ev[i]));
  g_profile = on ? 1 : 0;
  return PTM_OK;
}

int ptm_last_kernel_ms(float* pre_ms, float* main_ms, float* post_ms) {
  if (!g_ev[0]) return fail(PTM_ERR_SHAPE, "profiling was never enabled");
  PTM_CUDA(cudaEventSynchronize(g_ev[3]));
  float a = 0, b = 0, c = 0;
  PTM_CUDA(cudaEventElapsedTime(&a, g_ev[0], g_ev[1]));
  PTM_CUDA(cudaEventElapsedTime(&b, g_ev[1], g_ev[2]));
  PTM_CUDA(cudaEventElapsedTime(&c, g_ev[2], g_ev[3]));
  if (pre_ms) *pre_ms = a;
  if (main_ms) *main_ms = b;
  if (post_ms) *post_ms = c;
  return PTM_OK;
}
const char* ptm_last_error(void) { return g_err.c_str(); }
const char* ptm_version(void) { return "ptm_b200 0.1 (sm_100a)"; }

}  // extern "C"
