// ptm_xla_ffi.cc -- XLA FFI handlers over the C ABI of include/ptm_b200.h: the JAX side of
// the drop-in boundary (reference seams model/model.py:477,1369 -- the model
// interface -- and 1001,1038,1089,1126,1208 -- chol_info_fn).
//
// One handler per entry point and precision.  Operands: the [C, P] chain-state block JAX
// hands over (the chain batch arrives through vmap: integration/fused.py registers every
// call with vmap_method="expand_dims" / "broadcast_all" so that ONE launch sees all chains).
// Attributes: `handle` = the PtmProblem* from ptm_problem_create as an int64, `term` where an
// entry point needs it.  Results: the outputs plus the caller-sized device workspace
// (ptm_workspace_bytes) -- XLA allocates it, the library never does.  The stream is XLA's.
// Errors travel back as ffi::Error with ptm_last_error()'s text; no exceptions.
//
// Build (on a machine with jax):
//   g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax; print(jax.ffi.include_dir())")
//       -I/usr/local/cuda/include -Iinclude liesel_ptm_b200/csrc/ptm_xla_ffi.cc
//       -Lliesel_ptm_b200 -lptm_b200 -o liesel_ptm_b200/libptm_xla_ffi.so
// This image has no jax: the TU is compiled only when the header is found; tests type-check
// it against tests/ffi_stub (-DPTM_FFI_STUB), which mirrors the part of the API used here.
#if defined(PTM_FFI_STUB) || __has_include("xla/ffi/api/ffi.h")

#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>

#include "xla/ffi/api/ffi.h"
#include "ptm_b200.h"

namespace ffi = xla::ffi;

namespace {

inline PtmProblem* problem_of(int64_t handle) { return reinterpret_cast<PtmProblem*>(static_cast<intptr_t>(handle)); }

inline ffi::Error status(int rc) {
  if (rc == PTM_OK) return ffi::Error::Success();
  return ffi::Error(rc == PTM_ERR_CUDA ? ffi::ErrorCode::kInternal : ffi::ErrorCode::kInvalidArgument,
                    std::string("ptm_b200: ") + ptm_last_error());
}

// params must be [..., P] row-major (ravel_pytree rows; vmap with "expand_dims" delivers
// [C, 1, P]): every leading dimension is chain batch
template <ffi::DataType DT>
inline bool chain_block(const ffi::Buffer<DT>& params, const PtmProblem* prob, int64_t* C) {
  const auto dims = params.dimensions();
  if (dims.size() < 1 || dims[dims.size() - 1] != ptm_num_params(prob)) return false;
  int64_t c = 1;
  for (size_t i = 0; i + 1 < dims.size(); ++i) c *= dims[i];
  if (c < 1) return false;
  *C = c;
  return true;
}
#define PTM_FFI_CHAINS(params, prob, C)                                                       \
  int64_t C = 0;                                                                              \
  if (!chain_block(params, prob, &C))                                                         \
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "ptm_b200: params must have shape [..., P]")

// The C ABI has one symbol per precision; pick it by the buffer's element type.
template <ffi::DataType DT> struct Abi;
template <> struct Abi<ffi::F64> {
  static constexpr auto logp = ptm_logp_f64;
  static constexpr auto logp_grad = ptm_logp_grad_f64;
  static constexpr auto logp_grad_block = ptm_logp_grad_block_f64;
  static constexpr auto xtwx = ptm_xtwx_f64;
  static constexpr auto loc_precision = ptm_loc_precision_f64;
  static constexpr auto loc_precision_all = ptm_loc_precision_all_f64;
  static constexpr auto intercept_stats = ptm_intercept_stats_f64;
  static constexpr auto quadform = ptm_quadform_f64;
  static constexpr auto hessian_block = ptm_hessian_block_f64;
};
template <> struct Abi<ffi::F32> {
  static constexpr auto logp = ptm_logp_f32;
  static constexpr auto logp_grad = ptm_logp_grad_f32;
  static constexpr auto logp_grad_block = ptm_logp_grad_block_f32;
  static constexpr auto xtwx = ptm_xtwx_f32;
  static constexpr auto loc_precision = ptm_loc_precision_f32;
  static constexpr auto loc_precision_all = ptm_loc_precision_all_f32;
  static constexpr auto intercept_stats = ptm_intercept_stats_f32;
  static constexpr auto quadform = ptm_quadform_f32;
  static constexpr auto hessian_block = ptm_hessian_block_f32;
};

// model.log_prob(state): logprob.py:36-41, 50-69
template <ffi::DataType DT>
ffi::Error Logp(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle, ffi::ResultBuffer<DT> logp,
                ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::logp(prob, params.typed_data(), C, logp->typed_data(), ws->typed_data(),
                              ws->element_count(), stream));
}

// value_and_grad in one fused sweep: logprob.py:104-140, iwls_fixed.py:91-116 (custom_vjp fwd)
template <ffi::DataType DT>
ffi::Error LogpGrad(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle, ffi::ResultBuffer<DT> logp,
                    ffi::ResultBuffer<DT> grad, ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::logp_grad(prob, params.typed_data(), C, logp->typed_data(), grad->typed_data(),
                                   ws->typed_data(), ws->element_count(), stream));
}

// the packed [C, 1 + P] block of the observation-axis all-reduce (jax.lax.psum on the result)
template <ffi::DataType DT>
ffi::Error LogpGradBlock(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle,
                         ffi::ResultBuffer<DT> block, ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::logp_grad_block(prob, params.typed_data(), C, block->typed_data(), ws->typed_data(),
                                         ws->element_count(), stream));
}

// GaussianLocCholInfo / GaussianScaleCholInfo: iwls_proposals.py:55-89, 118-144
template <ffi::DataType DT>
ffi::Error Xtwx(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle, int64_t term,
                ffi::ResultBuffer<DT> xtwx, ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::xtwx(prob, (int32_t)term, params.typed_data(), C, xtwx->typed_data(), ws->typed_data(),
                              ws->element_count(), stream));
}

template <ffi::DataType DT>
ffi::Error LocPrecision(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle, int64_t term,
                        ffi::ResultBuffer<DT> precision, ffi::ResultBuffer<DT> chol,
                        ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::loc_precision(prob, (int32_t)term, params.typed_data(), C, precision->typed_data(),
                                       chol->typed_data(), ws->typed_data(), ws->element_count(), stream));
}

// every location term of a sampler sweep from ONE pass (the weights depend only on sigma)
template <ffi::DataType DT>
ffi::Error LocPrecisionAll(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle,
                           ffi::ResultBuffer<DT> precision, ffi::ResultBuffer<DT> chol,
                           ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::loc_precision_all(prob, params.typed_data(), C, precision->typed_data(),
                                           chol->typed_data(), ws->typed_data(), ws->element_count(), stream));
}

// LocationIntercept / LogScaleIntercept sufficient statistics: predictor.py:80-100, 126-130
template <ffi::DataType DT>
ffi::Error InterceptStats(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle,
                          ffi::ResultBuffer<DT> stats, ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::intercept_stats(prob, params.typed_data(), C, stats->typed_data(), ws->typed_data(),
                                         ws->element_count(), stream));
}

// beta' K beta of the tau^2 Gibbs kernels: gam/kernel.py:29-39
template <ffi::DataType DT>
ffi::Error Quadform(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle, ffi::ResultBuffer<DT> quad) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::quadform(prob, params.typed_data(), C, quad->typed_data(), stream));
}

// observed information of a parameter block (jacfwd(grad), logprob.py:108-113; CholInfo /
// PTMCholInfoFixed / kernel.IWLSKernel._chol_info: iwls_proposals.py:168-184, 227-245,
// 279-298, 333-364, kernel.py:240-258)
template <ffi::DataType DT>
ffi::Error HessianBlock(cudaStream_t stream, ffi::Buffer<DT> params, int64_t handle, int64_t block, int64_t mode,
                        ffi::ResultBuffer<DT> hess, ffi::ResultBuffer<DT> chol, ffi::ResultBuffer<ffi::U8> ws) {
  PtmProblem* prob = problem_of(handle);
  PTM_FFI_CHAINS(params, prob, C);
  return status(Abi<DT>::hessian_block(prob, (int32_t)block, (int32_t)mode, params.typed_data(), C,
                                       hess->typed_data(), chol->typed_data(), ws->typed_data(),
                                       ws->element_count(), stream));
}

}  // namespace

#define PTM_BIND_BEGIN(DT) \
  ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Arg<ffi::Buffer<DT>>().Attr<int64_t>("handle")
#define PTM_WS .Ret<ffi::Buffer<ffi::U8>>()

#define PTM_DEFINE_FOR(SFX, DT)                                                                             \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmLogp##SFX, Logp<DT>, PTM_BIND_BEGIN(DT).Ret<ffi::Buffer<DT>>() PTM_WS);   \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmLogpGrad##SFX, LogpGrad<DT>,                                             \
                                PTM_BIND_BEGIN(DT).Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>() PTM_WS);    \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmLogpGradBlock##SFX, LogpGradBlock<DT>,                                   \
                                PTM_BIND_BEGIN(DT).Ret<ffi::Buffer<DT>>() PTM_WS);                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmXtwx##SFX, Xtwx<DT>,                                                     \
                                PTM_BIND_BEGIN(DT).Attr<int64_t>("term").Ret<ffi::Buffer<DT>>() PTM_WS);     \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(                                                                            \
      PtmLocPrecision##SFX, LocPrecision<DT>,                                                               \
      PTM_BIND_BEGIN(DT).Attr<int64_t>("term").Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>() PTM_WS);        \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmLocPrecisionAll##SFX, LocPrecisionAll<DT>,                               \
                                PTM_BIND_BEGIN(DT).Ret<ffi::Buffer<DT>>().Ret<ffi::Buffer<DT>>() PTM_WS);    \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmInterceptStats##SFX, InterceptStats<DT>,                                 \
                                PTM_BIND_BEGIN(DT).Ret<ffi::Buffer<DT>>() PTM_WS);                           \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmQuadform##SFX, Quadform<DT>, PTM_BIND_BEGIN(DT).Ret<ffi::Buffer<DT>>()); \
  XLA_FFI_DEFINE_HANDLER_SYMBOL(PtmHessianBlock##SFX, HessianBlock<DT>,                                     \
                                PTM_BIND_BEGIN(DT)                                                          \
                                    .Attr<int64_t>("block")                                                 \
                                    .Attr<int64_t>("mode")                                                  \
                                    .Ret<ffi::Buffer<DT>>()                                                 \
                                    .Ret<ffi::Buffer<DT>>() PTM_WS)

PTM_DEFINE_FOR(F64, ffi::F64);
PTM_DEFINE_FOR(F32, ffi::F32);

#else
// no XLA FFI headers in this environment: nothing to compile (see the header comment)
#endif
