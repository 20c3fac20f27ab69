// XLA-FFI (jax.ffi) adapter over the C ABI of libpof_b200.so -- the binding BASELINE.json's north_star asks for
// ("host code stays Python/JAX and calls hand-written sm_100a CUDA kernels through jax.ffi custom calls").
//
// Every handler is a thin shim: it pulls the CUDA stream from the call context, the device pointers and shapes from the
// XLA buffers, scalar options from attributes, and forwards to ONE entry point of include/pof_b200.h.  The workspace
// is an XLA-allocated uint8 result buffer (nothing is allocated inside the library), so the calls are pure functions
// of their operands and can sit inside jax.jit / jax.lax.while_loop; no host synchronisation happens.
//
// NOT BUILT IN THIS IMAGE: the XLA-FFI headers (xla/ffi/api/ffi.h) ship inside jaxlib, which is not installed here
// (SURVEY.md 7.0).  jax_ffi/build.py compiles this file only when `import jaxlib` succeeds and the headers are found;
// the Python side that registers and calls the handlers is jax_ffi/pof_jax.py.  Until it has been compiled against a
// real jaxlib this file is a reviewed sketch of the adapter, not tested code.
#include <cstdint>
#include <string>

#include <cuda_runtime.h>

#include "xla/ffi/api/ffi.h"

#include "../../include/pof_b200.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using F64Out = ffi::ResultBuffer<ffi::F64>;
using U8Out = ffi::ResultBuffer<ffi::U8>;

namespace {

ffi::Error Status(int rc, const char* what) {
  if (rc == 0) return ffi::Error::Success();
  if (rc > 0) return ffi::Error::Internal(std::string(what) + ": CUDA error " + std::to_string(rc));
  return ffi::Error::InvalidArgument(std::string(what) + ": argument error " + std::to_string(rc));
}

// (q+1)^2 entries of the process-noise factor block travel as a float64 ATTRIBUTE array: the library reads them on
// the host (they become kernel parameters), so they must not be a device buffer.
using QL = ffi::Span<const double>;

// ---- pof.parallel_filtsmooth.linear_filtsmooth(x0, dtm, dom)      reference parallel_filtsmooth/__init__.py:5-10
//      operands: x0_mean (D), x0_chol (D,D), H (n,d,D), c (n,d), means_prev (N,D)
//      results : means (N,D), chols (N,D,D), scalars (8), workspace (bytes from pof_workspace_bytes)
ffi::Error FiltSmoothImpl(cudaStream_t stream, int64_t chunk_len, bool calibrate, int64_t flags, QL qL, F64 x0_mean,
                          F64 x0_chol, F64 H, F64 c, F64 means_prev, F64Out means, F64Out chols, F64Out scalars,
                          U8Out ws) {
  const auto dims = H.dimensions();  // (n, d, D)
  if (dims.size() != 3) return ffi::Error::InvalidArgument("H must have shape (n, d, D)");
  const int64_t N = dims[0] + 1;
  const int d = static_cast<int>(dims[1]);
  const int D = static_cast<int>(dims[2]);
  const int q = D / d - 1;
  // the pass compares the new means with the previous ones in place (POF_S_NOT_CLOSE): start from a copy
  cudaError_t e = cudaMemcpyAsync(means->typed_data(), means_prev.typed_data(), sizeof(double) * N * D,
                                  cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return ffi::Error::Internal("cudaMemcpyAsync failed");
  // ctx = NULL: everything in line on XLA's stream (a pof_ctx_t with its side stream can be held in an FFI state object
  // instead, see pof_b200.h)
  return Status(pof_linear_filtsmooth_f64(stream, nullptr, static_cast<uint32_t>(flags), N, d, q, chunk_len, qL.begin(),
                                          x0_mean.typed_data(), x0_chol.typed_data(), H.typed_data(), c.typed_data(),
                                          means->typed_data(), chols->typed_data(), nullptr, nullptr,
                                          calibrate ? 1 : 0, scalars->typed_data(), ws->typed_data(),
                                          ws->size_bytes()),
                "pof_linear_filtsmooth_f64");
}

// ---- the fused loop body: pof.step.ieks_step for a built-in pof.ivp vector field     reference step.py:33-45
//      operands: x0_mean, x0_chol, means_prev (N,D);  results: means, chols, scalars, workspace
ffi::Error IeksIterationImpl(cudaStream_t stream, int64_t ivp_id, int64_t d, int64_t q, int64_t chunk_len,
                             bool calibrate, int64_t flags, double scale0, double scale1, QL qL,
                             ffi::Span<const double> params, F64 x0_mean, F64 x0_chol, F64 means_prev, F64Out means,
                             F64Out chols, F64Out scalars, U8Out ws) {
  const auto dims = means_prev.dimensions();  // (N, D)
  if (dims.size() != 2) return ffi::Error::InvalidArgument("means must have shape (N, D)");
  const int64_t N = dims[0];
  cudaError_t e = cudaMemcpyAsync(means->typed_data(), means_prev.typed_data(), sizeof(double) * N * dims[1],
                                  cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return ffi::Error::Internal("cudaMemcpyAsync failed");
  return Status(pof_ieks_iteration_f64(stream, nullptr, static_cast<uint32_t>(flags), static_cast<int>(ivp_id),
                                       params.begin(), static_cast<int>(params.size()), N, static_cast<int>(d),
                                       static_cast<int>(q), chunk_len, qL.begin(), scale0, scale1,
                                       x0_mean.typed_data(), x0_chol.typed_data(), means->typed_data(),
                                       chols->typed_data(), calibrate ? 1 : 0, scalars->typed_data(), ws->typed_data(),
                                       ws->size_bytes()),
                "pof_ieks_iteration_f64");
}

// ---- vmap(linearize)(om, states[1:]) for the built-in vector fields      reference step.py:12-22, observations.py:35-40
ffi::Error LinearizeImpl(cudaStream_t stream, int64_t ivp_id, int64_t d, int64_t q, double scale0, double scale1,
                         ffi::Span<const double> params, F64 means_t1, F64Out H, F64Out c) {
  const int64_t n = means_t1.dimensions()[0];
  return Status(pof_linearize_ivp_f64(stream, static_cast<int>(ivp_id), params.begin(), static_cast<int>(params.size()),
                                      n, static_cast<int>(d), static_cast<int>(q), scale0, scale1,
                                      means_t1.typed_data(), H->typed_data(), c->typed_data()),
                "pof_linearize_ivp_f64");
}

// ---- the two associative operators (tuples of batched arrays packed per element)   reference filter.py:117-142,
//      smoother.py:53-63
ffi::Error FilterCombineImpl(cudaStream_t stream, int64_t D, F64 e1, F64 e2, F64Out out) {
  return Status(pof_filter_combine_f64(stream, e1.dimensions()[0], static_cast<int>(D), e1.typed_data(),
                                       e2.typed_data(), out->typed_data(), 0u),
                "pof_filter_combine_f64");
}
ffi::Error SmoothCombineImpl(cudaStream_t stream, int64_t D, F64 e1, F64 e2, F64Out out) {
  return Status(pof_smooth_combine_f64(stream, e1.dimensions()[0], static_cast<int>(D), e1.typed_data(),
                                       e2.typed_data(), out->typed_data(), 0u),
                "pof_smooth_combine_f64");
}

// ---- final calibration + E0 projection       reference solver.py:66-71
ffi::Error ProjectImpl(cudaStream_t stream, int64_t d, int64_t q, double scale0, F64 mult, F64 means, F64 chols,
                       F64Out ymean, F64Out ychol) {
  return Status(pof_project_f64(stream, means.dimensions()[0], static_cast<int>(d), static_cast<int>(q), scale0,
                                mult.typed_data(), means.typed_data(), chols.typed_data(), ymean->typed_data(),
                                ychol->typed_data()),
                "pof_project_f64");
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(PofFiltSmooth, FiltSmoothImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("chunk_len")
                                  .Attr<bool>("calibrate")
                                  .Attr<int64_t>("flags")
                                  .Attr<QL>("qL")
                                  .Arg<F64>()   // x0_mean
                                  .Arg<F64>()   // x0_chol
                                  .Arg<F64>()   // H
                                  .Arg<F64>()   // c
                                  .Arg<F64>()   // means_prev
                                  .Ret<F64>()   // means
                                  .Ret<F64>()   // chols
                                  .Ret<F64>()   // scalars
                                  .Ret<ffi::Buffer<ffi::U8>>());  // workspace

XLA_FFI_DEFINE_HANDLER_SYMBOL(PofIeksIteration, IeksIterationImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ivp_id")
                                  .Attr<int64_t>("d")
                                  .Attr<int64_t>("q")
                                  .Attr<int64_t>("chunk_len")
                                  .Attr<bool>("calibrate")
                                  .Attr<int64_t>("flags")
                                  .Attr<double>("scale0")
                                  .Attr<double>("scale1")
                                  .Attr<QL>("qL")
                                  .Attr<ffi::Span<const double>>("params")
                                  .Arg<F64>()   // x0_mean
                                  .Arg<F64>()   // x0_chol
                                  .Arg<F64>()   // means_prev
                                  .Ret<F64>()   // means
                                  .Ret<F64>()   // chols
                                  .Ret<F64>()   // scalars
                                  .Ret<ffi::Buffer<ffi::U8>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(PofLinearize, LinearizeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("ivp_id")
                                  .Attr<int64_t>("d")
                                  .Attr<int64_t>("q")
                                  .Attr<double>("scale0")
                                  .Attr<double>("scale1")
                                  .Attr<ffi::Span<const double>>("params")
                                  .Arg<F64>()
                                  .Ret<F64>()
                                  .Ret<F64>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(PofFilterCombine, FilterCombineImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("D").Arg<F64>()
                                  .Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PofSmoothCombine, SmoothCombineImpl,
                              ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>().Attr<int64_t>("D").Arg<F64>()
                                  .Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PofProject, ProjectImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("d")
                                  .Attr<int64_t>("q")
                                  .Attr<double>("scale0")
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Arg<F64>()
                                  .Ret<F64>()
                                  .Ret<F64>());
