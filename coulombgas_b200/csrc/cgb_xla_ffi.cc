// XLA FFI adapter over the C ABI of libcgb200.so (include/cgb200.h).
//
// NOT built in this image: jax/jaxlib (and therefore xla/ffi/api/ffi.h, found via jax.ffi.include_dir()) are not
// installable here (no wheels, no network).  It is compile-gated on CGB_HAVE_XLA_FFI and kept deliberately
// trivial: every handler only unpacks buffers/attributes and forwards to the tested C entry point.
//   build (where jax is installed):
//     g++ -shared -fPIC -O2 -DCGB_HAVE_XLA_FFI -I$(python -c 'import jax.ffi;print(jax.ffi.include_dir())') \
//         -I../../include cgb_xla_ffi.cc -L.. -lcgb200 -o ../libcgb200_xla.so
// Threading: under jax.pmap XLA calls a handler from one host thread per device, concurrently; the C ABI uses only the
// stream XLA passes in and a per-call workspace.  Its one piece of host state is the per-device event that serialises
// sweeps issued on DIFFERENT streams of one device (the warp kernel's weights are constant-bank operands of a per-device
// block, include/cgb200.h); it is mutex-protected, so concurrent handler calls are safe.
// The PRNG key is an OPERAND: the (2,) uint32 device array the reference carries per device (src/VMC.py:20,
// src/MCMC.py:24) is read by the kernels themselves -- nothing about it is baked into the compiled computation.
// The workspace is an XLA scratch result: its content is undefined at every call, which is fine -- every launcher
// initialises what it reads (queue, counters it uses per call); only the cumulative debug counters at the start of the
// workspace (pivot / float64 fallbacks) are meaningless under XLA and must not be read there.
#ifdef CGB_HAVE_XLA_FFI
#include <cuda_runtime_api.h>
#include "cgb200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error rc_to_error(int rc, const char* what) {
    if (rc == 0) return ffi::Error::Success();
    return ffi::Error(rc < 0 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal,
                      std::string(what) + ": " + cgb_error_string(rc));
}

// logp(x, params, state_idx)  — replaces src/logpsi.py:174-181
static ffi::Error LogpImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::S32> state_idx,
                           ffi::Buffer<ffi::F64> ktable, ffi::Buffer<ffi::F64> params, double L, double twist_x,
                           double twist_y, ffi::ResultBuffer<ffi::F64> logp, ffi::ResultBuffer<ffi::U8> workspace) {
    const auto d = x.dimensions();                       // (B, n, 2)
    const int B = (int)d[0], n = (int)d[1], M = (int)ktable.dimensions()[0];
    const double tw[2] = {twist_x, twist_y};
    return rc_to_error(cgb_logp_f64io(stream, x.typed_data(), state_idx.typed_data(), ktable.typed_data(), M, tw,
                                      params.typed_data(), B, n, L, logp->typed_data(), nullptr, nullptr,
                                      workspace->typed_data(), workspace->size_bytes(), CGB_KERNEL_AUTO), "cgb_logp_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbLogp, LogpImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Attr<double>("L").Attr<double>("twist_x").Attr<double>("twist_y")
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());

// mcmc(logp_fn, x_init, key, mc_steps, mc_stddev) + wrap  — replaces src/MCMC.py:6-40, src/VMC.py:24.
// x is an input_output_alias (donated, src/VMC.py:11): result 0 aliases operand 0.  `key` = jax.random.key_data(key_MCMC),
// a (2,) uint32 device buffer, consumed on the device.
static ffi::Error SweepImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::S32> state_idx,
                            ffi::Buffer<ffi::F64> ktable, ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::U32> key,
                            int64_t mc_steps, double mc_stddev, double L, double twist_x, double twist_y,
                            ffi::ResultBuffer<ffi::F64> x_out, ffi::ResultBuffer<ffi::F64> logp,
                            ffi::ResultBuffer<ffi::S32> naccept, ffi::ResultBuffer<ffi::F64> stats,
                            ffi::ResultBuffer<ffi::U8> workspace) {
    const auto d = x.dimensions();
    const int B = (int)d[0], n = (int)d[1], M = (int)ktable.dimensions()[0];
    const double tw[2] = {twist_x, twist_y};
    if (key.element_count() != 2 || stats->element_count() < CGB_STATS_LEN)
        return ffi::Error(ffi::ErrorCode::kInvalidArgument, "key must be (2,) uint32, stats (>= 8,) float64");
    if (x_out->typed_data() != x.typed_data())          // not aliased by XLA: copy first, then sweep in place
        cudaMemcpyAsync(x_out->typed_data(), x.typed_data(), x.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    return rc_to_error(cgb_mcmc_sweep_f64io(stream, x_out->typed_data(), state_idx.typed_data(), ktable.typed_data(), M,
                                            tw, params.typed_data(), nullptr, nullptr, /*seed=*/0,
                                            const_cast<uint32_t*>(key.typed_data()), (int)mc_steps, mc_stddev, L,
                                            B, n, /*wrap=*/1, logp->typed_data(), naccept->typed_data(), nullptr, nullptr,
                                            stats->typed_data(), workspace->typed_data(), workspace->size_bytes(),
                                            CGB_KERNEL_AUTO), "cgb_mcmc_sweep_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbMcmcSweep, SweepImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Arg<ffi::Buffer<ffi::U32>>()
        .Attr<int64_t>("mc_steps").Attr<double>("mc_stddev").Attr<double>("L").Attr<double>("twist_x").Attr<double>("twist_y")
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::F64>>()
        .Ret<ffi::Buffer<ffi::U8>>());

// potential_energy(x, kappa, G, L, rs)  — replaces src/potential.py:69-77.  `items` is the merged reciprocal table
// built once on the host by cgb_ewald_prepare_host (G and kappa are static in the reference, main.py:170).
static ffi::Error EwaldImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F32> items, double L, double rs,
                            double kappa, int64_t gmax, ffi::ResultBuffer<ffi::F64> V) {
    const auto d = x.dimensions();
    return rc_to_error(cgb_ewald_f64io(stream, x.typed_data(), (int)d[0], (int)d[1], L, rs, kappa, items.typed_data(),
                                       (int)items.dimensions()[0], (int)gmax, V->typed_data(), nullptr), "cgb_ewald_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbEwald, EwaldImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F32>>()
        .Attr<double>("L").Attr<double>("rs").Attr<double>("kappa").Attr<int64_t>("gmax")
        .Ret<ffi::Buffer<ffi::F64>>());

// sampler(params_van, key, batch) / log_prob(params_van, state_idx)  — replace src/sampler.py:30-44 over the Transformer of
// src/autoregressive.py:50-96 (SURVEY section 8 f/N2).  params = the flat vector of cgb200.h (18 M + 4368 doubles).
static ffi::Error VanSampleImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::F64> ktable,
                                ffi::Buffer<ffi::U32> key, int64_t n, ffi::ResultBuffer<ffi::S32> state_idx,
                                ffi::ResultBuffer<ffi::F64> logprob) {
    const auto d = state_idx->dimensions();              // (B, n)
    if (key.element_count() != 2) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "key must be (2,) uint32");
    return rc_to_error(cgb_van_sample_f64io(stream, params.typed_data(), params.element_count(), ktable.typed_data(),
                                            (int)ktable.dimensions()[0], (int)n, (int)d[0], nullptr, /*seed=*/0,
                                            key.typed_data(), state_idx->typed_data(), logprob->typed_data(), nullptr),
                       "cgb_van_sample_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbVanSample, VanSampleImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::U32>>()
        .Attr<int64_t>("n")
        .Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::F64>>());

static ffi::Error VanLogProbImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::F64> ktable,
                                 ffi::Buffer<ffi::S32> state_idx, ffi::ResultBuffer<ffi::F64> logprob) {
    const auto d = state_idx.dimensions();
    return rc_to_error(cgb_van_logprob_f64io(stream, params.typed_data(), params.element_count(), ktable.typed_data(),
                                             (int)ktable.dimensions()[0], (int)d[1], (int)d[0], state_idx.typed_data(),
                                             logprob->typed_data(), nullptr), "cgb_van_logprob_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbVanLogProb, VanLogProbImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>()
        .Ret<ffi::Buffer<ffi::F64>>());
#endif  // CGB_HAVE_XLA_FFI
