// XLA FFI adapter over the C ABI of libcgb200.so (include/cgb200.h).
//
// NOT built in this image: jax/jaxlib (and therefore xla/ffi/api/ffi.h, found via jax.ffi.include_dir()) are not
// installable here (no wheels, no network).  It is compile-gated on CGB_HAVE_XLA_FFI and kept deliberately
// trivial: every handler only unpacks buffers/attributes and forwards to the tested C entry point.
//   build (where jax is installed):
//     g++ -shared -fPIC -O2 -DCGB_HAVE_XLA_FFI -I$(python -c 'import jax.ffi;print(jax.ffi.include_dir())') \
//         -I../../include cgb_xla_ffi.cc -L.. -lcgb200 -o ../libcgb200_xla.so
// Threading: under jax.pmap XLA calls a handler from one host thread per device, concurrently; the C ABI is
// reentrant (no global mutable state, per-call workspace) and uses only the stream XLA passes in.
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
                                      workspace->typed_data(), workspace->size_bytes()), "cgb_logp_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbLogp, LogpImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Attr<double>("L").Attr<double>("twist_x").Attr<double>("twist_y")
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::U8>>());

// mcmc(logp_fn, x_init, key, mc_steps, mc_stddev) + wrap  — replaces src/MCMC.py:6-40, src/VMC.py:24.
// x is an input_output_alias (donated, src/VMC.py:11): result 0 aliases operand 0.
static ffi::Error SweepImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::S32> state_idx,
                            ffi::Buffer<ffi::F64> ktable, ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::U32> key,
                            int64_t mc_steps, double mc_stddev, double L, double twist_x, double twist_y, int64_t seed_hi,
                            int64_t seed_lo, ffi::ResultBuffer<ffi::F64> x_out, ffi::ResultBuffer<ffi::F64> logp,
                            ffi::ResultBuffer<ffi::S32> naccept, ffi::ResultBuffer<ffi::U8> workspace) {
    const auto d = x.dimensions();
    const int B = (int)d[0], n = (int)d[1], M = (int)ktable.dimensions()[0];
    const double tw[2] = {twist_x, twist_y};
    if (x_out->typed_data() != x.typed_data())          // not aliased by XLA: copy first, then sweep in place
        cudaMemcpyAsync(x_out->typed_data(), x.typed_data(), x.size_bytes(), cudaMemcpyDeviceToDevice, stream);
    // The jax PRNG key (2 x uint32) lives on the device; the Python side passes its two words as attributes
    // (seed_hi/seed_lo) so that no device->host copy is needed here.  `key` is kept as an operand for tracing.
    (void)key;
    const uint64_t seed = ((uint64_t)(uint32_t)seed_hi << 32) | (uint32_t)seed_lo;
    return rc_to_error(cgb_mcmc_sweep_f64io(stream, x_out->typed_data(), state_idx.typed_data(), ktable.typed_data(), M,
                                            tw, params.typed_data(), nullptr, nullptr, seed, (int)mc_steps, mc_stddev, L,
                                            B, n, /*wrap=*/1, logp->typed_data(), naccept->typed_data(), nullptr, nullptr,
                                            workspace->typed_data(), workspace->size_bytes()), "cgb_mcmc_sweep_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbMcmcSweep, SweepImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::S32>>().Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Arg<ffi::Buffer<ffi::U32>>()
        .Attr<int64_t>("mc_steps").Attr<double>("mc_stddev").Attr<double>("L").Attr<double>("twist_x").Attr<double>("twist_y")
        .Attr<int64_t>("seed_hi").Attr<int64_t>("seed_lo")
        .Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::F64>>().Ret<ffi::Buffer<ffi::S32>>().Ret<ffi::Buffer<ffi::U8>>());

// potential_energy(x, kappa, G, L, rs)  — replaces src/potential.py:69-77.  `items` is the merged reciprocal table
// built once on the host by cgb_ewald_prepare_host (G and kappa are static in the reference, main.py:170).
static ffi::Error EwaldImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> x, ffi::Buffer<ffi::F32> items, double L, double rs,
                            double kappa, int64_t gmax, ffi::ResultBuffer<ffi::F64> V) {
    const auto d = x.dimensions();
    return rc_to_error(cgb_ewald_f64io(stream, x.typed_data(), (int)d[0], (int)d[1], L, rs, kappa, items.typed_data(),
                                       (int)items.dimensions()[0], (int)gmax, V->typed_data()), "cgb_ewald_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbEwald, EwaldImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F32>>()
        .Attr<double>("L").Attr<double>("rs").Attr<double>("kappa").Attr<int64_t>("gmax")
        .Ret<ffi::Buffer<ffi::F64>>());

// sampler(params_van, key, batch) / log_prob(params_van, state_idx)  — replace src/sampler.py:30-44 over the Transformer of
// src/autoregressive.py:50-96 (SURVEY section 8 f/N2).  params = the flat vector of cgb200.h (18 M + 4368 doubles).
static ffi::Error VanSampleImpl(cudaStream_t stream, ffi::Buffer<ffi::F64> params, ffi::Buffer<ffi::F64> ktable, int64_t n,
                                int64_t seed_hi, int64_t seed_lo, ffi::ResultBuffer<ffi::S32> state_idx,
                                ffi::ResultBuffer<ffi::F64> logprob) {
    const auto d = state_idx->dimensions();              // (B, n)
    const uint64_t seed = ((uint64_t)(uint32_t)seed_hi << 32) | (uint32_t)seed_lo;
    return rc_to_error(cgb_van_sample_f64io(stream, params.typed_data(), params.element_count(), ktable.typed_data(),
                                            (int)ktable.dimensions()[0], (int)n, (int)d[0], nullptr, seed,
                                            state_idx->typed_data(), logprob->typed_data(), nullptr), "cgb_van_sample_f64io");
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(CgbVanSample, VanSampleImpl,
    ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
        .Arg<ffi::Buffer<ffi::F64>>().Arg<ffi::Buffer<ffi::F64>>()
        .Attr<int64_t>("n").Attr<int64_t>("seed_hi").Attr<int64_t>("seed_lo")
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
