// C ABI of libcgb200.so (include/cgb200.h): argument checking, host-side table preparation, launches.
#include <math.h>
#include <string.h>
#include <map>
#include <utility>
#include "cgb_internal.h"

using namespace cgb;

extern "C" {

int cgb_abi_version(void) { return CGB_ABI_VERSION; }

const char* cgb_error_string(int code) {
    switch (code) {
        case CGB_OK: return "ok";
        case CGB_E_BADARG: return "bad argument";
        case CGB_E_UNSUPPORTED: return "unsupported shape or table";
        case CGB_E_WORKSPACE: return "workspace missing or too small";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
    }
}

int cgb_ktable_twist_host(const double* kt, int M, double twist_out[2]) {
    if (!kt || M <= 0 || !twist_out) return CGB_E_BADARG;
    const double t0 = kt[0] - rint(kt[0]), t1 = kt[1] - rint(kt[1]);
    for (int j = 0; j < M; ++j) {
        const double a = kt[2 * j] - t0, b = kt[2 * j + 1] - t1;
        if (fabs(a - rint(a)) > 1e-9 || fabs(b - rint(b)) > 1e-9) return CGB_E_UNSUPPORTED;
        if (fabs(rint(a)) > 4096 || fabs(rint(b)) > 4096) return CGB_E_UNSUPPORTED;
    }
    twist_out[0] = t0;
    twist_out[1] = t1;
    return CGB_OK;
}

int cgb_ewald_prepare_host(const int64_t* G, int nG, double kappa, float* items_out, int* n_items, int* gmax) {
    if (!G || nG <= 0 || !items_out || !n_items || !gmax || !(kappa > 0)) return CGB_E_BADARG;
    std::map<std::pair<int, int>, std::pair<double, double>> acc;   // (gx>=0, gy>=0) -> (w_plus, w_minus)
    int gm = 0;
    const double pi = 3.14159265358979323846;
    for (int t = 0; t < nG; ++t) {
        long gx = (long)G[2 * t], gy = (long)G[2 * t + 1];
        if (gx == 0 && gy == 0) return CGB_E_BADARG;                 // kpoints() excludes G = 0
        if (labs(gx) > 255 || labs(gy) > 255) return CGB_E_UNSUPPORTED;
        const double gn = sqrt((double)(gx * gx + gy * gy));
        const double w = erfc(pi * gn / kappa) / gn;                 // g_G, src/potential.py:58 (dim 2)
        if (gx < 0 || (gx == 0 && gy < 0)) { gx = -gx; gy = -gy; }   // cos is even in G
        auto& slot = acc[{(int)gx, (int)labs(gy)}];
        if (gy >= 0) slot.first += w; else slot.second += w;
        if (gx > gm) gm = (int)gx;
        if (labs(gy) > gm) gm = (int)labs(gy);
    }
    int k = 0;
    for (const auto& kv : acc) {
        items_out[4 * k + 0] = (float)kv.first.first;
        items_out[4 * k + 1] = (float)kv.first.second;
        items_out[4 * k + 2] = (float)kv.second.first;
        items_out[4 * k + 3] = (float)kv.second.second;
        ++k;
    }
    *n_items = k;
    *gmax = gm;
    return CGB_OK;
}

// The FP32 Slater LU of the warp kernel is accepted while its a-posteriori error indicator sum_k k / |u_kk|^2 stays below
// this value (calibrated on the B200 against the float64 LU: tools/calib_slater32.py, profiles/r02_slater32_calibration.txt);
// above it the evaluation is redone in float64.
static const float kSlater32ErrsumMax = 300.0f;

static size_t workspace_total(int B) { return ws_total_bytes(B); }   // layout: cgb_internal.h (WS_*)

int cgb_logp_workspace_bytes(int B, int n, size_t* bytes) {
    if (!bytes || B < 0) return CGB_E_BADARG;
    if (n < 2 || n > CGB_MAX_N) return CGB_E_UNSUPPORTED;
    *bytes = workspace_total(B);
    return CGB_OK;
}
int cgb_mcmc_workspace_bytes(int B, int n, size_t* bytes) { return cgb_logp_workspace_bytes(B, n, bytes); }

static int run_sweep(void* stream, const double* x, double* x_out, const int32_t* state_idx, const double* ktable,
                     int M, const double twist[2], const double* params, const double* noise, const double* uniforms,
                     uint64_t seed, uint32_t* key_dev, int mc_steps, double sigma, double L, int B, int n, int wrap,
                     double* out_logp, int32_t* out_naccept, uint8_t* acc_trace, double* logp_trace, double* out_parts,
                     float* out_jac, double* out_stats, void* workspace, size_t workspace_bytes, unsigned flags) {
    if (B == 0) return CGB_OK;
    if (!x || !state_idx || !ktable || !twist || !params || B < 0 || M <= 0 || !(L > 0) || mc_steps < 0)
        return CGB_E_BADARG;
    if ((noise == nullptr) != (uniforms == nullptr)) return CGB_E_BADARG;
    if (n < 2 || n > CGB_MAX_N) return CGB_E_UNSUPPORTED;
    const unsigned kk = flags & CGB_KERNEL_MASK;
    if (kk > CGB_KERNEL_SPEC8) return CGB_E_BADARG;
    if (!workspace || workspace_bytes < workspace_total(B) || ((uintptr_t)workspace & 15)) return CGB_E_WORKSPACE;
    if ((uintptr_t)params & 15) return CGB_E_BADARG;            // float64 parameters are read as 16-byte pairs
    cudaStream_t st = (cudaStream_t)stream;
    SweepParams p;
    p.n_sms = 148;
    p.x = x; p.x_out = x_out; p.state_idx = state_idx; p.ktable = ktable; p.M = M;
    p.tw0 = twist[0]; p.tw1 = twist[1]; p.wpack = (const float*)workspace; p.params = params;
    p.noise = noise; p.uniforms = uniforms; p.seed = seed; p.key_dev = key_dev; p.mc_steps = mc_steps; p.sigma = sigma;
    p.L = L; p.B = B; p.n = n; p.wrap = wrap; p.out_logp = out_logp; p.out_naccept = out_naccept;
    p.out_acc_trace = acc_trace; p.out_logp_trace = logp_trace; p.out_parts = out_parts; p.out_jac = out_jac;
    p.out_stats = out_stats;
    p.force_pivot = (flags & CGB_FLAG_FORCE_PIVOT) ? 1 : 0;
    p.force_f64_slater = (flags & CGB_FLAG_FORCE_F64_SLATER) ? 1 : 0;
    p.slater32_errsum_max = kSlater32ErrsumMax;
    p.advance_key = (flags & CGB_FLAG_ADVANCE_KEY) ? 1 : 0;
    char* ws = (char*)workspace;
    p.fallback_counter = (unsigned long long*)(ws + WS_CTR);
    p.wconst = ws + WS_WCONST;
    p.wshared = (const double*)(ws + WS_WSHARED);
    p.qctr = (unsigned long long*)(ws + WS_CTR + 64);
    p.prof = (unsigned long long*)(ws + WS_CTR + 128);
    p.qcap = ws_qcap(B);
    p.queue = (unsigned long long*)(ws + WS_DYN);
    {
        size_t o = WS_DYN + ws_round(8 * (size_t)p.qcap);
        p.done = (int*)(ws + o); o += ws_round(sizeof(int) * (size_t)B);
        p.ws_logp = (double*)(ws + o); o += ws_round(sizeof(double) * (size_t)B);
        p.ws_nacc = (int*)(ws + o);
    }
    // kernel choice (measured: tools/time_small_batches.py, profiles/r01_small_batches_spec.txt).
    //  * throughput: the warp-per-walker kernel with its ready queue (cgb_wsweep.cu) for the instantiated n <= 32 shapes;
    //  * latency: when the batch cannot fill the machine with one warp per walker (B200: 148 SMs x 12 warps) the idle
    //    SMs are spent on SPECULATION -- the proposals of the next 2-3 steps of every chain are evaluated at once
    //    (bit-identical trajectory): B <= 64: clusters of 8 CTAs of the CTA-per-walker kernel (cgb_sweep.cu), B <= 148:
    //    8 warps per walker in the warp kernel, B <= 444: 4 warps per walker;
    //  * in between (n = 13: 444 < B <= 768) and for every other n: the CTA-per-walker kernel, speculative (clusters of
    //    8 or 4) whenever the clusters of the whole batch are resident at once (n = 57: B <= 74).
    // flags & CGB_KERNEL_MASK overrides the choice (tests, A/B timings); an override that has no instance for this n
    // falls back to the CTA-per-walker kernel.
    const bool spec_ok = wsweep_supported(n) && mc_steps > 0;
    const int sms = wsweep_spec_capacity(n) / 3;
    bool use_warp = wsweep_supported(n) && B > (n <= 16 ? 768 : 384);
    p.cspec = -1;
    p.seg_len = 0;
    p.seg_override = (flags & CGB_FLAG_FORCE_SEGMENTS) ? 2 : ((flags & CGB_FLAG_NO_SEGMENTS) ? 0 : -1);
    p.spec = 0;
    if (spec_ok && B > 64) p.spec = B <= sms ? 8 : (B <= 3 * sms ? 4 : 0);
    if (kk == CGB_KERNEL_CTA || kk == CGB_KERNEL_CTA4 || kk == CGB_KERNEL_CTA8) {
        use_warp = false; p.spec = 0; p.cspec = kk == CGB_KERNEL_CTA ? 0 : (kk == CGB_KERNEL_CTA4 ? 4 : 8);
    } else if (kk == CGB_KERNEL_WARP) {
        use_warp = wsweep_supported(n); p.spec = 0;
    } else if (kk == CGB_KERNEL_SPEC4 || kk == CGB_KERNEL_SPEC8) {
        p.spec = !spec_ok ? 0 : (kk == CGB_KERNEL_SPEC8 ? 8 : 4);
    }
    if (p.spec) use_warp = true;
    if (spec_ok && B <= 64 && kk == CGB_KERNEL_AUTO) use_warp = false;
    if (use_warp) return launch_wsweep(p, ws + WS_WCONST, st);
    int rc = launch_prep_params(p, (float*)workspace, st);
    if (rc) return rc;
    return launch_sweep(p, st);
}

int cgb_logp_f64io(void* stream, const double* x, const int32_t* state_idx, const double* ktable, int M,
                   const double twist[2], const double* params, int B, int n, double L, double* out_logp,
                   double* out_parts, float* out_jac, void* workspace, size_t workspace_bytes, unsigned flags) {
    if (!out_logp && B > 0) return CGB_E_BADARG;
    if (flags & CGB_FLAG_ADVANCE_KEY) return CGB_E_BADARG;
    return run_sweep(stream, x, nullptr, state_idx, ktable, M, twist, params, nullptr, nullptr, 0, nullptr, 0, 0.0, L, B, n, 0,
                     out_logp, nullptr, nullptr, nullptr, out_parts, out_jac, nullptr, workspace, workspace_bytes, flags);
}

int cgb_mcmc_sweep_f64io(void* stream, double* x_inout, const int32_t* state_idx, const double* ktable, int M,
                         const double twist[2], const double* params, const double* noise, const double* uniforms,
                         uint64_t seed, uint32_t* key_dev, int mc_steps, double mc_stddev, double L, int B, int n, int wrap,
                         double* out_logp, int32_t* out_naccept, uint8_t* out_accept_trace, double* out_logp_trace,
                         double* out_stats, void* workspace, size_t workspace_bytes, unsigned flags) {
    if (mc_steps < 1) return CGB_E_BADARG;
    if ((flags & CGB_FLAG_ADVANCE_KEY) && !key_dev) return CGB_E_BADARG;
    if (key_dev && ((uintptr_t)key_dev & 3)) return CGB_E_BADARG;
    if (out_stats && ((uintptr_t)out_stats & 7)) return CGB_E_BADARG;
    return run_sweep(stream, x_inout, x_inout, state_idx, ktable, M, twist, params, noise, uniforms, seed, key_dev, mc_steps,
                     mc_stddev, L, B, n, wrap, out_logp, out_naccept, out_accept_trace, out_logp_trace, nullptr,
                     nullptr, out_stats, workspace, workspace_bytes, flags);
}

int cgb_ewald_f64io(void* stream, const double* x, int B, int n, double L, double rs, double kappa,
                    const float* items, int n_items, int gmax, double* out_V, double* out_stats) {
    if (B == 0) return CGB_OK;
    if (!x || !items || !out_V || B < 0 || n_items <= 0 || !(L > 0) || !(kappa > 0)) return CGB_E_BADARG;
    if (n < 2 || n > 128 || gmax < 1 || gmax > 255) return CGB_E_UNSUPPORTED;
    if (((uintptr_t)items & 15)) return CGB_E_BADARG;
    if (out_stats && ((uintptr_t)out_stats & 7)) return CGB_E_BADARG;
    EwaldParams p;
    p.x = x; p.B = B; p.n = n; p.L = L; p.rs = rs; p.kappa = kappa;
    p.items = (const float4*)items; p.n_items = n_items; p.gmax = gmax; p.out = out_V; p.stats = out_stats;
    return launch_ewald(p, (cudaStream_t)stream);
}

size_t cgb_van_nparams(int M) { return M > 0 ? van_nparams(M) : 0; }

static int run_van(void* stream, const double* params, size_t nparams, const double* ktable, int M, int n, int B,
                   const int32_t* state_in, const double* gumbel, uint64_t seed, const uint32_t* key_dev, int32_t* out_state,
                   double* out_logprob, double* out_logits) {
    if (B == 0) return CGB_OK;
    if (!params || !ktable || B < 0 || M <= 0 || n <= 0) return CGB_E_BADARG;
    if (n > CGB_MAX_N || M < n || M > 256) return CGB_E_UNSUPPORTED;
    if (nparams != van_nparams(M)) return CGB_E_UNSUPPORTED;      // nlayers 2, modelsize 16, nheads 4, nhidden 32 only
    VanParams p;
    p.params = params; p.ktable = ktable; p.M = M; p.n = n; p.B = B; p.state_in = state_in; p.gumbel = gumbel;
    p.seed = seed; p.key_dev = key_dev; p.out_state = out_state; p.out_logprob = out_logprob; p.out_logits = out_logits;
    return launch_van(p, (cudaStream_t)stream);
}

int cgb_van_sample_f64io(void* stream, const double* params, size_t nparams, const double* ktable, int M, int n, int B,
                         const double* gumbel, uint64_t seed, const uint32_t* key_dev, int32_t* out_state_idx,
                         double* out_logprob, double* out_logits) {
    if (B != 0 && !out_state_idx) return CGB_E_BADARG;
    return run_van(stream, params, nparams, ktable, M, n, B, nullptr, gumbel, seed, key_dev, out_state_idx, out_logprob, out_logits);
}

int cgb_van_logprob_f64io(void* stream, const double* params, size_t nparams, const double* ktable, int M, int n, int B,
                          const int32_t* state_idx, double* out_logprob, double* out_logits) {
    if (B != 0 && (!state_idx || !out_logprob)) return CGB_E_BADARG;
    return run_van(stream, params, nparams, ktable, M, n, B, state_idx, nullptr, 0, nullptr, nullptr, out_logprob, out_logits);
}

int cgb_microbench(void* stream, int kind, int blocks, int iters, float* out) {
    if (kind < 0 || kind > 2 || blocks <= 0 || iters <= 0 || !out) return CGB_E_BADARG;
    return launch_microbench(kind, blocks, iters, out, (cudaStream_t)stream);
}

int cgb_debug_rng(void* stream, uint64_t seed, const uint32_t* key_dev, int B, int S, int n, float* out_normal,
                  double* out_uniform, const uint32_t* kat_in, uint32_t* kat_out, int n_kat) {
    if (B < 0 || S < 0 || n < 0 || n_kat < 0 || (n_kat > 0 && (!kat_in || !kat_out))) return CGB_E_BADARG;
    return launch_rng_dump(seed, key_dev, B, S, n, out_normal, out_uniform, kat_in, kat_out, n_kat, (cudaStream_t)stream);
}

}  // extern "C"
