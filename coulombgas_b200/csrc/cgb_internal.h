// Internal declarations shared by the translation units of libcgb200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include "../../include/cgb200.h"

namespace cgb {

// FP32 weight pack produced by prep_params_kernel (floats)
enum : int {
    W_WT = 0,      // Wt*log2e   [5][16]   two-particle layer       (src/flow.py:46)
    W_BT = 80,     // bt*log2e   [16]
    W_W0C = 96,    // W0[4:9]    [5][16]   rows of W0 that see mean_j t0_ij (flow.py:34-36)
    W_B0 = 176,    // b0         [16]
    W_W1A = 192,   // W1[0:16]   [16][16]  acts on h_i
    W_W1B = 448,   // W1[16:32]  [16][16]  acts on mean_k h_k
    W_W1C = 704,   // W1[32:48]  [16][16]  acts on mean_j t_ij
    W_B1 = 960,    // b1         [16]
    W_WF = 976,    // Wf         [16][2]
    W_TOTAL = 1008
};

struct SweepParams {
    const double* x;          // (B,n,2) initial walkers
    double* x_out;            // (B,n,2) final walkers (may alias x); unused when mc_steps == 0
    const int32_t* state_idx; // (B,n)
    const double* ktable;     // (M,2)
    int M;
    double tw0, tw1;
    const float* wpack;       // W_TOTAL floats
    const double* params;     // the caller's flat float64 parameters (forward pass of the single-particle stream)
    const double* noise;      // (S,B,n,2) or null
    const double* uniforms;   // (S,B) or null
    uint64_t seed;            // throughput mode: Philox key (host value) ...
    const uint32_t* key_dev;  // ... or, when non-null, a device-resident (2,) uint32 PRNG key read by the kernel (key[0] << 32 | key[1])
    int mc_steps;
    double sigma;
    double L;
    int B, n, wrap;
    double* out_logp;         // (B)
    int32_t* out_naccept;     // (B) or null
    uint8_t* out_acc_trace;   // (S,B) or null
    double* out_logp_trace;   // (S+1,B) or null
    double* out_parts;        // (B,2) or null
    float* out_jac;           // (B,2n,2n) or null
    int force_pivot;          // debug: always use the pivoted scalar LU for the Jacobian
    int force_f64_slater;     // debug: always use the float64 Slater LU (warp kernel)
    float slater32_errsum_max;    // FP32 Slater LU is kept while sum_k k / |u_kk|^2 stays below this (else redone in float64)
    double* out_stats;        // CGB_STATS_LEN doubles or null: [0] += accepted moves per walker (zeroed by the prologue)
    int advance_key;          // prologue: key_dev += golden-ratio increment before the sweep reads it
    unsigned long long* fallback_counter;   // evaluations redone with the pivoted LU (in the workspace)
    // ---- warp-per-walker kernel only (cgb_wsweep.cu): per-call scheduler state in the caller's workspace
    const void* wconst;           // staging copy of the constant-bank weight block (device)
    const double* wshared;        // float64 weights of the single-particle stream in the CTA-shared layout (TMA source, 16-byte aligned)
    unsigned long long* qctr;     // [0] queue head, [1] queue tail, [2] spin diagnostics (initialised per call)
    unsigned long long* queue;    // (qcap) ready-queue slots
    unsigned qcap;                // power of two >= 2*B
    int* done;                    // (B) steps completed per walker
    double* ws_logp;              // (B) current log p per walker
    int* ws_nacc;                 // (B) accepted moves so far
    int seg_len;                  // CTA kernel: steps per chain segment (0 = whole chains, static assignment); set by the launcher
    int seg_override;             // -1 = launcher's choice, else forced segment length (tests)
    int cspec;                    // CTA kernel's speculative mode: -1 = auto (widest tree that fits), 0 = off, 4/8 = CTAs per walker
    unsigned long long* prof;     // developer builds (CGBW_PROF): 12 per-phase cycle sums in the workspace, else unused
    unsigned n_sms;               // multiprocessors of the device (CTA i starts on SM i % n_sms)
    int spec;                     // 4 / 8: speculative mode, warps per walker (one CTA evaluates the next 2-3 / 3-4 steps' proposals at once)
};

// workspace layout (bytes); the B-dependent arrays start at WS_DYN
enum : size_t {
    WS_WPACK = 0,           // W_TOTAL floats (CTA-per-walker kernel)
    WS_CTR = 4096,          // +0: pivot-fallback counter (u64, cumulative); +8: float64-Slater-fallback counter; +64: queue head, tail, spin diagnostics (3 x u64);
                            // +128: 12 x u64 per-phase cycle sums (developer builds with CGBW_PROF only)
    WS_WCONST = 4352,       // staging copy of the constant-bank weight block (warp-per-walker kernel)
    WS_WSHARED = 4352 + 11264,   // float64 single-particle weights in the CTA-shared layout: source of the kernel's TMA bulk copy
    WS_DYN = 4352 + 11264 + 5120
};
inline unsigned ws_qcap(int B) { unsigned c = 64; while (c < 2u * (unsigned)(B > 0 ? B : 1)) c <<= 1; return c; }
inline size_t ws_round(size_t b) { return (b + 255) & ~size_t(255); }
inline size_t ws_total_bytes(int B) {
    return WS_DYN + ws_round(8 * (size_t)ws_qcap(B)) + ws_round(sizeof(int) * (size_t)B) + ws_round(sizeof(double) * (size_t)B) +
           ws_round(sizeof(int) * (size_t)B);
}

inline int sweep_num_warps(int n) { int w = (n + 7) / 8; return w < 2 ? 2 : w; }
size_t sweep_smem_bytes(int n);
int launch_sweep(const SweepParams& p, cudaStream_t st);
int launch_prep_params(const SweepParams& p, float* wpack, cudaStream_t st);      // + stats zero, key advance
// warp-per-walker sweep (n <= 32, instantiated shapes only): returns CGB_E_UNSUPPORTED when n has no instance
bool wsweep_supported(int n);
int launch_wsweep(const SweepParams& p, void* wconst_staging, cudaStream_t st);
int wsweep_spec_capacity(int n);

struct EwaldParams {
    const double* x; int B, n; double L, rs, kappa;
    const float4* items; int n_items, gmax;
    double* out;
    int wpw;                      // warps per walker (1 or 2), chosen by launch_ewald
    double* stats;                // CGB_STATS_LEN doubles or null: [2] sum V, [3] sum V^2, [4] B; [7] scratch counter
};
int launch_ewald(const EwaldParams& p, cudaStream_t st);
int launch_microbench(int kind, int blocks, int iters, float* out, cudaStream_t st);
int launch_rng_dump(uint64_t seed, const uint32_t* key_dev, int B, int S, int n, float* out_normal, double* out_uniform,
                    const uint32_t* kat_in, uint32_t* kat_out, int n_kat, cudaStream_t st);

// autoregressive momentum sampler (cgb_van.cu)
struct VanParams {
    const double* params; const double* ktable; int M, n, B;
    const int32_t* state_in;      // (B,n) or null: given choices (log_prob mode)
    const double* gumbel;         // (B,n,M) or null
    uint64_t seed;
    const uint32_t* key_dev;      // device-resident (2,) uint32 key (overrides seed), or null
    int32_t* out_state;           // (B,n) or null
    double* out_logprob;          // (B) or null
    double* out_logits;           // (B,n,M) masked logits, or null
};
size_t van_nparams(int M);
int launch_van(const VanParams& p, cudaStream_t st);

}  // namespace cgb
