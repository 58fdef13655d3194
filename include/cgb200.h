/*
 * cgb200.h — C ABI of libcgb200.so: B200 (sm_100a) kernels for the forward-sampling hot path of
 * fermiflow/CoulombGas.  Reference file:line citations are relative to the reference tree.
 *
 * The reference has no plugin / FFI interface: its hot path is three Python callables.
 * Each launcher below replaces one of them and is what a jax.ffi custom call binds:
 *
 *   cgb_logp_f64io         <->  logp(x, params, state_idx)                 src/logpsi.py:174-181
 *                               (= 2*Re logpsi: flow src/flow.py:39-55, Slater log-det src/slater.py:4-19,
 *                                + 0.5*slogdet(jacfwd(flow)) src/logpsi.py:22-31)
 *   cgb_mcmc_sweep_f64io   <->  mcmc(logp_fn, x_init, key, mc_steps, mc_stddev)   src/MCMC.py:6-40
 *                               and the wrap x -= L*floor(x/L)                     src/VMC.py:24
 *   cgb_ewald_f64io        <->  potential_energy(x, kappa, G, L, rs)       src/potential.py:36-77
 *   cgb_van_sample_f64io   <->  sampler(params_van, key, batch)            src/sampler.py:30-38   (next row N2:
 *   cgb_van_logprob_f64io  <->  log_prob(params_van, state_idx)           src/sampler.py:40-44    the caller before the sweep)
 *
 * Conventions (all launchers):
 *   - asynchronous on the caller's stream (`stream` is a cudaStream_t passed as void*), device pointers
 *     only, no allocation, no synchronisation (one call per stream at a time per workspace); the caller owns
 *     every buffer; workspace size comes from the *_workspace_bytes(B, n) query and grows with the batch
 *     (ready queue + per-walker scheduler state of the sweep kernel).  No environment variable is read on
 *     the launch path: kernel choice and debug switches are the per-call `flags` argument.
 *     The warp-per-walker sweep kernel (n = 9, 13, 21, 25, 29) reads its weights as constant-bank operands
 *     from a per-device __constant__ block that every call refreshes stream-ordered.  Calls on DIFFERENT
 *     streams of one device are made safe by the library itself: it keeps one event per device ("last
 *     kernel that read the block") and makes the refresh of the next call wait for it, i.e. such sweeps are
 *     serialised on the device (they would each fill the GPU anyway) and every call sees its own parameters.
 *     The one exception is a stream that is being CAPTURED into a CUDA graph: no cross-stream dependency can
 *     be added there, so graphs holding sweeps with different parameters must not be replayed concurrently
 *     on one device.  This per-device event (mutex-protected) is the only host-side state of the library.
 *   - I/O buffers are float64 (the reference runs with jax_enable_x64, main.py:2-3); internal arithmetic is
 *     FP32 with float64 range reduction / log-det accumulation / Metropolis test.
 *   - return 0 on success, CGB_E_* (<0) for bad arguments / unsupported shapes, or the positive
 *     cudaError_t of a failed launch.  Nothing throws.
 *   - supported model shape: dim = 2, depth = 2, spsize = tpsize = 16 (the shape of every BASELINE.json
 *     config), 2 <= n <= 64, k-table rows = integer + one common twist (src/orbitals.py:46-55).
 */
#ifndef CGB200_H
#define CGB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CGB_OK               0
#define CGB_E_BADARG        -1   /* null pointer, non-positive size, ... */
#define CGB_E_UNSUPPORTED   -2   /* n out of range, dim != 2, k-table not integer+twist, ... */
#define CGB_E_WORKSPACE     -3   /* workspace missing or too small */

#define CGB_NPARAMS        1074  /* flat flow parameter vector, see cgb_pack order below */
#define CGB_MAX_N            64
#define CGB_ABI_VERSION       2

/* per-call flags (logp / sweep launchers).  Bits 0-3: kernel choice, for A/B measurements and tests only -- results do not
 * depend on it (same noise/uniform streams, same decisions in the same order); 0 lets the library choose. */
#define CGB_KERNEL_AUTO       0u
#define CGB_KERNEL_CTA        1u   /* CTA-per-walker kernel, sequential chains */
#define CGB_KERNEL_CTA4       2u   /* ... clusters of 4 CTAs per walker (speculative) */
#define CGB_KERNEL_CTA8       3u   /* ... clusters of 8 */
#define CGB_KERNEL_WARP       4u   /* warp-per-walker kernel with the ready queue (n = 9, 13, 21, 25, 29) */
#define CGB_KERNEL_SPEC4      5u   /* ... speculative, 4 warps per walker */
#define CGB_KERNEL_SPEC8      6u   /* ... speculative, 8 warps per walker */
#define CGB_KERNEL_MASK      15u
#define CGB_FLAG_FORCE_PIVOT  0x100u   /* debug: Jacobian log-det by the partially pivoted fallback LU */
#define CGB_FLAG_ADVANCE_KEY  0x200u   /* sweep: key_dev <- key_dev + golden-ratio increment BEFORE it is used (see below) */
#define CGB_FLAG_FORCE_F64_SLATER 0x400u   /* debug: Slater log-det by the float64 LU even where the FP32 LU passes its monitor */
#define CGB_FLAG_FORCE_SEGMENTS  0x800u   /* debug: CTA-per-walker kernel cuts every chain into 2-step segments handed from CTA to CTA
                                            * (the load-balancing mode it picks itself for n > 32 and a few walkers per CTA slot) */
#define CGB_FLAG_NO_SEGMENTS    0x1000u   /* A/B timing: never cut chains into segments */

/* Statistics vector of one sampling step (device, CGB_STATS_LEN float64): what the reference's pmean sites reduce
 * (src/MCMC.py:38-39 accept rate; the V_mean / V2_mean entries of src/VMC.py:46-53).  One all-reduce (sum) of entries 0..4
 * over the ranks gives the global statistics.  The sweep launcher zeroes the vector and fills [0], [1]; the Ewald launcher
 * fills [2], [3], [4] (deterministically: one fixed-order reduction by the last CTA to finish); [7] is scratch and must be
 * zero when the Ewald launcher is called on a vector the sweep launcher has not initialised. */
#define CGB_STATS_LEN          8
#define CGB_STATS_NACCEPT      0   /* sum over walkers of accepted moves */
#define CGB_STATS_NPROPOSED    1   /* B * mc_steps */
#define CGB_STATS_SUM_V        2   /* sum_b V_b  (without Vconst) */
#define CGB_STATS_SUM_V2       3   /* sum_b V_b^2 */
#define CGB_STATS_NWALKERS     4   /* B */

/* Flat parameter layout (float64, CGB_NPARAMS entries), haiku module names of src/flow.py:
 *   Wt (5x16)  bt (16)   "fermi_net/~/linear_2"   two-particle layer   src/flow.py:13,46
 *   W0 (9x16)  b0 (16)   "fermi_net/~/linear"     single-particle #0   src/flow.py:11,45
 *   W1 (48x16) b1 (16)   "fermi_net/~/linear_1"   single-particle #1   src/flow.py:11,52
 *   Wf (16x2)  bf (2)    "fermi_net/linear"       final                src/flow.py:54
 * weights row-major (in, out) exactly as haiku stores `w`. */

int cgb_abi_version(void);
const char* cgb_error_string(int code);

/* ---- host helpers (pure CPU, no CUDA calls) ------------------------------------------------- */

/* Validate a k-table `sp_indices_twist` (M x 2, float64, host) = integer + common twist
 * (main.py:86-90).  Writes the twist.  Returns CGB_E_UNSUPPORTED otherwise. */
int cgb_ktable_twist_host(const double* ktable_host, int M, double twist_out[2]);

/* Fold the reference's G list (nG x 2 integers, src/potential.py:7-17) and kappa into the merged
 * reciprocal-space item table the Ewald kernel consumes: items_out[4*t + {0,1,2,3}] =
 * {gx >= 0, gy >= 0, weight of (gx,gy)&(-gx,-gy), weight of (gx,-gy)&(-gx,gy)} as float32 bit-compatible
 * floats, weights = g_G of src/potential.py:58 (2D).  items_out must hold 4*nG floats; *n_items <= nG.
 * Also returns max |G| component in *gmax (table size the kernel needs). */
int cgb_ewald_prepare_host(const int64_t* G_host, int nG, double kappa, float* items_out,
                           int* n_items, int* gmax);

/* ---- workspace queries ---------------------------------------------------------------------- */
int cgb_logp_workspace_bytes(int B, int n, size_t* bytes);
int cgb_mcmc_workspace_bytes(int B, int n, size_t* bytes);

/* ---- logp  (src/logpsi.py:174-181) ---------------------------------------------------------- */
/* x (B,n,2) f64; state_idx (B,n) int32 in [0,M); ktable (M,2) f64 device; twist from
 * cgb_ktable_twist_host; params (CGB_NPARAMS) f64 device; out_logp (B) f64.
 * Optional debug outputs (may be NULL): out_parts (B,2) f64 = {log|det D| (incl. -n log L), log|det J|};
 * out_jac (B,2n,2n) f32 = the analytic flow Jacobian. */
int cgb_logp_f64io(void* stream, const double* x, const int32_t* state_idx, const double* ktable, int M,
                   const double twist[2], const double* params, int B, int n, double L,
                   double* out_logp, double* out_parts, float* out_jac, void* workspace,
                   size_t workspace_bytes, unsigned flags);

/* ---- Metropolis sweep  (src/MCMC.py:22-40 + src/VMC.py:24) ---------------------------------- */
/* x_inout (B,n,2) f64 is advanced in place by mc_steps all-electron Gaussian moves (donated buffer,
 * src/VMC.py:11).  Parity mode: noise (mc_steps,B,n,2) f64 ~ N(0,1) and uniforms (mc_steps,B) f64 in
 * [0,1) are supplied by the caller (they replace jax.random.normal/uniform, MCMC.py:26,29).
 * Throughput mode: noise == uniforms == NULL and the kernel draws from Philox4x32-10.  The Philox key is
 * `seed`, or -- when key_dev != NULL -- the device-resident (2,) uint32 PRNG key the reference carries per
 * device inside pmap (src/VMC.py:20, src/MCMC.py:24; key[0] << 32 | key[1]), read by the kernel itself: no
 * host round trip, no recompilation per step.  With CGB_FLAG_ADVANCE_KEY the launcher first replaces the
 * device key by key + 0x9E3779B97F4A7C15 (mod 2^64) and the sweep uses the new value: a captured CUDA graph
 * of the step then draws fresh noise on every replay (the analogue of `key, subkey = split(key)`).
 * wrap != 0 applies x -= L*floor(x/L) at the end (VMC.py:24).
 * out_stats (CGB_STATS_LEN f64, device) or NULL: zeroed, then [0] = accepted moves, [1] = B*mc_steps.
 * Outputs: out_logp (B) f64 final log p per walker; out_naccept (B) int32 accepted moves per walker
 * (accept_rate = sum / (mc_steps*B), MCMC.py:38-39 — the cross-device mean is the caller's psum);
 * out_accept_trace (mc_steps,B) uint8 or NULL; out_logp_trace (mc_steps+1,B) f64 proposal log p
 * (row 0 = initial) or NULL.
 * The library picks the kernel by (n, B): ready-queue warp-per-walker kernel for batches that fill the
 * GPU; for small batches (the reference's per-device shards: 256 walkers at n=29, 64 at n=49/57) a
 * speculative mode that evaluates the proposals of the next 2-3 steps of every chain at once on the
 * otherwise idle SMs -- same noise/uniform streams, same decisions in the same order, so the result
 * does not depend on the choice.  `flags`: CGB_KERNEL_* override for A/B measurements. */
int cgb_mcmc_sweep_f64io(void* stream, double* x_inout, const int32_t* state_idx, const double* ktable,
                         int M, const double twist[2], const double* params, const double* noise,
                         const double* uniforms, uint64_t seed, uint32_t* key_dev, int mc_steps,
                         double mc_stddev, double L, int B, int n, int wrap, double* out_logp,
                         int32_t* out_naccept, uint8_t* out_accept_trace, double* out_logp_trace,
                         double* out_stats, void* workspace, size_t workspace_bytes, unsigned flags);

/* ---- Ewald energy  (src/potential.py:36-77) ------------------------------------------------- */
/* x (B,n,2) f64; items (n_items x 4 floats, device) and gmax from cgb_ewald_prepare_host;
 * out_V (B) f64 = 2*rs/L * psi(x/L)   (caller adds Vconst, src/VMC.py:40).
 * out_stats (CGB_STATS_LEN f64, device) or NULL: [2] = sum V, [3] = sum V^2, [4] = B (see CGB_STATS_*). */
int cgb_ewald_f64io(void* stream, const double* x, int B, int n, double L, double rs, double kappa,
                    const float* items, int n_items, int gmax, double* out_V, double* out_stats);

/* ---- autoregressive momentum sampler (SURVEY section 8 f/N2: the caller right before the sweep) -------------- */
/* Replaces `sampler(params_van, key, batch)` and `log_prob(params_van, state_idx)` of
 * make_autoregressive_sampler (src/sampler.py:30-44) for the causal Transformer of src/autoregressive.py:50-96
 * with the reference's default shape (main.py:24-27: nlayers 2, modelsize 16, nheads 4, nhidden 32; input = the
 * 2 components of the occupied k vector).  params: flat float64 device vector of cgb_van_nparams(M) = 18 M + 4368
 * entries in the order  x1hat[M] | embedding_mlp w[2][16] b[16] | per layer {query, key, value, linear: w[16][16]
 * b[16]; mlp/linear w[16][32] b[32]; mlp/linear_1 w[32][16] b[16]} | output_mlp w[16][M] b[M]   (haiku [in][out]).
 * ktable (M,2) f64 device = sp_indices_twist.  n <= 64, n <= M <= 256.
 * sample: gumbel (B,n,M) f64 = the Gumbel noise jax.random.categorical would add to the logits (parity mode), or
 *   NULL -> Philox4x32-10 keyed by (seed; walker, draw, orbital); key_dev != NULL: the key is the device-resident
 *   (2,) uint32 array instead of `seed` (src/VMC.py:20-22).  out_state_idx (B,n) int32 strictly increasing;
 *   out_logprob (B) f64 = log_prob of the drawn states, or NULL; out_logits (B,n,M) f64 masked logits, or NULL.
 * logprob: the same pass with the choices given (sampler.py:40-44). */
size_t cgb_van_nparams(int M);
int cgb_van_sample_f64io(void* stream, const double* params, size_t nparams, const double* ktable, int M, int n,
                         int B, const double* gumbel, uint64_t seed, const uint32_t* key_dev, int32_t* out_state_idx,
                         double* out_logprob, double* out_logits);
int cgb_van_logprob_f64io(void* stream, const double* params, size_t nparams, const double* ktable, int M, int n,
                          int B, const int32_t* state_idx, double* out_logprob, double* out_logits);

/* ---- roofline micro-benchmarks (bench.py only): FFMA / MUFU.EX2 saturating kernels ---------- */
/* Runs `iters` dependent-chain iterations on `blocks` x 256 threads; writes a checksum to out (1 float).
 * kind 0: 8 independent FFMA chains per thread (2*8*iters flop per thread);
 * kind 1: 8 independent ex2.approx chains (8*iters MUFU ops per thread);
 * kind 2: 8 independent DFMA chains (float64). */
int cgb_microbench(void* stream, int kind, int blocks, int iters, float* out);

/* ---- debug export of the throughput mode's random streams (tests only) ----------------------- */
/* out_normal (S,B,n,2) f32 / out_uniform (S,B) f64 (either may be NULL): exactly the proposal noise and Metropolis
 * uniforms the sweep kernels draw for (walker b, step s+1, particle i) under `seed` / `key_dev`.
 * kat_in (n_kat,6) u32 = Philox counter[4] + key[2], kat_out (n_kat,4) u32 = the raw Philox4x32-10 block (known-answer
 * vectors of Random123).  All device pointers. */
int cgb_debug_rng(void* stream, uint64_t seed, const uint32_t* key_dev, int B, int S, int n, float* out_normal,
                  double* out_uniform, const uint32_t* kat_in, uint32_t* kat_out, int n_kat);

#ifdef __cplusplus
}
#endif
#endif /* CGB200_H */
