// Autoregressive momentum sampler (VAN) of fermiflow/CoulombGas for sm_100a: the step right before the Metropolis
// sweep (src/VMC.py:22 -> src/sampler.py:30-44 -> src/autoregressive.py:15-96).  SURVEY.md section 8 f/N2.
//
// The reference draws the n occupied orbitals of a walker one at a time and re-evaluates the WHOLE causal Transformer
// on the padded state for every draw (n network passes of n positions each, dozens of small XLA launches per pass).
// Causality makes position t of the network a function of the first t+1 choices only, so here one warp per walker
// runs the n draws back to back and evaluates each position exactly once, appending its keys / values to a
// per-warp cache in shared memory (O(n) positions instead of O(n^2)).  Everything is float64 like the reference
// (main.py:2-3): the model is tiny (5 826 parameters at M = 81), so there is no precision engineering to do.
//
// Model shape (the reference's defaults, main.py:24-27, and every shipped run): 2 layers, model size 16, 4 heads of
// key size 4, MLP hidden size 32, input = the 2 components of the occupied k vector.  Per position, per layer
// (autoregressive.py:77-85): h += Linear(concat_heads(softmax(q k^T / 2, causal) v));  h += W2 tanh(W1 h + b1) + b2.
// Logits of draw i: x1hat for i = 0 (autoregressive.py:92-94), else output_mlp(tanh(h_{i-1})); masked so that indices are
// strictly increasing and leave room for the remaining draws (sampler.py:6-10,27-28: masked logits = -1e50).
// Draw: jax.random.categorical = argmax(logits + Gumbel) (first maximum); the Gumbel noise is an input in parity mode and
// comes from Philox4x32-10 keyed by (walker, draw, orbital) in throughput mode.  log_prob (sampler.py:40-44) is the
// same pass with the choices given.
#include "cgb_common.cuh"
#include "cgb_internal.h"

namespace cgb {

namespace van {
constexpr int D = 16, H = 4, KS = 4, HID = 32, NL = 2;
constexpr int KVP = 33;                                // KV-cache row pitch (doubles): odd, so the per-position score loads of 32 lanes spread over all banks
// flat parameter layout (doubles), M = number of orbitals:  x1hat[M] | emb w[2][16] b[16] |
//   per layer: q w[16][16] b[16] | k .. | v .. | o .. | mlp1 w[16][32] b[32] | mlp2 w[32][16] b[16] |  out w[16][M] b[M]
constexpr int LIN = D * D + D;                         // 272
constexpr int LSZ = 4 * LIN + (D * HID + HID) + (HID * D + D);   // 2160
__host__ __device__ inline int o_emb(int M) { return M; }
__host__ __device__ inline int o_layer(int M, int l) { return M + 48 + l * LSZ; }
__host__ __device__ inline int o_out(int M) { return M + 48 + NL * LSZ; }
__host__ __device__ inline size_t nparams(int M) { return (size_t)18 * M + 4368; }
}  // namespace van

// exp(x) for x <= 0 and tanh(x) in float64 without the library's special-case handling: range reduction to |f| <= ln2/2 and
// a degree-12 Taylor polynomial (relative error < 1e-14, far inside the 1e-10 parity budget of the log-probabilities);
// tanh = (1 - e)/(1 + e), e = exp(-2|x|), with a Newton-refined FP32 reciprocal.
__device__ __forceinline__ double exp_nonpos(double x) {
    const double y = fmax(x * 1.4426950408889634074, -1000.0);
    const double kf = rint(y);
    const double f = fma(kf, -0.69314718055994530942, x < -700.0 ? -700.0 : x);     // x - kf ln2 (kf ln2 exact to 1e-17 |kf|)
    double p = 2.08767569878680989792e-09;                                          // 1/12!
    p = fma(p, f, 2.50521083854417187751e-08);
    p = fma(p, f, 2.75573192239858906526e-07);
    p = fma(p, f, 2.75573192239858906526e-06);
    p = fma(p, f, 2.48015873015873015873e-05);
    p = fma(p, f, 1.98412698412698412698e-04);
    p = fma(p, f, 1.38888888888888888889e-03);
    p = fma(p, f, 8.33333333333333333333e-03);
    p = fma(p, f, 4.16666666666666666667e-02);
    p = fma(p, f, 1.66666666666666666667e-01);
    p = fma(p, f, 0.5);
    p = fma(p, f, 1.0);
    p = fma(p, f, 1.0);
    return p * __hiloint2double(((int)kf + 1023) << 20, 0);
}
__device__ __forceinline__ double tanh_fast(double x) {
    const double e = exp_nonpos(-2.0 * fabs(x));
    const double d = 1.0 + e;
    double r = (double)rcp_approx((float)d);
    r = r * (2.0 - d * r);
    r = r * (2.0 - d * r);
    return copysign((1.0 - e) * r, x);
}
// max over the warp of a float64 value, taken in FP32 (rounded up): only used as the softmax shift, where any common
// value within a few ulp(FP32) of the maximum is as good as the maximum itself
__device__ __forceinline__ double warp_shift_f64(double v) {
    float f = __double2float_ru(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) f = fmaxf(f, __shfl_xor_sync(0xffffffffu, f, o));
    return (double)f;
}
__device__ __forceinline__ double warp_max_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_f64v(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256) van_kernel(const VanParams p) {
    using namespace van;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int M = p.M, n = p.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    double* const sP = reinterpret_cast<double*>(smem_raw);                     // parameters, CTA-shared
    const int np = (int)nparams(M);
    for (int t = threadIdx.x; t < np; t += blockDim.x) sP[t] = p.params[t];
    __syncthreads();
    const int np_pad = (np + 1) & ~1;
    // per-warp slice: KV cache [NL][n][KVP] (keys 0..15, values 16..31), scratch x[16] q[16] att[16] hid[32] y[16], wts[H][64]
    const int slice = NL * n * KVP + 96 + H * 64;
    double* const sKV = sP + np_pad + (size_t)warp * slice;
    double* const sx = sKV + NL * n * KVP;
    double* const sq = sx + 16;
    double* const satt = sq + 16;
    double* const shid = satt + 16;
    double* const sy = shid + 32;
    double* const swt = sy + 16;
    const double* W_emb = sP + o_emb(M);
    const double* W_out = sP + o_out(M);
    const int c = lane & 15;

    for (int b = blockIdx.x * nwarps + warp; b < p.B; b += gridDim.x * nwarps) {
        int prev = -1;
        double logp = 0.0;
        for (int i = 0; i < n; ++i) {
            // ---------------- network position i-1 (autoregressive.py:72-90), incremental ----------------
            if (i > 0) {
                const int t = i - 1;
                const double kx = __ldg(p.ktable + 2 * prev), ky = __ldg(p.ktable + 2 * prev + 1);    // sp_indices[state_idx]
                __syncwarp();
                if (lane < 16) sx[c] = tanh_fast(fma(kx, W_emb[c], fma(ky, W_emb[16 + c], W_emb[32 + c])));
                __syncwarp();
#pragma unroll 1
                for (int l = 0; l < NL; ++l) {
                    const double* Wl = sP + o_layer(M, l);
                    double* kv = sKV + ((size_t)l * n + t) * KVP;
                    {   // lanes 0..15: query then value; lanes 16..31: key
                        const double* Wa = Wl + (lane < 16 ? 0 : LIN);
                        double a = Wa[D * D + c], v = Wl[2 * LIN + D * D + c];
#pragma unroll
                        for (int e = 0; e < D; ++e) {
                            const double xe = sx[e];
                            a = fma(xe, Wa[e * D + c], a);
                            v = fma(xe, Wl[2 * LIN + e * D + c], v);
                        }
                        if (lane < 16) { sq[c] = a; kv[16 + c] = v; } else kv[c] = a;
                    }
                    __syncwarp();
                    // attention weights per head over positions tau <= t   (hk.MultiHeadAttention: q.k/sqrt(key_size), softmax)
#pragma unroll
                    for (int h = 0; h < H; ++h) {
                        double s0 = -INFINITY, s1 = -INFINITY;
                        if (lane <= t) {
                            const double* kk = sKV + ((size_t)l * n + lane) * KVP + KS * h;
                            s0 = 0.5 * (sq[KS * h] * kk[0] + sq[KS * h + 1] * kk[1] + sq[KS * h + 2] * kk[2] + sq[KS * h + 3] * kk[3]);
                        }
                        if (lane + 32 <= t) {
                            const double* kk = sKV + ((size_t)l * n + lane + 32) * KVP + KS * h;
                            s1 = 0.5 * (sq[KS * h] * kk[0] + sq[KS * h + 1] * kk[1] + sq[KS * h + 2] * kk[2] + sq[KS * h + 3] * kk[3]);
                        }
                        const double mx = warp_shift_f64(fmax(s0, s1));
                        const double e0 = lane <= t ? exp_nonpos(s0 - mx) : 0.0, e1 = lane + 32 <= t ? exp_nonpos(s1 - mx) : 0.0;
                        const double inv = 1.0 / warp_sum_f64v(e0 + e1);
                        swt[h * 64 + lane] = e0 * inv;
                        swt[h * 64 + 32 + lane] = e1 * inv;
                    }
                    __syncwarp();
                    if (lane < 16) {
                        double a = 0.0;
                        const double* wv = swt + (c >> 2) * 64;
                        for (int tau = 0; tau <= t; ++tau) a = fma(wv[tau], sKV[((size_t)l * n + tau) * KVP + 16 + c], a);
                        satt[c] = a;
                    }
                    __syncwarp();
                    if (lane < 16) {                                  // output projection + residual (autoregressive.py:83)
                        const double* Wo = Wl + 3 * LIN;
                        double a = Wo[D * D + c];
#pragma unroll
                        for (int e = 0; e < D; ++e) a = fma(satt[e], Wo[e * D + c], a);
                        sx[c] += a;
                    }
                    __syncwarp();
                    {                                                 // DenseBlock (autoregressive.py:43-48): 16 -> 32 tanh -> 16
                        const double* W1 = Wl + 4 * LIN;
                        double a = W1[D * HID + lane];
#pragma unroll
                        for (int e = 0; e < D; ++e) a = fma(sx[e], W1[e * HID + lane], a);
                        shid[lane] = tanh_fast(a);
                    }
                    __syncwarp();
                    if (lane < 16) {
                        const double* W2 = Wl + 4 * LIN + D * HID + HID;
                        double a = W2[HID * D + c];
#pragma unroll
                        for (int j = 0; j < HID; ++j) a = fma(shid[j], W2[j * D + c], a);
                        sx[c] += a;                                   // residual (autoregressive.py:85)
                    }
                    __syncwarp();
                }
                if (lane < 16) sy[c] = tanh_fast(sx[c]);                   // autoregressive.py:87
                __syncwarp();
            }
            // ---------------- logits of draw i, mask, draw / given choice, log-softmax ----------------
            const int hi = M - n + i;                                 // sampler.py:7: tril(ones(n, M), k = M - n)
            double best = -INFINITY, lmax = -INFINITY;
            int bestj = 0x7fffffff;
            double lg[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int j = lane + 32 * r;
                lg[r] = -INFINITY;
                if (j < M) {
                    double a;
                    if (i == 0) a = sP[j];                            // x1hat (autoregressive.py:92-94)
                    else {
                        a = W_out[D * M + j];
#pragma unroll
                        for (int e = 0; e < D; ++e) a = fma(sy[e], W_out[e * M + j], a);
                    }
                    const bool ok = j > prev && j <= hi;              // sampler.py:8-10
                    a = ok ? a : -1e50;                               // sampler.py:28
                    lg[r] = a;
                    lmax = fmax(lmax, a);
                    if (p.out_logits) p.out_logits[((size_t)b * n + i) * M + j] = a;
                    if (!p.state_in) {
                        double g;
                        if (p.gumbel) g = __ldg(p.gumbel + ((size_t)b * n + i) * M + j);
                        else {
                            g = rng_gumbel(rng_seed_of(p.seed, p.key_dev), (uint32_t)b, (uint32_t)i, (uint32_t)j);
                        }
                        const double v = a + g;
                        if (ok && v > best) { best = v; bestj = j; } // first maximum inside the lane (j ascending); masked
                                                                     // entries (-1e50 + g) can never win in the reference either
                    }
                }
            }
            int idx;
            if (p.state_in) {
                idx = __ldg(p.state_in + (size_t)b * n + i);
                idx = min(max(idx, 0), M - 1);
            } else {
                const double gbest = warp_max_f64(best);
                idx = __reduce_min_sync(0xffffffffu, best == gbest ? bestj : 0x7fffffff);   // argmax takes the first maximum
            }
            // log-softmax of the masked logits at idx  (sampler.py:41-43)
            lmax = warp_shift_f64(lmax);
            double se = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (lane + 32 * r < M) se += exp_nonpos(lg[r] - lmax);
            se = warp_sum_f64v(se);
            const int own = idx >> 5;
            double li = 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r)
                if (r == own) li = lg[r];
            li = __shfl_sync(0xffffffffu, li, idx & 31);
            logp += li - (lmax + log(se));
            if (p.out_state && lane == 0) p.out_state[(size_t)b * n + i] = idx;
            prev = idx;
        }
        if (p.out_logprob && lane == 0) p.out_logprob[b] = logp;
    }
}

size_t van_nparams(int M) { return van::nparams(M); }

int launch_van(const VanParams& p, cudaStream_t st) {
    using namespace van;
    if (p.M < p.n || p.M > 256 || p.n < 1 || p.n > 64) return CGB_E_UNSUPPORTED;
    const size_t np_pad = (nparams(p.M) + 1) & ~size_t(1);
    const size_t slice = (size_t)NL * p.n * KVP + 96 + H * 64;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int wpc = 8;
    while (wpc > 1 && 8 * (np_pad + wpc * slice) > 200 * 1024) wpc >>= 1;
    while (wpc > 1 && (long long)p.B < (long long)sms * wpc) wpc >>= 1;
    const size_t smem = 8 * (np_pad + wpc * slice);
    if (smem > 227 * 1024) return CGB_E_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(van_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int occ = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, van_kernel, wpc * 32, smem);
    if (occ < 1) occ = 1;
    const long long need = ((long long)p.B + wpc - 1) / wpc;
    const int grid = (int)(need < (long long)sms * occ ? need : (long long)sms * occ);
    van_kernel<<<grid, wpc * 32, smem, st>>>(p);
    return (int)cudaGetLastError();
}

}  // namespace cgb
