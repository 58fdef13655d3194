// Ewald Coulomb energy of a walker batch for sm_100a  (src/potential.py:36-77).
//
// The reference evaluates  sum_G g_G sum_{i<j} cos(2 pi G.r_ij)  pairwise (O(|G| n^2)); here the same sum
// is taken in structure-factor form  1/2 (|S(G)|^2 - n),  S(G) = sum_i e^{2 pi i G.u_i}  (O(|G| n)), with
// e^{2 pi i G.u} = px_i[|Gx|] * py_i[|Gy|]^(*) from per-particle power tables, and the (Gx,Gy)/(Gx,-Gy)/(-G)
// quadruple sharing one set of four real accumulators (host-side merge in cgb_ewald_prepare_host).
// Real space: nearest image only, erfc(kappa d)/d (potential.py:45-50).
//
// One WARP per walker (persistent over the batch), no block barrier once the weight table is staged:
//   * lane i builds the power tables of particles i, i+32: ONE sincospi per coordinate (phase reduced in float64,
//     hi/lo corrected), higher harmonics by complex multiplication -- both coordinates and the even / odd harmonics as four
//     independent chains;
//   * the reciprocal sum is a small complex matrix product  S+-(gx,gy) = sum_i px_i[gx] py_i[gy]^(*)  over the
//     (gmax+1)^2 quadrant, register-tiled 2 x 4 per lane: per particle 3 broadcast LDS.128 feed 16 packed FMAs (FFMA2:
//     (ac, ad) += xr (yr, yi), (bc, bd) += xi (yr, yi)); the merged weights (w+, w-) sit in a dense CTA-shared table (zero
//     outside the |G| <= Gmax disc);
//   * the n(n-1)/2 real-space pairs are enumerated as a circulant (particle i meets i+d mod n, d = 1..n/2): no index decode,
//     differences from a hi/lo float split of the reduced coordinates, a 15-instruction erfc (erfc_pos below);
//   * half of every scheduler's warps run the two loops in the other order (one is FMA-pipe-bound, the other latency-bound);
//   * batches that cannot fill the GPU with one warp per walker give every walker TWO warps: the pair shares the walker's
//     tables, splits the particles of the table build, the tiles and the pairs, and meets at a named barrier (64 threads);
//   * a batch of one to three waves gets CTAs of equal size on every SM (launch_ewald).
#include "cgb_common.cuh"
#include "cgb_internal.h"

namespace cgb {

struct EwLay {                      // per-warp shared-memory slice (bytes), n particles, harmonics 0..gmax
    int npx, npy, pitch_x, pitch_y, o_px, o_py, total;
    __host__ __device__ EwLay(int n, int gmax) {
        npx = (gmax + 2) & ~1;                 // px harmonics padded to a multiple of 2 (tile width in gx)
        npy = (gmax + 4) & ~3;                 // py harmonics padded to a multiple of 4 (tile width in gy)
        pitch_x = npx + 2;                     // float2 units; +2 keeps 16-byte alignment and spreads rows over the banks
        pitch_y = npy + 2;
        o_px = 16 * n;                         // float4 u[n] (hi/lo reduced coordinates) first
        o_py = o_px + 8 * n * pitch_x;
        total = (o_py + 8 * n * pitch_y + 15) & ~15;
    }
};

__device__ __forceinline__ float2 cmulf(const float2 a, const float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// erfc(x) for x >= 0: t exp(-x^2 + P(t)), t = 1 / (1 + x/2), P of degree 9 -- the classic Chebyshev-fitted form (fractional error
// < 1.2e-7 in exact arithmetic; evaluated in FP32: < 1.6e-6 relative for x < 3, 3.7e-7 absolute anywhere; tools/gpu_diag.py and
// the Ewald parity tests bound the effect on V: < 1e-6 of the energy scale even for random walkers with pairs 0.005 of a cell apart).
// 15 instructions against ~45 of erfcf, which was 19 % of the kernel.
__device__ __forceinline__ float erfc_pos(const float x) {
    const float t = rcp_approx(fmaf(0.5f, x, 1.f));
    float p = 0.17087277f;
    p = fmaf(p, t, -0.82215223f); p = fmaf(p, t, 1.48851587f); p = fmaf(p, t, -1.13520398f); p = fmaf(p, t, 0.27886807f);
    p = fmaf(p, t, -0.18628806f); p = fmaf(p, t, 0.09678418f); p = fmaf(p, t, 0.37409196f); p = fmaf(p, t, 1.00002368f);
    p = fmaf(p, t, -1.26551223f);
    return t * ex2_approx((p - x * x) * 1.4426950408889634f);
}

#ifndef CGBE_WPW2_MAX_BATCH_PER_SM
#define CGBE_WPW2_MAX_BATCH_PER_SM 8       // two warps per walker while B <= this x #SMs (16 for n > 32: two particles per lane)
#endif
#ifndef CGBE_MAX_THREADS
#define CGBE_MAX_THREADS 256
#endif

__global__ void __launch_bounds__(CGBE_MAX_THREADS) ewald_kernel(const EwaldParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ double s_part[8];                 // one per walker of the CTA (at most 8)
    const int n = p.n, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int wpw = p.wpw;                                   // warps per walker: 1 or 2
    const int grp = wpw == 2 ? warp >> 1 : warp, sub = wpw == 2 ? warp & 1 : 0, ngrp = nwarps / wpw;
    const int glane = lane + 32 * sub, gsize = 32 * wpw;
    auto group_sync = [&]() {
        if (wpw == 2) asm volatile("bar.sync %0, 64;" ::"r"(1 + grp) : "memory");
        else __syncwarp();
    };
    const EwLay L_(n, p.gmax);
    const int tiles_x = L_.npx >> 1, tiles_y = L_.npy >> 2, ntiles = tiles_x * tiles_y;
    // the first walker's coordinates are requested before the weight table is staged: the two global-memory latencies overlap
    double2 xpre = make_double2(0.0, 0.0);
    {
        const int b = blockIdx.x * ngrp + grp;
        if (b < p.B && glane < n) xpre = __ldg(reinterpret_cast<const double2*>(p.x) + (size_t)b * n + glane);
    }
    // dense weights (w+, w-)[gx][gy] over the padded quadrant, CTA-shared
    float2* const Wd = reinterpret_cast<float2*>(smem_raw);
    const int wd_bytes = (8 * L_.npx * L_.npy + 15) & ~15;
    // the table is zeroed now and filled after the first walker's power tables are built (nothing reads it before the first tile
    // epilogue): the items' global-memory latency hides behind the table build
    float4 itpre = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((int)threadIdx.x < p.n_items) itpre = __ldg(p.items + threadIdx.x);
    for (int t = threadIdx.x; t < L_.npx * L_.npy; t += blockDim.x) Wd[t] = make_float2(0.f, 0.f);
    unsigned char* const sm = smem_raw + wd_bytes + grp * L_.total;
    float4* const uf = reinterpret_cast<float4*>(sm);        // reduced coordinates split hi / lo: (hi_x, hi_y, lo_x, lo_y)
    float2* const px = reinterpret_cast<float2*>(sm + L_.o_px);
    float2* const py = reinterpret_cast<float2*>(sm + L_.o_py);
    const double invL = 1.0 / p.L;
    const float kap = (float)p.kappa, fn = (float)n;

    for (int b0 = blockIdx.x * ngrp; b0 < p.B; b0 += gridDim.x * ngrp) {      // the same trip count for every warp of the CTA
        const int b = b0 + grp;
        const bool live = b < p.B;
        group_sync();
        // ---- reduced coordinates and power tables e^{2 pi i g u}, g = 0..gmax (padding harmonics = 0)
        for (int i = glane; live && i < n; i += gsize) {
            double2 xv = (b0 == blockIdx.x * ngrp && i == glane) ? xpre
                                                                 : __ldg(reinterpret_cast<const double2*>(p.x) + (size_t)b * n + i);
            xv.x *= invL; xv.y *= invL;                                   // x / L  (potential.py:77)
            xv.x -= floor(xv.x); xv.y -= floor(xv.y);
            const float uhx = (float)xv.x, uhy = (float)xv.y;
            const float ulx = (float)(xv.x - (double)uhx), uly = (float)(xv.y - (double)uhy);
            uf[i] = make_float4(uhx, uhy, ulx, uly);
            // both coordinates at once, even and odd harmonics as two independent chains (each multiplies by p1^2): four
            // independent complex products per step instead of a serial chain of 2 (gmax + 1) per coordinate
            float2 p1[2], p2[2], ev[2], od[2];
#pragma unroll
            for (int comp = 0; comp < 2; ++comp) {
                const float uh = comp ? uhy : uhx, ul = comp ? uly : ulx;
                float sn, cs;
                sincospif(2.f * uh, &sn, &cs);
                const float pl = 2.f * kPi * ul;                          // first-order hi/lo correction of the phase
                p1[comp] = make_float2(fmaf(-pl, sn, cs), fmaf(pl, cs, sn));
                p2[comp] = cmulf(p1[comp], p1[comp]);
                ev[comp] = make_float2(1.f, 0.f);
                od[comp] = p1[comp];
            }
            float2* const rowx = px + i * L_.pitch_x;
            float2* const rowy = py + i * L_.pitch_y;
            for (int g = 0; g < L_.npy; g += 2) {                          // npy >= npx
                const bool in0 = g <= p.gmax, in1 = g + 1 <= p.gmax;
                if (g < L_.npx)
                    *reinterpret_cast<float4*>(rowx + g) = make_float4(in0 ? ev[0].x : 0.f, in0 ? ev[0].y : 0.f, in1 ? od[0].x : 0.f, in1 ? od[0].y : 0.f);
                *reinterpret_cast<float4*>(rowy + g) = make_float4(in0 ? ev[1].x : 0.f, in0 ? ev[1].y : 0.f, in1 ? od[1].x : 0.f, in1 ? od[1].y : 0.f);
                ev[0] = cmulf(ev[0], p2[0]); od[0] = cmulf(od[0], p2[0]);
                ev[1] = cmulf(ev[1], p2[1]); od[1] = cmulf(od[1], p2[1]);
            }
        }
        if (b0 == blockIdx.x * ngrp) {                   // first walker of the CTA: stage the weights (CTA-uniform branch)
            __syncthreads();
            if ((int)threadIdx.x < p.n_items) Wd[(int)itpre.x * L_.npy + (int)itpre.y] = make_float2(0.5f * itpre.z, 0.5f * itpre.w);
            for (int t = threadIdx.x + blockDim.x; t < p.n_items; t += blockDim.x) {
                const float4 it = __ldg(p.items + t);
                Wd[(int)it.x * L_.npy + (int)it.y] = make_float2(0.5f * it.z, 0.5f * it.w);
            }
            __syncthreads();
        } else group_sync();
        float acc = 0.f, accr = 0.f;          // two accumulators added in a fixed order: the result does not depend on the warp
        // ---- reciprocal space: 2 x 4 register tile of (gx, gy) per lane
        auto reciprocal_space = [&]() {
        for (int tile = glane; live && tile < ntiles; tile += gsize) {
            const int tx = tile / tiles_y, ty = tile - tx * tiles_y;
            const float4* pxa = reinterpret_cast<const float4*>(px + 2 * tx);
            const float4* pya = reinterpret_cast<const float4*>(py + 4 * ty);
            // (ac, ad) += xr (yr, yi) and (bc, bd) += xi (yr, yi): the (re, im) pairs of py are register pairs as loaded, so every
            // particle costs 16 packed FMAs (FFMA2, scalar operand broadcast) instead of 32 scalar ones
            float2 A[2][4], Bq[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) { A[a][c] = make_float2(0.f, 0.f); Bq[a][c] = make_float2(0.f, 0.f); }
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
                const float4 xa = pxa[i * (L_.pitch_x >> 1)];
                const float4 y0 = pya[i * (L_.pitch_y >> 1)], y1 = pya[i * (L_.pitch_y >> 1) + 1];
                const float xr[2] = {xa.x, xa.z}, xi[2] = {xa.y, xa.w};
                const float2 Y[4] = {make_float2(y0.x, y0.y), make_float2(y0.z, y0.w), make_float2(y1.x, y1.y), make_float2(y1.z, y1.w)};
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        A[a][c] = __ffma2_rn(make_float2(xr[a], xr[a]), Y[c], A[a][c]);
                        Bq[a][c] = __ffma2_rn(make_float2(xi[a], xi[a]), Y[c], Bq[a][c]);
                    }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float2 w = Wd[(2 * tx + a) * L_.npy + 4 * ty + c];
                    const float spr = A[a][c].x - Bq[a][c].y, spi = A[a][c].y + Bq[a][c].x;      // S+ = sum px py
                    const float smr = A[a][c].x + Bq[a][c].y, smi = Bq[a][c].x - A[a][c].y;      // S- = sum px conj(py)
                    acc = fmaf(w.x, fmaf(spr, spr, spi * spi) - fn, acc);
                    acc = fmaf(w.y, fmaf(smr, smr, smi * smi) - fn, acc);
                }
        }
        };
        // ---- real space, nearest image (potential.py:45-50).  The n(n-1)/2 pairs are enumerated as a circulant: particle i
        // meets i+d (mod n), d = 1..n/2 (for even n the last offset only from the lower half) -- no index decode, the own
        // coordinates stay in registers, the partners of a warp are consecutive shared-memory words.  Differences from the
        // hi/lo split (error < 1e-8 of a cell), nearest image by the round-to-nearest magic constant.
        auto real_space = [&]() {
            const int half = n >> 1;
            const bool wide = n > 32;                     // two particles per lane (one per thread when the walker has two warps)
            const int d0 = wide ? 1 : 1 + sub, dstep = wide ? 1 : wpw;
            for (int i = wide ? glane : lane; live && i < n; i += wide ? gsize : 32) {
                const float4 ui = uf[i];
                const int dmax = (2 * half == n && i >= half) ? half - 1 : half;
#pragma unroll 2
                for (int d = d0; d <= dmax; d += dstep) {
                    int j = i + d;
                    j = j >= n ? j - n : j;
                    const float4 uj = uf[j];
                    float fx = (ui.x - uj.x) + (ui.z - uj.z), fy = (ui.y - uj.y) + (ui.w - uj.w);
                    fx -= (fx + 12582912.f) - 12582912.f;                      // rint for |f| < 2^22
                    fy -= (fy + 12582912.f) - 12582912.f;
                    const float d2 = fmaf(fx, fx, fy * fy);
                    const float rd = rsqrt_approx(d2);                         // d2 >= 1e-12 for distinct particles: no denormals
                    accr = fmaf(erfc_pos(kap * (d2 * rd)), rd, accr);
                }
            }
        };
        // the tile loop is bound by the FMA pipe, the pair loop by its dependency chain (rsqrt -> rcp -> Horner -> ex2): half of the
        // warps of every scheduler (warp = 4 k + scheduler) take them in the other order, so the two phases overlap on the SM
        if ((warp >> 2) & 1) { real_space(); reciprocal_space(); }
        else { reciprocal_space(); real_space(); }
        double a = (double)(acc + accr);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if (wpw == 2) {                                   // the second warp hands its partial sum to the first (fixed order)
            if (sub == 1 && lane == 0) s_part[grp] = a;
            group_sync();
            if (sub == 0) a += s_part[grp];
        }
        if (lane == 0 && sub == 0 && live) {
            const double g0 = -2.0 * sqrt(3.14159265358979323846) / p.kappa;       // potential.py:59
            a += g0 * (0.5 * n * (n - 1));                                        // potential.py:62
            p.out[b] = 2.0 * p.rs * invL * a;                                     // potential.py:77
        }
    }
    // ---- statistics of the batch (the V_mean / V2_mean entries of the pmean at src/VMC.py:46-53): the LAST CTA to finish
    // reduces out[0..B) in a fixed order -> deterministic sums, no second kernel, no floating-point atomics
    if (p.stats) {
        __shared__ bool s_last;
        __shared__ double s_red[2][CGBE_MAX_THREADS / 32];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned long long t = atomicAdd(reinterpret_cast<unsigned long long*>(p.stats + 7), 1ull);
            s_last = t == (unsigned long long)gridDim.x - 1ull;
        }
        __syncthreads();
        if (s_last) {
            __threadfence();
            double s1 = 0.0, s2 = 0.0;
            for (int b = threadIdx.x; b < p.B; b += blockDim.x) { const double v = __ldcg(p.out + b); s1 += v; s2 = fma(v, v, s2); }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            if (lane == 0) { s_red[0][warp] = s1; s_red[1][warp] = s2; }
            __syncthreads();
            if (threadIdx.x == 0) {
                double t1 = 0.0, t2 = 0.0;
                for (int w = 0; w < nwarps; ++w) { t1 += s_red[0][w]; t2 += s_red[1][w]; }
                p.stats[CGB_STATS_SUM_V] = t1; p.stats[CGB_STATS_SUM_V2] = t2; p.stats[CGB_STATS_NWALKERS] = (double)p.B;
                *reinterpret_cast<unsigned long long*>(p.stats + 7) = 0ull;
            }
        }
    }
}

int launch_ewald(const EwaldParams& p, cudaStream_t st) {
    const EwLay L_(p.n, p.gmax);
    const int wd_bytes = (8 * L_.npx * L_.npy + 15) & ~15;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // two warps per walker while one warp per walker cannot fill the machine (see the header of this file)
    EwaldParams q = p;
    // measured (profiles/r02_time_ewald.txt): n=29: B=256 5.1 vs 6.9 us, B=2048 11.9 vs 8.3 us (28 warps per SM in two 448-thread CTAs at
    // 72 registers: the half-warps' shorter chains do not pay for the second barrier-coupled warp); n=57: B=512 8.3 vs 12.5 us
    q.wpw = (long long)p.B <= (long long)(p.n > 32 ? 16 : CGBE_WPW2_MAX_BATCH_PER_SM) * sms ? 2 : 1;
    // walkers per CTA: 8 (4 with two warps each) while three CTAs still fit one SM's shared memory, fewer for large n / gmax;
    // small batches use smaller CTAs so that the walkers spread over all SMs
    int gpc = q.wpw == 2 ? CGBE_MAX_THREADS / 64 : 8;
    while (gpc > 1 && (size_t)wd_bytes + (size_t)gpc * L_.total > 72 * 1024) --gpc;
    if ((long long)p.B < (long long)sms * gpc) gpc = p.B / sms > 1 ? p.B / sms : 1;
    // a batch of a few waves: even out the warps per SM (2048 walkers on 148 SMs are 13.8 per SM: two CTAs of 7 on every SM
    // instead of two CTAs of 8 on 108 SMs and one on 40)
    if ((long long)p.B >= (long long)sms * gpc && (long long)p.B <= 3ll * sms * gpc) {
        const int per_sm = (int)(((long long)p.B + sms - 1) / sms), ctas = (per_sm + gpc - 1) / gpc;
        gpc = (per_sm + ctas - 1) / ctas;
    }
    const int wpc = gpc * q.wpw;
    const size_t smem = (size_t)wd_bytes + (size_t)gpc * L_.total;
    if (smem > 227 * 1024) return CGB_E_UNSUPPORTED;
    // attribute + occupancy queries are cached per (device, shape): the steady-state call is one launch
    static thread_local struct { int dev; size_t smem; int wpc; int occ; } cache = {-1, 0, 0, 0};
    if (cache.dev != dev || cache.smem != smem || cache.wpc != wpc) {
        cudaError_t e = cudaFuncSetAttribute(ewald_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ewald_kernel, wpc * 32, smem);
        if (occ < 1) occ = 1;
        cache = {dev, smem, wpc, occ};
    }
    occ = cache.occ;
    const long long need = ((long long)p.B + gpc - 1) / gpc;
    const int grid = (int)(need < (long long)sms * occ ? need : (long long)sms * occ);
    ewald_kernel<<<grid, wpc * 32, smem, st>>>(q);
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------
// roofline denominators: FFMA / MUFU.EX2 / DFMA saturating micro-kernels (bench.py only)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) microbench_kernel(int kind, int iters, float* out) {
    const float a = 1.0000001f, bcoef = 1e-9f;
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = 0.001f * (threadIdx.x + k);
    if (kind == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = fmaf(x[k], a, bcoef);
        }
    } else if (kind == 1) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) x[k] = ex2_approx(-fabsf(x[k]));
        }
    } else {
        double y[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = x[k];
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 8; ++k) y[k] = fma(y[k], 1.0000001, 1e-9);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = (float)y[k];
    }
    float sacc = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) sacc += x[k];
    if (sacc == 123.456f) out[0] = sacc;      // never true; defeats dead-code elimination
}

int launch_microbench(int kind, int blocks, int iters, float* out, cudaStream_t st) {
    microbench_kernel<<<blocks, 256, 0, st>>>(kind, iters, out);
    return (int)cudaGetLastError();
}

// ---- debug export of the throughput mode's random streams (tests: Philox known answers, moments / KS of the draws) -------
// out_normal (S,B,n,2) f32 and out_uniform (S,B) f64: exactly what the sweep kernels draw for (walker b, step s + 1, particle i);
// kat_in / kat_out (n_kat x 6 / x 4 words): raw Philox4x32-10 blocks for given (counter[4], key[2]).
__global__ void rng_dump_kernel(uint64_t seed_host, const uint32_t* key_dev, int B, int S, int n, float* out_normal,
                                double* out_uniform, const uint32_t* kat_in, uint32_t* kat_out, int n_kat) {
    const uint64_t seed = rng_seed_of(seed_host, key_dev);
    const size_t total = (size_t)S * B * n;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(t % n);
        const size_t sb = t / n;
        const int b = (int)(sb % B), s = (int)(sb / B);
        if (out_normal) {
            const float2 g = rng_normal2(seed, (uint32_t)b, (uint32_t)(s + 1), (uint32_t)i);
            out_normal[2 * t] = g.x; out_normal[2 * t + 1] = g.y;
        }
        if (out_uniform && i == 0) out_uniform[sb] = rng_uniform(seed, (uint32_t)b, (uint32_t)(s + 1));
    }
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n_kat; t += gridDim.x * blockDim.x) {
        const uint4 r = philox4x32_10(make_uint4(kat_in[6 * t], kat_in[6 * t + 1], kat_in[6 * t + 2], kat_in[6 * t + 3]),
                                      make_uint2(kat_in[6 * t + 4], kat_in[6 * t + 5]));
        kat_out[4 * t] = r.x; kat_out[4 * t + 1] = r.y; kat_out[4 * t + 2] = r.z; kat_out[4 * t + 3] = r.w;
    }
}

int launch_rng_dump(uint64_t seed, const uint32_t* key_dev, int B, int S, int n, float* out_normal, double* out_uniform,
                    const uint32_t* kat_in, uint32_t* kat_out, int n_kat, cudaStream_t st) {
    rng_dump_kernel<<<592, 256, 0, st>>>(seed, key_dev, B, S, n, out_normal, out_uniform, kat_in, kat_out, n_kat);
    return (int)cudaGetLastError();
}

}  // namespace cgb
