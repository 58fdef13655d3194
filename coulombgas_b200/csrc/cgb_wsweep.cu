// Warp-per-walker Metropolis sweep / batched log p kernel for sm_100a (n <= 32).
//
// One WARP evaluates  log p = 2 log|det D| + log|det J|  (src/logpsi.py:22-31,179) of one walker: lane i owns
// particle i, i.e. row i of the pair tensor (src/flow.py:22-26,46), block-row i of the flow Jacobian (which
// replaces jax.jacfwd, src/logpsi.py:27-28) and row i of the Slater matrix (src/slater.py:14-17).  Consequences:
//   * no block barrier inside an evaluation: every exchange is a broadcast LDS from the warp's private shared-memory
//     slice or a warp collective (REDUX / SHFL), ordered by __syncwarp; 12 warps per SM hide each other's latency
//     chains (the two LU factorisations are sequential in the pivot index).  The four warps of a CTA do START their
//     evaluations together (they pop the ready queue as a group, two block barriers per evaluation): running in lock
//     step they share instruction-cache lines -- the kernel is 126 KiB of SASS against a 32 KiB L1.5 instruction cache;
//   * the weights are operands from the CONSTANT bank (FFMA/DFMA R, R, c[3][..], R): no register, no LDS;
//   * both LU factorisations are register-resident AND rolled loops: every elimination step shifts the row one
//     register to the left while it updates it (the FMA writes position j-1), so the active column is always
//     register 0 and the loop body has compile-time register indices; the shrinking tail is cut by warp-uniform
//     breaks.  Only the pivot line travels through a double-buffered shared-memory line.  (A fully unrolled first
//     version, 330 kB of code, spent 47 % of its stall samples on instruction fetch.)
//     J: unpivoted 2x2-block LU (J = I + small for any physical flow; monitored, with a partially pivoted
//     shared-memory LU as fallback).  D: partially pivoted complex float64 LU with implicit row exchange, pivot
//     search by one REDUX per step;
//   * the Metropolis chain (src/MCMC.py:22-34) is scheduled at the granularity of ONE proposal through a bounded
//     multi-producer/multi-consumer ready queue in global memory: a free warp pops the walker that has been ready
//     longest, evaluates one proposal, writes the walker state (x, log p, accept count; L2-resident) and pushes
//     the walker back.  Every resident warp is busy until the last proposal, whatever the batch size (no wave
//     quantisation, no head-of-line blocking);
//   * batches too small to fill the warp slots run SPECULATIVELY (template parameter SPECW, below): the warps of a CTA
//     evaluate the proposals of the next 2-3 steps of one walker at once.
// Precision: FP32 pair passes with Kahan-compensated partner sums, float64 single-particle stream, float64
// Slater matrix and LU, float64 log-det sums and Metropolis test.
#include "cgb_common.cuh"
#include <stdlib.h>
#include <mutex>
#include "cgb_internal.h"

namespace cgb {

// ---------------------------------------------------------------------------------------------------
// Constant-bank weight block.  FP32 copies feed the pair passes and the Jacobian-side contractions,
// float64 copies feed the forward single-particle stream.  [in][out] like haiku's `w`.
struct WConst {
    float wt[5][16];      // Wt * log2(e)      two-particle layer (src/flow.py:46)
    float bt[16];         // bt * log2(e)
    float w0c[5][16];     // W0[4:9]           rows that see mean_j t0_ij (flow.py:34-36; s0 == 0, flow.py:16-18)
    float w1a[16][16];    // W1[0:16]          acts on h_i
    float w1b[16][16];    // W1[16:32]         acts on mean_k h_k
    float w1c[16][16];    // W1[32:48]         acts on mean_j t_ij
    float wf[16][2];      // Wf
    double expc[11];      // Taylor coefficients 1/10! .. 1/0! of the float64 exp (single-particle stream): constant-bank operands
    double logc[11];      // 1/21, 1/19, .., 1/3, 1 of the float64 log1p series
    float sabM[16];       // pass 1: magic constant 2^e >= max |a_c| of channel c (fixed-point accumulation of sum_k |a_k|)
    double sabS[16];      // ... and the weight 2^(e-23) of one unit of that fixed-point sum
    double dw0c[5][16];
    double db0[16];
    double dw1a[16][16];
    double dw1c[16][16];
    double db1[16];
    double dwf[16][2];
};
static_assert(sizeof(WConst) <= 11264, "WS_WCONST staging area too small");
__constant__ WConst cW;

// Prologue of a sweep, ONE launch: weight block, ready queue (all walkers ready for step 0), per-walker step counters,
// statistics vector, device key advance.
// ---------------------------------------------------------------------------------------------------
// CTA-shared float64 weights of the forward single-particle stream, regrouped [c4][e][kk] (output channel c = 4 c4 + kk)
// so that one group of 4 output channels reads contiguous 32-byte rows by broadcast LDS.128.  (As runtime-indexed
// constant-bank loads these 8 kB thrashed the constant cache: 16 % of all stall samples on 2 % of the instructions.)
struct WShared {
    double w1a[4][16][4];
    double w1c[4][16][4];
    double w0c[4][5][4];
};
constexpr int W_SHARED_BYTES = ((sizeof(WShared) + 15) & ~15) + 128;      // + the speculative mode's log p exchange slots

__global__ void wsweep_prologue_kernel(const double* __restrict__ p, WConst* __restrict__ w, double* __restrict__ wsh, unsigned long long* slots,
                                       unsigned cap, unsigned B, unsigned long long* ctr, int* done, double* stats, double nprop,
                                       uint32_t* key_dev, int advance_key) {
    {   // ready queue (bounded MPMC, Vyukov style): slot = (sequence << 32) | walker.  Slot i starts "filled" with walker i
        // for i < B and "free for push i" otherwise; head = 0, tail = B.  cap == 0: speculative mode, no queue.
        const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i < cap) slots[i] = i < B ? (((unsigned long long)(i + 1) << 32) | i) : ((unsigned long long)i << 32);
        if (cap && i < B) done[i] = 0;
        if (i == 0) {
            if (cap) { ctr[0] = 0ull; ctr[1] = (unsigned long long)B; ctr[2] = 0ull; }    // head, tail, spin diagnostics
            if (stats) { for (int t = 0; t < CGB_STATS_LEN; ++t) stats[t] = 0.0; stats[CGB_STATS_NPROPOSED] = nprop; }
            if (key_dev && advance_key) {
                const uint64_t k = (((uint64_t)key_dev[0] << 32) | (uint64_t)key_dev[1]) + 0x9E3779B97F4A7C15ull;
                key_dev[0] = (uint32_t)(k >> 32); key_dev[1] = (uint32_t)k;
            }
        }
    }
    if (blockIdx.x != 0) return;
    {   // float64 weights of the single-particle stream in the layout the sweep kernel's CTAs fetch by one TMA bulk copy
        constexpr int S_W0 = 96, S_W1 = 256;
        WShared* const sh = reinterpret_cast<WShared*>(wsh);
        for (int t = threadIdx.x; t < 256; t += blockDim.x) {
            const int e = t >> 4, c = t & 15;
            sh->w1a[c >> 2][e][c & 3] = p[S_W1 + t];                 // W1[e][c]        (acts on h_i)
            sh->w1c[c >> 2][e][c & 3] = p[S_W1 + 512 + t];           // W1[32 + e][c]   (acts on mean_j t_ij)
        }
        for (int t = threadIdx.x; t < 80; t += blockDim.x) {
            const int f = t >> 4, c = t & 15;
            sh->w0c[c >> 2][f][c & 3] = p[S_W0 + 64 + t];            // W0[4 + f][c]
        }
    }
    // source offsets in the flat 1074 vector (include/cgb200.h)
    constexpr int S_WT = 0, S_BT = 80, S_W0 = 96, S_B0 = 240, S_W1 = 256, S_B1 = 1024, S_WF = 1040;
    constexpr double LOG2E = 1.4426950408889634;
    for (int t = threadIdx.x; t < 80; t += blockDim.x) {
        (&w->wt[0][0])[t] = (float)(p[S_WT + t] * LOG2E);
        (&w->w0c[0][0])[t] = (float)p[S_W0 + 64 + t];
        (&w->dw0c[0][0])[t] = p[S_W0 + 64 + t];
    }
    for (int t = threadIdx.x; t < 16; t += blockDim.x) {
        {   // |a_c| <= |bt| + |Wt0| + |Wt1| + |Wt2| + |Wt3| + sqrt(2) |Wt4|  (features: cos, cos, sin, sin, d <= sqrt 2; log2 units)
            double A = fabs(p[S_BT + t]);
            for (int f = 0; f < 5; ++f) A += fabs(p[S_WT + 16 * f + t]) * (f == 4 ? 1.4142135623730951 : 1.0);
            A = fmax(A * LOG2E * 1.0001, 1e-30);
            int e = 0;
            frexp(A, &e);                        // A = m 2^e, m in [0.5, 1)  ->  2^e > A
            e = e < -100 ? -100 : (e > 100 ? 100 : e);
            w->sabM[t] = (float)ldexp(1.0, e);
            w->sabS[t] = ldexp(1.0, e - 23);
        }
        w->bt[t] = (float)(p[S_BT + t] * LOG2E);
        w->db0[t] = p[S_B0 + t];
        w->db1[t] = p[S_B1 + t];
    }
    for (int t = threadIdx.x; t < 256; t += blockDim.x) {
        (&w->w1a[0][0])[t] = (float)p[S_W1 + t];
        (&w->w1b[0][0])[t] = (float)p[S_W1 + 256 + t];
        (&w->w1c[0][0])[t] = (float)p[S_W1 + 512 + t];
        (&w->dw1a[0][0])[t] = p[S_W1 + t];
        (&w->dw1c[0][0])[t] = p[S_W1 + 512 + t];
    }
    if (threadIdx.x == 0) {
        double f = 1.0;
        for (int t = 0; t <= 10; ++t) { w->expc[10 - t] = 1.0 / f; f *= (double)(t + 1); w->logc[10 - t] = 1.0 / (double)(2 * t + 1); }
    }
    for (int t = threadIdx.x; t < 32; t += blockDim.x) {
        (&w->wf[0][0])[t] = (float)p[S_WF + t];
        (&w->dwf[0][0])[t] = p[S_WF + t];       // bf dropped: a common shift of z is a global phase of det D
    }
}

// ---------------------------------------------------------------------------------------------------
// per-warp shared-memory slice
template <int N>
struct WLay {
    static constexpr int PB = N | 1;                      // float4 blocks per staged Jacobian block-row (odd: conflict-free)
    static constexpr int J_bytes = 16 * N * PB;
    // early aliases inside the J region (all dead before pass 2 writes J):
    static constexpr int a_sig0 = 0;                      // float[N][20]   sigma0 (pitch 20: conflict-free LDS/STS.128)
    static constexpr int a_hbuf = 80 * N;                 // double[N][17]  h_i (float64)
    static constexpr int a_scr = 80 * N;                  // float[N][68]   R|T rows (overlays h once the forward pass is done)
    static constexpr int a_s1 = a_scr + 272 * N;          // float[N][20]   sigma1
    static constexpr int a_hbar = a_s1 + 80 * N;          // double[16]
    static constexpr int a_gb = a_hbar + 128;             // double[16]
    static constexpr int alias_bytes = a_gb + 128;
    static constexpr int RJ = ((J_bytes > alias_bytes ? J_bytes : alias_bytes) + 15) & ~15;
    static constexpr int o_J = 0;
    static constexpr int o_H = RJ;                        // float[N][16][2] H_k; afterwards the LU pivot lines
    static constexpr int H_bytes = 128 * N > 32 * (N + 8) ? 128 * N : 32 * (N + 8);
    static constexpr int o_pos = o_H + H_bytes;           // float4[N] half-angle table
    static constexpr int o_k = o_pos + 16 * N;            // int2[N]   integer part of the occupied k rows
    static constexpr int total = (o_k + 8 * N + 15) & ~15;
    // late aliases of the J region: Slater power tables double2[N][PT], PT = (2*(mmax+1))|1, or (large |m|) the
    // staged Slater rows double2[N][PB]
    static constexpr int MMAX_TAB = (RJ / (16 * N) - 1) / 2 - 1;
    static_assert(16 * N * PB <= RJ, "Slater staging must fit the J region");
    static_assert(16 * ((N + 18) & ~1) <= H_bytes, "FP32 Slater pivot lines must fit the H region");
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float4 shfl4w(const float4 v, int src) {
    return make_float4(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src),
                       __shfl_sync(0xffffffffu, v.z, src), __shfl_sync(0xffffffffu, v.w, src));
}
__device__ __forceinline__ double2 shfl2dw(const double2 v, int src) {
    return make_double2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}
// shared-memory loads the compiler may not move: software prefetch of the next chunk of a pivot line
__device__ __forceinline__ float4 lds_pinned_f4(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ double2 lds_pinned_d2(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
    return v;
}
// predicated shared-memory store (no divergent branch around a single-lane publish)
__device__ __forceinline__ void sts_pred_f4(unsigned addr, const float4 v, const bool pred) {
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %5, 0; @p st.shared.v4.f32 [%0], {%1, %2, %3, %4}; }"
                 :: "r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void sts_pred_f2x2(unsigned addr, const float2 a, const float2 b, const bool pred) {
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %5, 0; @p st.shared.v4.f32 [%0], {%1, %2, %3, %4}; }"
                 :: "r"(addr), "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "r"((int)pred) : "memory");
}
__device__ __forceinline__ void sts_pred_f2(unsigned addr, const float2 a, const bool pred) {
    asm volatile("{ .reg .pred p; setp.ne.s32 p, %3, 0; @p st.shared.v2.f32 [%0], {%1, %2}; }"
                 :: "r"(addr), "f"(a.x), "f"(a.y), "r"((int)pred) : "memory");
}
// packed FP32 FMA of sm_100 (FFMA2: two IEEE fmas per lane and instruction; the scalar operand is broadcast by the
// instruction itself, `R.F32`): same results as two fmaf, half the issue slots
__device__ __forceinline__ float2 ffma2s(const float a, const float2 b, const float2 c) {
    return __ffma2_rn(make_float2(a, a), b, c);
}
__device__ __forceinline__ double warp_sum_f64(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
#ifndef CGBW_CTA_SYNC
#define CGBW_CTA_SYNC 1     // 1: the warps of a CTA pop the ready queue together (lock step: shared instruction-cache lines)
#endif
#ifndef CGBW_PIN_LDS
#define CGBW_PIN_LDS 1      // 1: the LUs' next-chunk pivot-line loads are volatile asm (the compiler otherwise sinks them to their use)
#endif
#ifndef CGBW_PACK_PRE
#define CGBW_PACK_PRE 1     // pass-1 pre-activations of channel pairs through FFMA2 (weight pairs via uniform registers): +0.9 %
#endif
#ifndef CGBW_PACK_PRE23
#define CGBW_PACK_PRE23 1   // the same for the pre-activations of pass 2+3: +0.9 %
#endif
#ifndef CGBW_UFORM
#define CGBW_UFORM 1        // pass 2+3: accumulate U = Q + sum_c (sigma'_c R_c) Wt_c^T (2x5) and contract with E once per pair
#endif
#ifndef CGBW_PROF
#define CGBW_PROF 0         // 1: developer build that accumulates per-phase cycle counts of every warp (tools/phase_clocks.py)
#endif
#if CGBW_PROF
#define PROF_MARK(i) do { if (lane == 0) { const long long t_ = clock64(); s_prof[threadIdx.x >> 5][i] += (unsigned long long)(t_ - s_prof_t0[threadIdx.x >> 5]); s_prof_t0[threadIdx.x >> 5] = t_; } } while (0)
#else
#define PROF_MARK(i) do { } while (0)
#endif
#ifndef CGBW_FIXSUM
#define CGBW_FIXSUM 1       // pass 1: sum_k |a_k| and sum_k d_k in 32-bit fixed point (FADD with a magic constant + IADD) instead of
#endif                      // F2F + DADD per term (the conversions run on the quarter-rate XU pipe); trig feature sums separable
#ifndef CGBW_STAGGER_NS
#define CGBW_STAGGER_NS 0   // ready-queue kernel: start delay between the co-resident CTAs of an SM (0 = none)
#endif
#ifndef CGBW_WPC
#define CGBW_WPC 4          // warps per CTA of the ready-queue kernel (12 / CGBW_WPC CTAs per SM)
#endif
#ifndef CGBW_KAHAN
#define CGBW_KAHAN 1        // 0: plain FP32 partner sums (precision experiment)
#endif
__device__ __forceinline__ void kahan_add_w(float& sum, float& comp, float x) {
#if !CGBW_KAHAN
    sum += x;
    return;
#endif
    const float y = x - comp;
    const float t = sum + y;
    comp = (t - sum) - y;
    sum = t;
}
// float64 softplus (forward value) + FP32 logistic sigmoid (Jacobian side) of a float64 pre-activation.
// softplus(x) = max(x,0) + log1p(t), t = exp(-|x|): exp by range reduction + degree-10 Taylor,
// log1p(t) = 2 atanh(t/(2+t)) by a degree-10 series in u^2 <= 1/9 with a Newton-refined FP32 reciprocal (~1e-14).
__device__ __forceinline__ void softplus_sigmoid_f64w(double pre, double& sp, float& sg) {
    const double y = fmax(-fabs(pre) * 1.4426950408889634074, -1000.0);
    const double kf = rint(y);
    const double f = (y - kf) * 0.69314718055994530942;
    double pz = cW.expc[0];
#pragma unroll
    for (int q = 1; q <= 10; ++q) pz = fma(pz, f, cW.expc[q]);
    const double t = pz * __hiloint2double(((int)kf + 1023) << 20, 0);
    const double d = 2.0 + t;
    const double r0 = (double)rcp_approx((float)d);
    const double r = fma(r0, fma(-d, r0, 1.0), r0);
    const double u = t * r, w = u * u;
    double a = cW.logc[0];
#pragma unroll
    for (int q = 1; q <= 10; ++q) a = fma(a, w, cW.logc[q]);
    sp = fmax(pre, 0.0) + 2.0 * u * a;
    const float ef = (float)t, rr = rcp_approx(1.f + ef);
    sg = pre >= 0.0 ? rr : ef * rr;
}

// pair features of (row i, partner j) from the half-angle table: cos/sin of theta = 2 pi (x_i - x_j)/L by angle
// addition and d^2 = sin^2(theta_x/2) + sin^2(theta_y/2)  (src/flow.py:22-25).  Table entries are laid out
// (cos pi u_x, cos pi u_y, sin pi u_x, sin pi u_y) so that the x and y components go through packed FP32 instructions.
struct WHalf { float2 sh, ch; };            // sin, cos of the half-angle differences (x, y)
__device__ __forceinline__ WHalf whalf_diff(const float4 pi_, const float2 nci, const float4 pj) {
    const float2 ci = make_float2(pi_.x, pi_.y), si = make_float2(pi_.z, pi_.w);
    const float2 cj = make_float2(pj.x, pj.y), sj = make_float2(pj.z, pj.w);
    WHalf h;
    h.sh = __ffma2_rn(si, cj, __fmul2_rn(nci, sj));          // s_i c_j - c_i s_j   (nci = -c_i)
    h.ch = __ffma2_rn(ci, cj, __fmul2_rn(si, sj));           // c_i c_j + s_i s_j
    return h;
}
struct WPairFeat { float2 c, s; float d2; };
__device__ __forceinline__ WPairFeat wpair_features(const float4 pi_, const float2 nci, const float4 pj) {
    const WHalf h = whalf_diff(pi_, nci, pj);
    WPairFeat f;
    f.c = __ffma2_rn(__fmul2_rn(h.sh, make_float2(-2.f, -2.f)), h.sh, make_float2(1.f, 1.f));
    f.s = __fmul2_rn(__fmul2_rn(h.sh, make_float2(2.f, 2.f)), h.ch);
    f.d2 = fmaf(h.sh.x, h.sh.x, h.sh.y * h.sh.y);
    return f;
}

// Partially pivoted scalar LU of the staged Jacobian (blocked shared-memory layout), one warp, lane owns scalar
// rows lane and lane+32.  Fallback of the unpivoted register LU; returns this lane's sum of log2|pivot|.
template <int N>
__device__ __noinline__ float pivoted_lu_smem(float* Jf, int lane) {
    constexpr int PB = WLay<N>::PB, N2 = 2 * N;
    auto el = [&](int rr, int cc) -> float& { return Jf[((rr >> 1) * PB + (cc >> 1)) * 4 + (rr & 1) * 2 + (cc & 1)]; };
    const int r0 = lane, r1 = lane + 32;
    bool d0 = !(r0 < N2), d1 = !(r1 < N2);
    float mylog = 0.f;
    for (int k = 0; k < N2; ++k) {
        const float c0 = d0 ? 0.f : el(r0, k), c1 = d1 ? 0.f : el(r1, k);
        const unsigned k0 = d0 ? 0u : ((__float_as_uint(fabsf(c0)) & ~0x3Fu) | (unsigned)r0);
        const unsigned k1 = d1 ? 0u : ((__float_as_uint(fabsf(c1)) & ~0x3Fu) | (unsigned)r1);
        const unsigned g = __reduce_max_sync(0xffffffffu, max(k0, k1));
        const int pr = (int)(g & 0x3Fu);
        const float piv = el(pr, k);
        if (!d0) {
            if (r0 == pr) { d0 = true; mylog += lg2_approx(fabsf(piv)); }
            else { const float l = -c0 / piv; for (int j = k + 1; j < N2; ++j) el(r0, j) = fmaf(l, el(pr, j), el(r0, j)); }
        }
        if (!d1) {
            if (r1 == pr) { d1 = true; mylog += lg2_approx(fabsf(piv)); }
            else { const float l = -c1 / piv; for (int j = k + 1; j < N2; ++j) el(r1, j) = fmaf(l, el(pr, j), el(r1, j)); }
        }
        __syncwarp();
    }
    return mylog;
}

// Float64 Slater log-determinant (fallback of the FP32 factorisation of the main kernel, and the path for wave vectors
// beyond the power tables): row r = lane r, partially pivoted complex LU with implicit row exchange, register-resident,
// shift-left (the active column is always register 0).  After its first update every unfinished row offers its new
// register 0; one REDUX picks the next pivot row, whose pivot is shuffled to all lanes (FP32-seeded reciprocal + one float64
// Newton step) and whose row is published through the shared-memory line.  Returns log|det| (without the L^-n factor), -inf for
// a singular matrix (slater.py:18).
template <int N>
__device__ __noinline__ double slater_logdet_f64(unsigned char* sm, const int2* sK, const int lane, const int kself, const bool rv,
                                                 const int mmax, const double wx, const double wy) {
    using L_ = WLay<N>;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PB = L_::PB;
    double2* const sPt = reinterpret_cast<double2*>(sm + L_::o_J);
    __syncwarp();
    double aD;
    {
    double2 Dr[N];
        if (mmax <= L_::MMAX_TAB) {
            const int PT = (2 * (mmax + 1)) | 1;
            double2* tb = sPt + kself * PT;
            if (rv) {
#pragma unroll 1
                for (int comp = 0; comp < 2; ++comp) {
                    const double wv = comp ? wy : wx;
                    double sn, cs;
                    sincospi(2.0 * (wv - rint(wv)), &sn, &cs);
                    double2* t = tb + comp * (mmax + 1);
                    double2 pw = make_double2(1.0, 0.0);
                    t[0] = pw;
                    for (int m = 1; m <= mmax; ++m) {
                        pw = make_double2(pw.x * cs - pw.y * sn, pw.x * sn + pw.y * cs);
                        t[m] = pw;
                    }
                }
            }
            __syncwarp();
            const double2* ty = tb + (mmax + 1);
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const int2 mk = sK[j];
                double2 X = tb[abs(mk.x)], Y = ty[abs(mk.y)];
                X.y = mk.x < 0 ? -X.y : X.y;
                Y.y = mk.y < 0 ? -Y.y : Y.y;
                Dr[j] = make_double2(X.x * Y.x - X.y * Y.y, X.x * Y.y + X.y * Y.x);
            }
        } else {
            double2* row = sPt + kself * PB;
#pragma unroll 1
            for (int j = 0; j < N; ++j) {
                const int2 mk = sK[j];
                double ph = fma((double)mk.x, wx, (double)mk.y * wy);
                ph -= rint(ph);
                double sn, cs;
                sincospi(2.0 * ph, &sn, &cs);
                if (rv) row[j] = make_double2(cs, sn);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < N; ++j) Dr[j] = row[j];
        }
        __syncwarp();     // H region (lines) free, tables read

        // ---------------- Slater LU: partially pivoted, implicit row exchange, register-resident, rolled ---------
        // Same shift-left scheme: the active column is register 0.  After its first update every unfinished
        // row offers its new register 0; one REDUX picks the next pivot row, which publishes 1/pivot (the
        // reciprocal chain overlaps the remaining updates) and then its row.  Finished rows compute garbage.
        bool done = !rv, singular = false;
        double mypn = 1.0;
        constexpr int LP = N + 8;
        double2* const dline0 = reinterpret_cast<double2*>(sm + L_::o_H);    // [2][LP]
        auto offer = [&](const double2 e) -> unsigned {
            const float mag = (float)(fabs(e.x) + fabs(e.y));
            const unsigned key = done ? 0u : ((__float_as_uint(mag) & ~0x1Fu) | (unsigned)lane);
            return __reduce_max_sync(FULL, key);
        };
        // 1/pivot, computed by EVERY lane from the shuffled pivot (no divergent branch, no shared-memory round trip):
        // FP32 seed of 1/|e|^2 from the FP32-rounded entry + one float64 Newton step (2^-23 squared)
        double pn = 1.0;
        auto pivot_inv = [&](const double2 e) -> double2 {
            pn = fma(e.x, e.x, e.y * e.y);
            const float ex = (float)e.x, ey = (float)e.y;
            const double r0 = (double)rcp_approx(fmaf(ex, ex, ey * ey));
            const double ip = r0 * (2.0 - pn * r0);
            return make_double2(e.x * ip, -e.y * ip);
        };
        unsigned wk = offer(Dr[0]);
        double2 iv = pivot_inv(shfl2dw(Dr[0], (int)(wk & 31u)));
        bool win = !done && lane == (int)(wk & 31u) && (wk >> 5) != 0u;
        if (win) {
            mypn = pn;
#pragma unroll
            for (int j = 1; j < N; ++j) dline0[j] = Dr[j];
        }
        __syncwarp();
#pragma unroll 1
        for (int k = 0; k < N; ++k) {
            const double2* cur = dline0 + (k & 1) * LP;
            double2* nxt = dline0 + ((k + 1) & 1) * LP;
            const unsigned dcur32 = (unsigned)__cvta_generic_to_shared(cur);
            const int lim = N - 1 - k;
            singular |= (wk >> 5) == 0u;
            done |= win;
            if (lim == 0) break;
            const double2 c = Dr[0];
            const double lx = -(c.x * iv.x - c.y * iv.y);               // l = -c / pivot
            const double ly = -(c.x * iv.y + c.y * iv.x);
            double2 pq[4] = {cur[1], cur[2], cur[3], cur[4]};
#pragma unroll
            for (int j0 = 1; j0 < N; j0 += 4) {
                if (j0 > lim) break;
                double2 nq[4];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#if CGBW_PIN_LDS
                    nq[t] = lds_pinned_d2(dcur32 + 16u * (unsigned)(j0 + 4 + t));
#else
                    nq[t] = cur[j0 + 4 + t];
#endif
#pragma unroll
                for (int t = 0; t < 4; ++t)                               // independent first halves (written straight to the
                    if (j0 + t < N) {                                     // shifted slot): covers the DFMA latency
                        const double2 o = Dr[j0 + t];
                        Dr[j0 + t - 1] = make_double2(fma(-ly, pq[t].y, o.x), fma(ly, pq[t].x, o.y));
                    }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int j = j0 + t;
                    if (j < N) {
                        double2 v = Dr[j - 1];
                        v.x = fma(lx, pq[t].x, v.x);
                        v.y = fma(lx, pq[t].y, v.y);
                        Dr[j - 1] = v;
                        if (j == 1) {
                            wk = offer(v);
                            iv = pivot_inv(shfl2dw(v, (int)(wk & 31u)));
                            win = !done && lane == (int)(wk & 31u) && (wk >> 5) != 0u;
                            if (win) mypn = pn;
                        }
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) pq[t] = nq[t];
            }
            if (win) {
#pragma unroll
                for (int j = 1; j < N - 1; ++j) {
                    if ((j & 3) == 1 && j >= lim) break;
                    nxt[j] = Dr[j];
                }
            }
            __syncwarp();
        }
        aD = (rv && mypn != 1.0) ? 0.5 * log(mypn) : 0.0;
        aD = warp_sum_f64(aD);
        if (singular) aD = -INFINITY;             // slogdet of a singular matrix: log|det| = -inf (slater.py:18)
    }
    return aD;
}

// SPECW = 4 or 8: speculative ("prefetching") Metropolis for batches that leave most warp slots empty (SPECW = 0: the
// ready-queue kernel).  One CTA of SPECW warps owns one walker for the whole sweep; per round the warps evaluate,
// concurrently, the complete binary decision tree of the next D = log2(SPECW) steps -- the next proposal (node 1), BOTH
// proposals of the step after it (node 2: first one rejected, node 3: accepted), ... -- and the last warp one node of
// level D (all-rejected or all-accepted path, whichever the running acceptance favours; in the first round: log p of
// the initial walker instead).  The proposal noise and the uniforms are keyed by (walker, step), so a node's
// configuration is known before the decisions above it; after ONE block barrier every warp replays the accept/reject
// decisions along the realised path (MCMC.py:28-33, same arithmetic, same order) and the chain advances D to D+1 steps
// per evaluation latency.  The trajectory is bit-identical to the sequential kernel's.
template <int N, int SPECW>
__global__ void __launch_bounds__(SPECW == 8 ? 256 : (SPECW == 4 ? 128 : 32 * CGBW_WPC), SPECW == 8 ? 1 : (SPECW == 4 ? 3 : 12 / CGBW_WPC))
wsweep_kernel(const SweepParams p) {
    constexpr bool SPEC = SPECW != 0;
    static_assert(SPECW == 0 || SPECW == 4 || SPECW == 8, "speculation tree: 4 or 8 warps");
    extern __shared__ __align__(16) unsigned char smem_raw[];
#if CGBW_PROF
    __shared__ unsigned long long s_prof[8][12];
    __shared__ long long s_prof_t0[8];
    if (threadIdx.x < 96) (&s_prof[0][0])[threadIdx.x] = 0ull;
#endif
    using L_ = WLay<N>;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PB = L_::PB;
    // lane id and the warp's slice offset pass through volatile asm so that the compiler keeps them in registers:
    // otherwise, at the 168-register cap, it re-derives them from S2R (tid, shared-window base) inside the hot loops.
    // (a plain mov is still rematerialised by ptxas; the result of a shuffle is not)
    const int lane = __shfl_sync(0xffffffffu, (int)(threadIdx.x & 31), (int)(threadIdx.x & 31));
    const int wofs = __shfl_sync(0xffffffffu, (int)(W_SHARED_BYTES + (threadIdx.x >> 5) * L_::total), 0);
    {   // CTA-shared float64 weights: regrouped by the prologue kernel into the WShared layout, one TMA bulk copy per CTA
        __shared__ __align__(8) unsigned long long s_mbar;
        static_assert(sizeof(WShared) % 16 == 0, "bulk copies move multiples of 16 bytes");
        tma_stage_block(smem_raw, p.wshared, (unsigned)sizeof(WShared), &s_mbar);   // (its barriers: before the warps go their own ways)
    }
    // the slice base as a pinned 32-bit shared-window address (the window base is a function of %cluster_ctarank that
    // the compiler otherwise re-derives with S2UR + ULEA at every use site)
    const unsigned sbase32 = __shfl_sync(0xffffffffu, (unsigned)__cvta_generic_to_shared(smem_raw) + (unsigned)wofs, 0);
    unsigned char* sm = reinterpret_cast<unsigned char*>(__cvta_shared_to_generic((size_t)sbase32));
    const WShared* const sW = reinterpret_cast<const WShared*>(__cvta_shared_to_generic((size_t)(sbase32 - (unsigned)wofs)));
    float4* const sJ = reinterpret_cast<float4*>(sm + L_::o_J);
    float* const sSig0 = reinterpret_cast<float*>(sm + L_::a_sig0);
    double* const sHbuf = reinterpret_cast<double*>(sm + L_::a_hbuf);
    float* const sScr = reinterpret_cast<float*>(sm + L_::a_scr);
    float* const sS1 = reinterpret_cast<float*>(sm + L_::a_s1);
    double* const sHbar = reinterpret_cast<double*>(sm + L_::a_hbar);
    double* const sGb = reinterpret_cast<double*>(sm + L_::a_gb);
    float* const sH = reinterpret_cast<float*>(sm + L_::o_H);
    float4* const sPos = reinterpret_cast<float4*>(sm + L_::o_pos);
    int2* const sK = reinterpret_cast<int2*>(sm + L_::o_k);
    double2* const sPt = reinterpret_cast<double2*>(sm + L_::o_J);

    const bool rv = lane < N;
    const int kself = rv ? lane : N - 1;
    const unsigned lanebits = 0xffffffe0u | (unsigned)lane;   // lane-dependent (not warp-uniform) mask for the prefetch keep-alive uses
    const float invn = 1.f / (float)N;
    const double invL = 1.0 / p.L;
    const double nlogL = (double)N * log(p.L);               // D carries L^{-dim/2} per entry (slater.py:17)
    // (2 pi / L) / n, pinned in a register (result of a shuffle): at the register cap ptxas otherwise rematerialises the
    // float64 product + conversion inside the pair loops
    const float alpha = __shfl_sync(0xffffffffu, (float)(2.0 * 3.14159265358979323846 * invL) * invn, 0);
    const unsigned long long total = (unsigned long long)p.B * (unsigned long long)(p.mc_steps + 1);
    const unsigned long long qmask = (unsigned long long)p.qcap - 1ull;

    // proposal noise of (walker b, step s) for this lane's particle and the Metropolis uniform of (b, s): inputs in
    // parity mode, Philox streams keyed by (walker, step) otherwise -- independent of which warp asks, and when
    auto add_noise = [&](double2& xq, const int b, const int s) {                  // MCMC.py:26
        double nzx, nzy;
        if (p.noise) {
            double2 nz = make_double2(0.0, 0.0);
            if (rv) nz = __ldg(reinterpret_cast<const double2*>(p.noise + (((size_t)(s - 1) * p.B + b) * N + lane) * 2));
            nzx = nz.x; nzy = nz.y;
        } else {
            const float2 g = rng_normal2(rng_seed_of(p.seed, p.key_dev), (uint32_t)b, (uint32_t)s, (uint32_t)lane);
            nzx = (double)g.x; nzy = (double)g.y;
        }
        xq.x += p.sigma * nzx;
        xq.y += p.sigma * nzy;
    };
    auto uniform_of = [&](const int b, const int s) -> double {                    // MCMC.py:29
        if (p.uniforms) return __ldg(p.uniforms + (size_t)(s - 1) * p.B + b);
        return rng_uniform(rng_seed_of(p.seed, p.key_dev), (uint32_t)b, (uint32_t)s);
    };

    // speculative mode: the walker's chain state, replicated identically in the four warps of the CTA
    int sp_b = blockIdx.x, sp_t = 0, sp_nacc = 0, sp_round = 0;
    bool sp_load = true;
    double2 sp_x = make_double2(0.0, 0.0);
    double sp_logp = 0.0;
    double* const sLp = reinterpret_cast<double*>(smem_raw + W_SHARED_BYTES - 128);    // [2][8] log p of the round's nodes

#if CGBW_PROF
    if (lane == 0) s_prof_t0[threadIdx.x >> 5] = clock64();
#endif
#if CGBW_STAGGER_NS
    // the co-resident CTAs of an SM (round-robin placement: CTA i on SM i % #SMs) start their first evaluation a third of an
    // evaluation apart, so that their phases (and with them the XU / FP64 pipe demand) are out of step from the beginning
    if constexpr (!SPEC) {
        const unsigned wave = blockIdx.x / p.n_sms;
        for (unsigned w = 0; w < wave; ++w) __nanosleep(CGBW_STAGGER_NS);
    }
#endif
    for (;;) {
        int b = 0, step = 0;
        bool active = true;
        double2 xv = make_double2(0.0, 0.0);
        double logp_cur = 0.0;
        if constexpr (SPEC) {
            if (sp_b >= p.B) break;
            b = sp_b;
            if (sp_load) {
                if (rv) sp_x = __ldcg(reinterpret_cast<const double2*>(p.x + ((size_t)b * N + lane) * 2));
                sp_t = 0; sp_nacc = 0; sp_load = false;
            }
            const int w = (int)(threadIdx.x >> 5);
            int node = w + 1;                                   // heap numbering: child of m is 2m + accepted
            if (w == SPECW - 1) node = sp_t == 0 ? 0 : (2 * sp_nacc > sp_t ? 2 * SPECW - 1 : SPECW);
            xv = sp_x;
            if (node > 0) {
                const int depth = 31 - __clz(node);
                step = sp_t + 1 + depth;
                active = step <= p.mc_steps;
                if (active) {
                    for (int j = 0; j < depth; ++j)
                        if ((node >> (depth - 1 - j)) & 1) add_noise(xv, b, sp_t + 1 + j);     // accepted moves above this node
                    add_noise(xv, b, step);
                }
            }
        } else {
        // ---------------- pop the walker that has been ready longest ---------------------------------------
        {
            unsigned long long idx = 0;
#if CGBW_CTA_SYNC
            // the warps of a CTA pop TOGETHER and therefore run their evaluations in approximate lock step: they share
            // instruction-cache lines (125 KiB of code against a 32 KiB L1.5 instruction cache; with 12 free-running warps
            // per SM 12 % of the stall samples were instruction fetch).  Measured: 6.27 -> 5.88 ms at n=29, B=2048.
            __shared__ unsigned long long s_idx;
            __syncthreads();
            if (threadIdx.x == 0) s_idx = atomicAdd(p.qctr, (unsigned long long)(blockDim.x >> 5));
            __syncthreads();
            if (s_idx >= total) break;
            idx = s_idx + (threadIdx.x >> 5);
            if (idx >= total) continue;
#else
            if (lane == 0) idx = atomicAdd(p.qctr, 1ull);
            idx = __shfl_sync(FULL, idx, 0);
            if (idx >= total) break;
#endif
            if (lane == 0) {
                unsigned long long* slot = p.queue + (idx & qmask);
                const unsigned want = (unsigned)(idx + 1ull);
                unsigned long long v;
                unsigned spins = 0;
                while ((unsigned)((v = ld_acquire_u64(slot)) >> 32) != want) { __nanosleep(40); ++spins; }
                if (spins) atomicAdd(p.qctr + 2, (unsigned long long)spins);
                st_release_u64(slot, (unsigned long long)(unsigned)(idx + (unsigned long long)p.qcap) << 32);   // free for push idx+cap
                b = (int)(unsigned)(v & 0xffffffffull);
            }
            b = __shfl_sync(FULL, b, 0);
        }
        __syncwarp();
        step = p.mc_steps > 0 ? __ldcg(p.done + b) : 0;

        // ---------------- walker state, proposal (MCMC.py:26) ------------------------------------------------
        if (rv) xv = __ldcg(reinterpret_cast<const double2*>((step == 0 ? p.x : p.x_out) + ((size_t)b * N + lane) * 2));
        if (step > 0) logp_cur = __ldcg(p.ws_logp + b);
        if (step > 0) add_noise(xv, b, step);
        }
        double lp = 0.0, ldD = 0.0, aJ = 0.0;
#ifdef CGBW_DIAG_ERRSUM
        double aJ_diag = 0.0;
#endif
        if (!SPEC || active) {
        // ---------------- reduced coordinates, occupied k rows ------------------------------------------------
        int mmax;
        {
            int2 mk = make_int2(0, 0);
            if (rv) {
                int si = __ldg(p.state_idx + (size_t)b * N + lane);
                si = min(max(si, 0), p.M - 1);
                mk.x = (int)rint(__ldg(p.ktable + 2 * si) - p.tw0);
                mk.y = (int)rint(__ldg(p.ktable + 2 * si + 1) - p.tw1);
                sK[lane] = mk;
            }
            mmax = __reduce_max_sync(FULL, max(abs(mk.x), abs(mk.y)));
        }
        double ux = xv.x * invL, uy = xv.y * invL;
        ux -= floor(ux); uy -= floor(uy);                       // features are L-periodic (flow.py:22-25)
        float4 pos;                                             // (cos pi ux, cos pi uy, sin pi ux, sin pi uy)
        {
            float cs2[2], sn2[2];
#pragma unroll 1
            for (int comp = 0; comp < 2; ++comp) {
                const double u = comp ? uy : ux;
                const float uh = (float)u, ul = (float)(u - (double)uh);
                float sn, cs;
                sincospif(uh, &sn, &cs);                         // half angle pi*u, first-order hi/lo correction
                const float pl = kPi * ul;
                const float c1 = fmaf(-pl, sn, cs), s1_ = fmaf(pl, cs, sn);
                if (comp) { cs2[1] = c1; sn2[1] = s1_; } else { cs2[0] = c1; sn2[0] = s1_; }
            }
            pos = make_float4(cs2[0], cs2[1], sn2[0], sn2[1]);
        }
        if (rv) sPos[lane] = pos;
        const float2 npc = make_float2(-pos.x, -pos.y);
        __syncwarp();

        PROF_MARK(0);
        // ---------------- pass 1: two-particle stream, means over partners (flow.py:34-36,46) ----------------
        double dzx = 0.0, dzy = 0.0;
        {
            double md[16], m0d[5];
            {
                // sum_k softplus2(a_k) = 1/2 (sum_k a_k + sum_k |a_k|) + log2 prod_k (1 + 2^-|a_k|):
                //   sum_k a_k follows from the feature sums (linear), sum_k |a_k| is accumulated in float64, and the product
                //   (29 factors in (1,2]) is one FFMA per term -- ONE logarithm per channel and evaluation instead of 29.
#if CGBW_FIXSUM
                // Fixed-point partner sums: t = |a| + M with M = 2^e > max |a| (bound from the weights, prologue kernel) has the
                // value |a| / 2^(e-23) in its mantissa field, and float bit patterns are linear in [M, 2M]: the INTEGER sum of the
                // bit patterns minus N bits(M) is sum_k |a_k| in units of 2^(e-23), exactly (mod 2^32 arithmetic).  One FADD + one
                // IADD per term instead of F2F + DADD; the rounding of |a| to the 2^(e-23) grid (e = 4: 2e-6) is below the FP32
                // error a_k carries already.  The four trigonometric feature sums are separable,
                //   sum_k cos(th_i - th_k) = cos th_i C + sin th_i S,   C = sum_k cos th_k, S = sum_k sin th_k  (float64 warp sums),
                // and need no pair loop at all; sum_k d_ik uses the same fixed-point scheme with M = 2 (d <= sqrt 2).
                double m0s[5];
                unsigned sabi[16], sdi = 0u;
                float P[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) { sabi[c] = 0u; P[c] = 1.f; }
                {
                    const float cthx = fmaf(-2.f * pos.z, pos.z, 1.f), sthx = 2.f * pos.z * pos.x;     // full angles 2 pi u
                    const float cthy = fmaf(-2.f * pos.w, pos.w, 1.f), sthy = 2.f * pos.w * pos.y;
                    const double Cx = warp_sum_f64(rv ? (double)cthx : 0.0), Sx = warp_sum_f64(rv ? (double)sthx : 0.0);
                    const double Cy = warp_sum_f64(rv ? (double)cthy : 0.0), Sy = warp_sum_f64(rv ? (double)sthy : 0.0);
                    m0s[0] = fma((double)cthx, Cx, (double)sthx * Sx); m0s[1] = fma((double)cthy, Cy, (double)sthy * Sy);
                    m0s[2] = fma((double)sthx, Cx, -(double)cthx * Sx); m0s[3] = fma((double)sthy, Cy, -(double)cthy * Sy);
                }
                float btr[16];                                    // biases in registers (an FFMA takes one constant operand)
#pragma unroll
                for (int c = 0; c < 16; ++c) btr[c] = cW.bt[c];
                float4 pn = sPos[0];
#pragma unroll 2
                for (int k = 0; k < N; ++k) {
                    const WPairFeat f = wpair_features(pos, npc, pn);
                    pn = sPos[k + 1 < N ? k + 1 : k];             // next partner's table entry ahead of its use
                    const float d = sqrt_approx(f.d2);            // 0 on the diagonal (flow.py:25)
                    sdi += __float_as_uint(d + 2.0f);
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) {              // pre-activations of a channel pair through packed FMAs
                        const int c = 2 * c2;
                        const float2 a2 = ffma2s(d, make_float2(cW.wt[4][c], cW.wt[4][c + 1]),
                                          ffma2s(f.s.y, make_float2(cW.wt[3][c], cW.wt[3][c + 1]),
                                          ffma2s(f.s.x, make_float2(cW.wt[2][c], cW.wt[2][c + 1]),
                                          ffma2s(f.c.y, make_float2(cW.wt[1][c], cW.wt[1][c + 1]),
                                          ffma2s(f.c.x, make_float2(cW.wt[0][c], cW.wt[0][c + 1]), make_float2(btr[c], btr[c + 1]))))));
                        sabi[c] += __float_as_uint(fabsf(a2.x) + cW.sabM[c]);
                        sabi[c + 1] += __float_as_uint(fabsf(a2.y) + cW.sabM[c + 1]);
                        const float2 e2 = make_float2(ex2_approx(-fabsf(a2.x)), ex2_approx(-fabsf(a2.y)));
                        const float2 pp = __ffma2_rn(make_float2(P[c], P[c + 1]), e2, make_float2(P[c], P[c + 1]));    // P (1 + 2^-|a|), packed
                        P[c] = pp.x; P[c + 1] = pp.y;
                    }
                }
                m0s[4] = (double)(int)(sdi - (unsigned)N * 0x40000000u) * (1.0 / 4194304.0);      // units of 2^-22
                double sab[16];
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    sab[c] = (double)(int)(sabi[c] - (unsigned)N * __float_as_uint(cW.sabM[c])) * cW.sabS[c];
#else
                double sab[16], m0s[5];
                float P[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) { sab[c] = 0.0; P[c] = 1.f; }
#pragma unroll
                for (int f = 0; f < 5; ++f) m0s[f] = 0.0;
                float btr[16];                                    // biases in registers (an FFMA takes one constant operand)
#pragma unroll
                for (int c = 0; c < 16; ++c) btr[c] = cW.bt[c];
                float4 pn = sPos[0];
#pragma unroll 2
                for (int k = 0; k < N; ++k) {
                    const WPairFeat f = wpair_features(pos, npc, pn);
                    pn = sPos[k + 1 < N ? k + 1 : k];             // next partner's table entry ahead of its use
                    const float d = sqrt_approx(f.d2);            // 0 on the diagonal (flow.py:25)
                    m0s[0] += (double)f.c.x; m0s[1] += (double)f.c.y; m0s[2] += (double)f.s.x; m0s[3] += (double)f.s.y;
                    m0s[4] += (double)d;
#if CGBW_PACK_PRE
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) {              // pre-activations of a channel pair through packed FMAs
                        const int c = 2 * c2;
                        const float2 a2 = ffma2s(d, make_float2(cW.wt[4][c], cW.wt[4][c + 1]),
                                          ffma2s(f.s.y, make_float2(cW.wt[3][c], cW.wt[3][c + 1]),
                                          ffma2s(f.s.x, make_float2(cW.wt[2][c], cW.wt[2][c + 1]),
                                          ffma2s(f.c.y, make_float2(cW.wt[1][c], cW.wt[1][c + 1]),
                                          ffma2s(f.c.x, make_float2(cW.wt[0][c], cW.wt[0][c + 1]), make_float2(btr[c], btr[c + 1]))))));
                        sab[c] += (double)fabsf(a2.x);
                        P[c] = fmaf(P[c], ex2_approx(-fabsf(a2.x)), P[c]);
                        sab[c + 1] += (double)fabsf(a2.y);
                        P[c + 1] = fmaf(P[c + 1], ex2_approx(-fabsf(a2.y)), P[c + 1]);
                    }
#else
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float a = fmaf(d, cW.wt[4][c], fmaf(f.s.y, cW.wt[3][c], fmaf(f.s.x, cW.wt[2][c],
                                        fmaf(f.c.y, cW.wt[1][c], fmaf(f.c.x, cW.wt[0][c], btr[c])))));
                        sab[c] += (double)fabsf(a);
                        P[c] = fmaf(P[c], ex2_approx(-fabsf(a)), P[c]);
                    }
#endif
                }
#endif
                const double invnd = 1.0 / (double)N;
#pragma unroll
                for (int f = 0; f < 5; ++f) m0d[f] = m0s[f] * invnd;
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    double sa = (double)N * (double)cW.bt[c];
#pragma unroll
                    for (int f = 0; f < 5; ++f) sa = fma(m0s[f], (double)cW.wt[f][c], sa);
                    const int pb = __float_as_int(P[c]);
                    const float mant = __int_as_float((pb & 0x007fffff) | 0x3f800000);         // P = 2^ex * mant, mant in [1,2)
                    const double lg = (double)((pb >> 23) - 127) + (double)lg2_approx(mant);   // abs error 2^-22 on a sum ~ n
                    md[c] = (0.5 * (sa + sab[c]) + lg) * (0.693147180559945309417 * invnd);
                }
            }
            PROF_MARK(1);
            // ---------------- single-particle stream (flow.py:45,51-55): forward values in float64 -----------
            // rolled over groups of 4 output channels (uniform constant-bank indices)
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
                double hq[4];
                float sq[4];
#pragma unroll
                double pre[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) pre[kk] = cW.db0[4 * c4 + kk];
#pragma unroll
                for (int f = 0; f < 5; ++f) {
                    const double2 wa = *reinterpret_cast<const double2*>(&sW->w0c[c4][f][0]);
                    const double2 wb = *reinterpret_cast<const double2*>(&sW->w0c[c4][f][2]);
                    pre[0] = fma(m0d[f], wa.x, pre[0]); pre[1] = fma(m0d[f], wa.y, pre[1]);
                    pre[2] = fma(m0d[f], wb.x, pre[2]); pre[3] = fma(m0d[f], wb.y, pre[3]);
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) softplus_sigmoid_f64w(pre[kk], hq[kk], sq[kk]);
                if (rv) {
                    *reinterpret_cast<float4*>(sSig0 + lane * 20 + 4 * c4) = make_float4(sq[0], sq[1], sq[2], sq[3]);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) sHbuf[lane * 17 + 4 * c4 + kk] = hq[kk];
                }
            }
            __syncwarp();
            {   // hbar_c = mean_k h_k[c]  (lane c), then gb_c = sum_e hbar_e W1b[e][c]
                const int c = lane & 15;
                double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
                for (int r = 0; r + 1 < N; r += 2) { a0 += sHbuf[r * 17 + c]; a1 += sHbuf[(r + 1) * 17 + c]; }
                if (N & 1) a0 += sHbuf[(N - 1) * 17 + c];
                if (lane < 16) sHbar[c] = (a0 + a1) / (double)N;
                __syncwarp();
                double gbv = 0.0;
                const double* w1b = p.params + 256 + 256 + c;       // W1[16+e][c]
#pragma unroll
                for (int e = 0; e < 16; ++e) gbv = fma(sHbar[e], __ldg(w1b + 16 * e), gbv);
                if (lane < 16) sGb[c] = gbv;
                __syncwarp();
            }
            double h[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) h[e] = sHbuf[kself * 17 + e];
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {
                float sq[4];
                double ga[4], gm[4];              // two accumulation chains per channel (h part, mean-t part)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) { ga[kk] = cW.db1[4 * c4 + kk] + sGb[4 * c4 + kk]; gm[kk] = 0.0; }
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const double2 a01 = *reinterpret_cast<const double2*>(&sW->w1a[c4][e][0]);
                    const double2 a23 = *reinterpret_cast<const double2*>(&sW->w1a[c4][e][2]);
                    const double2 c01 = *reinterpret_cast<const double2*>(&sW->w1c[c4][e][0]);
                    const double2 c23 = *reinterpret_cast<const double2*>(&sW->w1c[c4][e][2]);
                    ga[0] = fma(h[e], a01.x, ga[0]); ga[1] = fma(h[e], a01.y, ga[1]);
                    ga[2] = fma(h[e], a23.x, ga[2]); ga[3] = fma(h[e], a23.y, ga[3]);
                    gm[0] = fma(md[e], c01.x, gm[0]); gm[1] = fma(md[e], c01.y, gm[1]);
                    gm[2] = fma(md[e], c23.x, gm[2]); gm[3] = fma(md[e], c23.y, gm[3]);
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int c = 4 * c4 + kk;
                    double sp;
                    softplus_sigmoid_f64w(ga[kk] + gm[kk], sp, sq[kk]);
                    const double sv = sHbuf[kself * 17 + c] + sp;  // s_i (flow.py:52)
                    dzx = fma(sv, cW.dwf[c][0], dzx);
                    dzy = fma(sv, cW.dwf[c][1], dzy);
                }
                if (rv) *reinterpret_cast<float4*>(sS1 + lane * 20 + 4 * c4) = make_float4(sq[0], sq[1], sq[2], sq[3]);
            }
        }
        __syncwarp();

        PROF_MARK(2);
        // ---------------- pass 2b: H_i = d hbar / d x_i = (1/n) [ sum_k P_k E_ki - P_i sum_k E_ik ] ------------
        // P_k E_ki = sigma0_k . W0c^T (eps . E_ik), eps = (-,-,+,+,-), so H needs S[c][fd] = sum_{k!=i} sigma0_k[c] E_ik[fd] for
        // the 6 non-zero entries fd of E.  Four of them are cos/sin of theta_i - theta_k and SEPARATE:
        //   sum_k s0_k[c] sin(th_i - th_k) = sin th_i (sum_k s0_k[c] cos th_k) - cos th_i (sum_k s0_k[c] sin th_k),
        // i.e. O(n) work from 68 warp-wide sums (A/B per channel and direction, plus the plain sums for sum_k E_ik);
        // only the two d-feature entries (-alpha sin(th)/(4 d_ik)) need the pair loop: 32 FMA per pair instead of 96.
        {
            const float cthx = fmaf(-2.f * pos.z, pos.z, 1.f), sthx = 2.f * pos.z * pos.x;     // full angles 2 pi u
            const float cthy = fmaf(-2.f * pos.w, pos.w, 1.f), sthy = 2.f * pos.w * pos.y;
            float* const sTab = sScr;                 // [N][68] contributions, then [68] sums at row 0 (h is dead: region free)
            if (rv) {
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 s04 = *reinterpret_cast<const float4*>(sSig0 + lane * 20 + 4 * c4);
                    const float s0v[4] = {s04.x, s04.y, s04.z, s04.w};
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        *reinterpret_cast<float4*>(sTab + lane * 68 + 4 * (4 * c4 + kk)) =
                            make_float4(s0v[kk] * cthx, s0v[kk] * sthx, s0v[kk] * cthy, s0v[kk] * sthy);
                }
                *reinterpret_cast<float4*>(sTab + lane * 68 + 64) = make_float4(cthx, sthx, cthy, sthy);
            }
            __syncwarp();
            float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f;    // column sums: columns lane, lane+32, lane+64 (< 68)
            {
                const int c2 = lane + 64 < 68 ? lane + 64 : 64;
#pragma unroll 4
                for (int r = 0; r < N; ++r) {
                    cs0 += sTab[r * 68 + lane]; cs1 += sTab[r * 68 + 32 + lane]; cs2 += sTab[r * 68 + c2];
                }
            }
            __syncwarp();
            float* const sSum = sTab + N * 68 - 68;   // last row slot is rewritten with the sums (row N-1 already consumed)
            sSum[lane] = cs0; sSum[32 + lane] = cs1;
            if (lane < 4) sSum[64 + lane] = cs2;
            __syncwarp();

            float S4[16][2];                          // pairwise part: fd = 4x, 4y
#pragma unroll
            for (int c = 0; c < 16; ++c) { S4[c][0] = 0.f; S4[c][1] = 0.f; }
            float sE4x = 0.f, sE4y = 0.f;
            float4 pn = sPos[0];
#pragma unroll 2
            for (int k = 0; k < N; ++k) {
                const float4 pj = pn;
                pn = sPos[k + 1 < N ? k + 1 : k];
                {   // branch-free: the diagonal partner (sh = 0: e4 = 0 * finite) contributes exact zeros
                    const WHalf hd = whalf_diff(pos, npc, pj);
                    const float d2s = k != kself ? fmaf(hd.sh.x, hd.sh.x, hd.sh.y * hd.sh.y) : 1.f;
                    const float ai4d = (-0.5f * alpha) * rsqrt_approx(d2s);                                         // -alpha / (4 d) * 2
                    const float2 e4 = __fmul2_rn(__fmul2_rn(hd.sh, hd.ch), make_float2(ai4d, ai4d));                  // -alpha sin(th)/(4 d)
                    const float e4x = e4.x, e4y = e4.y;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 sk4 = *reinterpret_cast<const float4*>(sSig0 + k * 20 + 4 * c4);
                        const float skv[4] = {sk4.x, sk4.y, sk4.z, sk4.w};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const float2 s4 = ffma2s(skv[kk], e4, make_float2(S4[4 * c4 + kk][0], S4[4 * c4 + kk][1]));
                            S4[4 * c4 + kk][0] = s4.x; S4[4 * c4 + kk][1] = s4.y;
                        }
                    }
                    sE4x += e4x; sE4y += e4y;
                }
            }
            if (rv) {
                const float4 tot = *reinterpret_cast<const float4*>(sSum + 64);            // sum_k (cos thx, sin thx, cos thy, sin thy)
                // sum_{k != i} E_ik:  (0x, 2x, 1y, 3y); the k = i term of cos(0) = 1 is taken out
                const float sE0 = alpha * (sthx * tot.x - cthx * tot.y), sE1 = alpha - alpha * (cthx * tot.x + sthx * tot.y);
                const float sE3 = alpha * (sthy * tot.z - cthy * tot.w), sE4 = alpha - alpha * (cthy * tot.z + sthy * tot.w);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 s04 = *reinterpret_cast<const float4*>(sSig0 + lane * 20 + 4 * c4);
                    const float s0v[4] = {s04.x, s04.y, s04.z, s04.w};
                    float hv[8];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int c = 4 * c4 + kk;
                        const float4 ab = *reinterpret_cast<const float4*>(sSum + 4 * c);     // (Ax, Bx, Ay, By)[c]
                        const float S0 = alpha * (sthx * ab.x - cthx * ab.y);
                        const float S1 = alpha * s0v[kk] - alpha * (cthx * ab.x + sthx * ab.y);
                        const float S3 = alpha * (sthy * ab.z - cthy * ab.w);
                        const float S4v = alpha * s0v[kk] - alpha * (cthy * ab.z + sthy * ab.w);
                        const float hx = fmaf(cW.w0c[2][c], S1, -fmaf(cW.w0c[0][c], S0, cW.w0c[4][c] * S4[c][0]));
                        const float hy = fmaf(cW.w0c[3][c], S4v, -fmaf(cW.w0c[1][c], S3, cW.w0c[4][c] * S4[c][1]));
                        const float ox = fmaf(cW.w0c[0][c], sE0, fmaf(cW.w0c[2][c], sE1, cW.w0c[4][c] * sE4x));
                        const float oy = fmaf(cW.w0c[1][c], sE3, fmaf(cW.w0c[3][c], sE4, cW.w0c[4][c] * sE4y));
                        hv[2 * kk] = (hx - s0v[kk] * ox) * invn;
                        hv[2 * kk + 1] = (hy - s0v[kk] * oy) * invn;
                    }
                    *reinterpret_cast<float4*>(sH + lane * 32 + 8 * c4) = make_float4(hv[0], hv[1], hv[2], hv[3]);
                    *reinterpret_cast<float4*>(sH + lane * 32 + 8 * c4 + 4) = make_float4(hv[4], hv[5], hv[6], hv[7]);
                }
            }
        }
        __syncwarp();

        PROF_MARK(3);
        // ---------------- Jacobian-side per-row contractions (FP32), rolled over groups of 4 input channels ----
        //   WfS[c][a] = Wf[c][a] sigma1_i[c];  R = WfS W1c^T, T = WfS W1b^T, M = Wf^T + WfS W1a^T, Q = (M . sigma0) W0c^T
        float Q[2][5];
        {
            float wfs[16][2];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float4 s14 = *reinterpret_cast<const float4*>(sS1 + kself * 20 + 4 * c4);
                const float s1v[4] = {s14.x, s14.y, s14.z, s14.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    wfs[4 * c4 + kk][0] = cW.wf[4 * c4 + kk][0] * s1v[kk];
                    wfs[4 * c4 + kk][1] = cW.wf[4 * c4 + kk][1] * s1v[kk];
                }
            }
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int f = 0; f < 5; ++f) Q[a][f] = 0.f;
            __syncwarp();                    // every lane has read h (above) before the R|T rows overlay it
#pragma unroll 1
            for (int e4 = 0; e4 < 4; ++e4) {
                const float4 s04 = *reinterpret_cast<const float4*>(sSig0 + kself * 20 + 4 * e4);
                const float s0v[4] = {s04.x, s04.y, s04.z, s04.w};
                float r0v[4], r1v[4], t0v[4], t1v[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const int e = 4 * e4 + kk;
                    float2 r01 = make_float2(0.f, 0.f), t01 = make_float2(0.f, 0.f), mm01 = make_float2(cW.wf[e][0], cW.wf[e][1]);
#pragma unroll
                    for (int c = 0; c < 16; ++c) {                  // packed over the two rows a = 0, 1 of WfS
                        const float2 wv = make_float2(wfs[c][0], wfs[c][1]);
                        r01 = ffma2s(cW.w1c[e][c], wv, r01);
                        t01 = ffma2s(cW.w1b[e][c], wv, t01);
                        mm01 = ffma2s(cW.w1a[e][c], wv, mm01);
                    }
                    float r0 = r01.x, r1 = r01.y, t0 = t01.x, t1 = t01.y, mm0 = mm01.x, mm1 = mm01.y;
                    r0v[kk] = r0 * kLn2; r1v[kk] = r1 * kLn2;     // ln2: the pair pass uses Wt*log2e
                    t0v[kk] = t0; t1v[kk] = t1;
                    mm0 *= s0v[kk]; mm1 *= s0v[kk];
#pragma unroll
                    for (int f = 0; f < 5; ++f) {
                        Q[0][f] = fmaf(mm0, cW.w0c[f][e], Q[0][f]);
                        Q[1][f] = fmaf(mm1, cW.w0c[f][e], Q[1][f]);
                    }
                }
                if (rv) {
#if CGBW_UFORM
                    float* sc = sScr + lane * 68 + 8 * e4;       // [(R0,R1) x 16 | (T0,T1) x 16]: pairs over the two rows of the block
                    *reinterpret_cast<float4*>(sc) = make_float4(r0v[0], r1v[0], r0v[1], r1v[1]);
                    *reinterpret_cast<float4*>(sc + 4) = make_float4(r0v[2], r1v[2], r0v[3], r1v[3]);
                    *reinterpret_cast<float4*>(sc + 32) = make_float4(t0v[0], t1v[0], t0v[1], t1v[1]);
                    *reinterpret_cast<float4*>(sc + 36) = make_float4(t0v[2], t1v[2], t0v[3], t1v[3]);
#else
                    float* sc = sScr + lane * 68 + 4 * e4;       // [R0 | R1 | T0 | T1] x 16
                    *reinterpret_cast<float4*>(sc) = make_float4(r0v[0], r0v[1], r0v[2], r0v[3]);
                    *reinterpret_cast<float4*>(sc + 16) = make_float4(r1v[0], r1v[1], r1v[2], r1v[3]);
                    *reinterpret_cast<float4*>(sc + 32) = make_float4(t0v[0], t0v[1], t0v[2], t0v[3]);
                    *reinterpret_cast<float4*>(sc + 48) = make_float4(t1v[0], t1v[1], t1v[2], t1v[3]);
#endif
                }
            }
        }

        PROF_MARK(4);
        // ---------------- pass 2 + 3: Jacobian blocks  J_ik = Q_i E_ik + R_i F_ik + T_i H_k  -> staging -----------
#if CGBW_UFORM
        // R_i F_ik = R_i diag(sigma'_ik) Wt^T E_ik: the 2x5 matrix  U_ik = Q_i + sum_c (sigma'_c R_i[:,c]) Wt[:,c]^T  is accumulated
        // over the channels (5 packed FMAs per channel on the pair of block rows) and contracted with the 6 non-zeros of E_ik
        // ONCE per pair, instead of forming Wt^T E_ik per channel (10 -> 6 instructions per pair-channel).
        {
            float2 R2[16], T2[16], Q2[5];
            __syncwarp();                    // padding lanes (>= N) mirror row N-1, written by lane N-1 just above
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float* sc = sScr + kself * 68 + 8 * c4;
                const float4 a0 = *reinterpret_cast<const float4*>(sc), a1 = *reinterpret_cast<const float4*>(sc + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(sc + 32), b1 = *reinterpret_cast<const float4*>(sc + 36);
                R2[4 * c4] = make_float2(a0.x, a0.y); R2[4 * c4 + 1] = make_float2(a0.z, a0.w);
                R2[4 * c4 + 2] = make_float2(a1.x, a1.y); R2[4 * c4 + 3] = make_float2(a1.z, a1.w);
                T2[4 * c4] = make_float2(b0.x, b0.y); T2[4 * c4 + 1] = make_float2(b0.z, b0.w);
                T2[4 * c4 + 2] = make_float2(b1.x, b1.y); T2[4 * c4 + 3] = make_float2(b1.z, b1.w);
            }
#pragma unroll
            for (int f = 0; f < 5; ++f) Q2[f] = make_float2(Q[0][f], Q[1][f]);
            __syncwarp();                    // all R|T rows are in registers before the J blocks overwrite the region
            float2 rsx = make_float2(0.f, 0.f), rsy = make_float2(0.f, 0.f);
            float4 pn = sPos[0];
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
                const float4 pj = pn;
                pn = sPos[k + 1 < N ? k + 1 : k];
                {   // no branch around the diagonal partner: its lane computes a harmless dummy block that is neither stored nor summed
                    const bool offd = k != kself;
                    const WPairFeat f = wpair_features(pos, npc, pj);
                    const float d2s = offd ? f.d2 : 1.f;
                    const float i4d = 0.25f * rsqrt_approx(d2s);
                    // E_ik = (1/n) d t0_ik / d x_k   (5x2, 6 non-zeros)
                    const float2 e01 = __fmul2_rn(f.s, make_float2(alpha, alpha)), e23 = __fmul2_rn(f.c, make_float2(-alpha, -alpha));
                    const float2 e4 = __fmul2_rn(e01, make_float2(-i4d, -i4d));
                    const float d = d2s * (4.f * i4d);
                    float2 U[5];
#pragma unroll
                    for (int g = 0; g < 5; ++g) U[g] = Q2[g];
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) {
                        const int c = 2 * c2;
                        const float2 a2 = ffma2s(d, make_float2(cW.wt[4][c], cW.wt[4][c + 1]),
                                          ffma2s(f.s.y, make_float2(cW.wt[3][c], cW.wt[3][c + 1]),
                                          ffma2s(f.s.x, make_float2(cW.wt[2][c], cW.wt[2][c + 1]),
                                          ffma2s(f.c.y, make_float2(cW.wt[1][c], cW.wt[1][c + 1]),
                                          ffma2s(f.c.x, make_float2(cW.wt[0][c], cW.wt[0][c + 1]), make_float2(cW.bt[c], cW.bt[c + 1]))))));
                        const float sg0 = sigmoid2(a2.x), sg1 = sigmoid2(a2.y);
                        const float2 sr0 = __fmul2_rn(R2[c], make_float2(sg0, sg0)), sr1 = __fmul2_rn(R2[c + 1], make_float2(sg1, sg1));
#pragma unroll
                        for (int g = 0; g < 5; ++g) U[g] = ffma2s(cW.wt[g][c + 1], sr1, ffma2s(cW.wt[g][c], sr0, U[g]));
                    }
                    // J_ik (rows a = 0, 1 packed): column x from features 0, 2, 4, column y from 1, 3, 4
                    float2 jx = ffma2s(e01.x, U[0], ffma2s(e23.x, U[2], __fmul2_rn(U[4], make_float2(e4.x, e4.x))));
                    float2 jy = ffma2s(e01.y, U[1], ffma2s(e23.y, U[3], __fmul2_rn(U[4], make_float2(e4.y, e4.y))));
                    // mean-field coupling through hbar:  T_i H_k
                    const float4* hk = reinterpret_cast<const float4*>(sH + k * 32);
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) {
                        const float4 hv = hk[c2];                  // (Hx, Hy)[2 c2], (Hx, Hy)[2 c2 + 1]
                        jx = ffma2s(hv.x, T2[2 * c2], ffma2s(hv.z, T2[2 * c2 + 1], jx));
                        jy = ffma2s(hv.y, T2[2 * c2], ffma2s(hv.w, T2[2 * c2 + 1], jy));
                    }
                    if (rv && offd) sJ[lane * PB + k] = make_float4(jx.x, jy.x, jx.y, jy.y);
                    const float2 z2 = make_float2(0.f, 0.f);
                    rsx = __fadd2_rn(rsx, offd ? jx : z2); rsy = __fadd2_rn(rsy, offd ? jy : z2);
                }
            }
            // translation equivariance: sum_k J_ik = I  (reference tests/test_flow.py:29-32)
            if (rv) sJ[lane * PB + lane] = make_float4(1.f - rsx.x, -rsy.x, -rsx.y, 1.f - rsy.y);
        }
#else
        {
            float R[2][16], T[2][16];
            __syncwarp();                    // padding lanes (>= N) mirror row N-1, written by lane N-1 just above
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                const float* sc = sScr + kself * 68 + 4 * c4;
                const float4 a0 = *reinterpret_cast<const float4*>(sc), a1 = *reinterpret_cast<const float4*>(sc + 16);
                const float4 b0 = *reinterpret_cast<const float4*>(sc + 32), b1 = *reinterpret_cast<const float4*>(sc + 48);
                R[0][4 * c4] = a0.x; R[0][4 * c4 + 1] = a0.y; R[0][4 * c4 + 2] = a0.z; R[0][4 * c4 + 3] = a0.w;
                R[1][4 * c4] = a1.x; R[1][4 * c4 + 1] = a1.y; R[1][4 * c4 + 2] = a1.z; R[1][4 * c4 + 3] = a1.w;
                T[0][4 * c4] = b0.x; T[0][4 * c4 + 1] = b0.y; T[0][4 * c4 + 2] = b0.z; T[0][4 * c4 + 3] = b0.w;
                T[1][4 * c4] = b1.x; T[1][4 * c4 + 1] = b1.y; T[1][4 * c4 + 2] = b1.z; T[1][4 * c4 + 3] = b1.w;
            }
            __syncwarp();                    // all R|T rows are in registers before the J blocks overwrite the region
            float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
            float4 pn = sPos[0];
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
                const float4 pj = pn;
                pn = sPos[k + 1 < N ? k + 1 : k];
                if (k != kself) {
                    const WPairFeat f = wpair_features(pos, npc, pj);
                    const float i4d = 0.25f * rsqrt_approx(f.d2);
                    // E_ik = (1/n) d t0_ik / d x_k   (5x2, 6 non-zeros)
                    const float2 e01 = __fmul2_rn(f.s, make_float2(alpha, alpha)), e23 = __fmul2_rn(f.c, make_float2(-alpha, -alpha));
                    const float2 e4 = __fmul2_rn(e01, make_float2(-i4d, -i4d));
                    const float e0x = e01.x, e1y = e01.y, e2x = e23.x, e3y = e23.y, e4x = e4.x, e4y = e4.y;
                    const float d = f.d2 * (4.f * i4d);
                    float2 jt = __ffma2_rn(make_float2(Q[0][0], Q[0][1]), e01,
                                           __ffma2_rn(make_float2(Q[0][2], Q[0][3]), e23, __fmul2_rn(make_float2(Q[0][4], Q[0][4]), e4)));
                    float2 jb = __ffma2_rn(make_float2(Q[1][0], Q[1][1]), e01,
                                           __ffma2_rn(make_float2(Q[1][2], Q[1][3]), e23, __fmul2_rn(make_float2(Q[1][4], Q[1][4]), e4)));
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
#if CGBW_PACK_PRE23
                        float a;
                        {
                            const int ce = c & ~1;
                            const float2 a2 = ffma2s(d, make_float2(cW.wt[4][ce], cW.wt[4][ce + 1]),
                                              ffma2s(f.s.y, make_float2(cW.wt[3][ce], cW.wt[3][ce + 1]),
                                              ffma2s(f.s.x, make_float2(cW.wt[2][ce], cW.wt[2][ce + 1]),
                                              ffma2s(f.c.y, make_float2(cW.wt[1][ce], cW.wt[1][ce + 1]),
                                              ffma2s(f.c.x, make_float2(cW.wt[0][ce], cW.wt[0][ce + 1]), make_float2(cW.bt[ce], cW.bt[ce + 1]))))));
                            a = (c & 1) ? a2.y : a2.x;        // the pair is computed once: the compiler merges the two identical expressions
                        }
#else
                        const float a = fmaf(d, cW.wt[4][c], fmaf(f.s.y, cW.wt[3][c], fmaf(f.s.x, cW.wt[2][c],
                                        fmaf(f.c.y, cW.wt[1][c], fmaf(f.c.x, cW.wt[0][c], cW.bt[c])))));
#endif
                        const float sg = sigmoid2(a);
                        // F_ik[c,:] = sigma_t * Wt^T E_ik   (Wt carries log2e; R carries ln2)
                        const float fx = sg * fmaf(cW.wt[0][c], e0x, fmaf(cW.wt[2][c], e2x, cW.wt[4][c] * e4x));
                        const float fy = sg * fmaf(cW.wt[1][c], e1y, fmaf(cW.wt[3][c], e3y, cW.wt[4][c] * e4y));
                        // J rows += R_i[c] (fx, fy), packed over (x, y).  (Packing g = Wt^T E as well was measured slower:
                        // the weight pairs need uniform-register loads and the compiler hoists all 16 g's, spilling.)
                        const float2 fxy = make_float2(fx, fy);
                        jt = ffma2s(R[0][c], fxy, jt);
                        jb = ffma2s(R[1][c], fxy, jb);
                    }
                    float j00 = jt.x, j01 = jt.y, j10 = jb.x, j11 = jb.y;
                    // mean-field coupling through hbar:  T_i H_k
                    const float4* hk = reinterpret_cast<const float4*>(sH + k * 32);
#pragma unroll
                    for (int c2 = 0; c2 < 8; ++c2) {
                        const float4 hv = hk[c2];
                        const float2 ha = make_float2(hv.x, hv.y), hb = make_float2(hv.z, hv.w);
                        const float2 jt = ffma2s(T[0][2 * c2], ha, ffma2s(T[0][2 * c2 + 1], hb, make_float2(j00, j01)));
                        const float2 jb = ffma2s(T[1][2 * c2], ha, ffma2s(T[1][2 * c2 + 1], hb, make_float2(j10, j11)));
                        j00 = jt.x; j01 = jt.y; j10 = jb.x; j11 = jb.y;
                    }
                    if (rv) sJ[lane * PB + k] = make_float4(j00, j01, j10, j11);
                    rs0 += j00; rs1 += j01; rs2 += j10; rs3 += j11;
                }
            }
            // translation equivariance: sum_k J_ik = I  (reference tests/test_flow.py:29-32)
            if (rv) sJ[lane * PB + lane] = make_float4(1.f - rs0, -rs1, -rs2, 1.f - rs3);
        }
#endif
        __syncwarp();
        if (p.out_jac && step == 0) {
            float* oj = p.out_jac + (size_t)b * (2 * N) * (2 * N);
            const float* Jf = reinterpret_cast<const float*>(sJ);
            for (int q = lane; q < 4 * N * N; q += 32) {
                const int rr = q / (2 * N), cc = q - rr * (2 * N);
                oj[q] = Jf[((rr >> 1) * PB + (cc >> 1)) * 4 + (rr & 1) * 2 + (cc & 1)];
            }
        }

        PROF_MARK(5);
        // ---------------- Jacobian LU: unpivoted 2x2-block LU, register-resident, rolled (shift-left) --------------
        // Step k: every row eliminates the active block column (register 0) with the published pivot line k and
        // -P_kk^{-1}, writing the updated block j into register j-1.  Right after the first update the new diagonal
        // block of row k+1 is broadcast by shuffle and EVERY lane inverts it (no divergent branch: the reciprocal
        // chain interleaves with the remaining updates instead of serialising the warp); row k+1 then publishes
        // its row as line k+1.  Rows <= k compute garbage nobody reads.  log|det J| = sum_k log|det P_kk|.
        float lacc = 0.f;                         // sum_k log2|det P_kk|, identical in every lane
        bool lacc_per_lane = false;               // pivoted fallback: per-lane partial sums instead
        {
            constexpr int LP = N + 8;
            float4 Jr[N];
#pragma unroll
            for (int j = 0; j < N; ++j) Jr[j] = sJ[kself * PB + j];
            bool bad = false;
            unsigned lu_chk = 0u;
            float4* const line0 = reinterpret_cast<float4*>(sm + L_::o_H);      // [2][LP]: blocks 1..N-1 (+ prefetch padding)
            auto neg_inv = [&](const float4 pk) -> float4 {
                const float det = pk.x * pk.w - pk.y * pk.z;
                const float idet = rcp_approx(det);
                lacc += lg2_approx(fabsf(det));
                bad |= !(fabsf(det) > 0.05f);                               // also catches NaN
                return make_float4(-pk.w * idet, pk.y * idet, pk.z * idet, -pk.x * idet);   // -P^{-1}
            };
            float4 m = neg_inv(shfl4w(Jr[0], 0));
            if (lane == 0) {
#pragma unroll
                for (int j = 1; j < N; ++j) line0[j] = Jr[j];
            }
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < N - 1; ++k) {
                const float4* cur = line0 + (k & 1) * LP;
                float4* nxt = line0 + ((k + 1) & 1) * LP;
                const unsigned cur32 = (unsigned)__cvta_generic_to_shared(cur);
                const int lim = N - 1 - k;                                    // live trailing blocks: registers 1..lim
                const bool act = rv && lane > k;
                const float4 a = Jr[0];
                const float l00 = fmaf(a.x, m.x, a.y * m.z), l01 = fmaf(a.x, m.y, a.y * m.w);
                const float l10 = fmaf(a.z, m.x, a.w * m.z), l11 = fmaf(a.z, m.y, a.w * m.w);
                bad |= act && fmaxf(fmaxf(fabsf(l00), fabsf(l01)), fmaxf(fabsf(l10), fabsf(l11))) > 2.f;
                const unsigned nxt32 = (unsigned)__cvta_generic_to_shared(nxt);
                const bool pub = lane == k + 1;                               // the next pivot row publishes its row as it is updated
                float4 pq[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) pq[t] = lds_pinned_f4(cur32 + 16u * (unsigned)(1 + t));
#pragma unroll
                for (int j0 = 1; j0 < N; j0 += 4) {
                    if (j0 > lim) break;                                      // warp-uniform; at most three dead blocks of work
                    float4 nq[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t)                               // next chunk ahead of its use (line is padded)
                        nq[t] = lds_pinned_f4(cur32 + 16u * (unsigned)(j0 + 4 + t));
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int j = j0 + t;
                        if (j < N) {
                            const float4 pv = pq[t];
                            float4 v = Jr[j];
                            const float2 pa = make_float2(pv.x, pv.y), pb = make_float2(pv.z, pv.w);
                            const float2 vt = ffma2s(l00, pa, ffma2s(l01, pb, make_float2(v.x, v.y)));
                            const float2 vb = ffma2s(l10, pa, ffma2s(l11, pb, make_float2(v.z, v.w)));
                            v = make_float4(vt.x, vt.y, vb.x, vb.y);
                            Jr[j - 1] = v;
                            if (j == 1) m = neg_inv(shfl4w(v, k + 1));        // next pivot block, inverted by every lane
                            else sts_pred_f4(nxt32 + 16u * (unsigned)(j - 1), v, pub);
                        }
                    }
                    lu_chk ^= __float_as_uint(nq[3].w) & lanebits;   // a (per-lane) use inside this block keeps the prefetch where it is written
#pragma unroll
                    for (int t = 0; t < 4; ++t) pq[t] = nq[t];
                }
                __syncwarp();
            }
            bad = __any_sync(FULL, bad && rv) || (p.force_pivot != 0) || lu_chk == 0x7fc0dead;    // (lu_chk: never true)
            if (bad) {                              // rare: redo with the partially pivoted LU on the staged copy
                lacc = pivoted_lu_smem<N>(reinterpret_cast<float*>(sJ), lane);     // all 32 lanes own scalar rows here
                lacc_per_lane = true;
                if (lane == 0 && p.fallback_counter) atomicAdd(p.fallback_counter, 1ull);
            }
        }
        __syncwarp();

        PROF_MARK(6);
        // ---------------- Slater determinant: D_rj = exp(2 pi i m_j . w_r), w = u + dz/L  (src/slater.py:14-19) ---------------
        // The twist and any common displacement are pure row/column phases and drop out of |det|.  Lane r builds row r from
        // per-particle power tables X_r^m, Y_r^m, m = 0..mmax (one float64 sincospi per coordinate, float64 recurrence).
        // Fast path: the tables are rounded to FP32 and the matrix is factored by a partially pivoted complex FP32 LU
        // (register-resident, shift-left, implicit row exchange).  An a-posteriori bound of its rounding error,
        //   err^2 ~ eps^2 sum_k k / |u_kk|^2   (k accumulated products in pivot k, amplified by 1 / |pivot|),
        // is accumulated on the fly; evaluations that exceed the budget (near-singular D: small pivots) or that use
        // wave vectors beyond the tables are redone by the float64 LU (slater_logdet_f64, same tables in float64).
        double aD = 0.0;
        const double wx = fma(dzx, invL, ux), wy = fma(dzy, invL, uy);
        bool need64 = p.force_f64_slater != 0 || mmax > L_::MMAX_TAB;
        if (!need64) {
            const int PT = (2 * (mmax + 1)) | 1;
            float2* const sPf = reinterpret_cast<float2*>(sm + L_::o_J);
            float2* const tb = sPf + kself * PT;
            if (rv) {
#pragma unroll 1
                for (int comp = 0; comp < 2; ++comp) {
                    const double wv = comp ? wy : wx;
                    double sn, cs;
                    sincospi(2.0 * (wv - rint(wv)), &sn, &cs);
                    float2* t = tb + comp * (mmax + 1);
                    double2 pw = make_double2(1.0, 0.0);
                    t[0] = make_float2(1.f, 0.f);
                    for (int m = 1; m <= mmax; ++m) {
                        pw = make_double2(pw.x * cs - pw.y * sn, pw.x * sn + pw.y * cs);
                        t[m] = make_float2((float)pw.x, (float)pw.y);
                    }
                }
            }
            __syncwarp();
            float2 Dr[N];
            {
                const float2* ty = tb + (mmax + 1);
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    const int2 mk = sK[j];
                    float2 X = tb[abs(mk.x)], Y = ty[abs(mk.y)];
                    X.y = mk.x < 0 ? -X.y : X.y;
                    Y.y = mk.y < 0 ? -Y.y : Y.y;
                    Dr[j] = make_float2(fmaf(X.x, Y.x, -X.y * Y.y), fmaf(X.x, Y.y, X.y * Y.x));
                }
            }
            __syncwarp();     // H region (lines) free, tables read
            PROF_MARK(7);
            // ---- complex FP32 LU.  Step k: every unfinished row offers |entry|^2 of the active column (register 0); ONE
            // REDUX picks the pivot row.  Every lane has already computed the reciprocal of ITS OWN candidate (no
            // divergence, and the MUFU runs beside the REDUX instead of after it), so the winner's 1/pivot arrives by two
            // shuffles.  The update  v += l p  is two packed FMAs per entry:  (vx, vy) += px (lx, ly) + py (-ly, lx);
            // the pivot line travels through a double-buffered shared-memory line as plain (x, y) pairs, loaded one
            // chunk of 8 entries ahead of its use.
            constexpr int LP = (N + 18) & ~1;                   // float2 slots per line buffer: entry j sits in slot j + 1 (entry 1 is
                                                                // 16-byte aligned), + prefetch padding
            float2* const fline0 = reinterpret_cast<float2*>(sm + L_::o_H);     // [2][LP]
            bool done = !rv, singular = false;
            float mypn = 1.f, myerr = 0.f;
            unsigned chk = 0u;
            const float fone = p.B > 0 ? 1.f : 2.f;             // 1, opaque to the compiler
            float pn, rinv;
            float2 own;
            auto offer = [&](const float2 e) -> unsigned {
                pn = fmaf(e.x, e.x, e.y * e.y);
                rinv = rcp_approx(pn);
                own = make_float2(e.x * rinv, -e.y * rinv);     // 1 / e
                const unsigned key = done ? 0u : ((__float_as_uint(pn) & ~0x1Fu) | (unsigned)lane);
                return __reduce_max_sync(FULL, key);
            };
            unsigned wk = offer(Dr[0]);
            float2 iv = make_float2(__shfl_sync(FULL, own.x, (int)(wk & 31u)), __shfl_sync(FULL, own.y, (int)(wk & 31u)));
            bool win = !done && lane == (int)(wk & 31u) && (wk >> 5) != 0u;
            if (win) {
                mypn = pn; myerr = rinv;
#pragma unroll
                for (int j = 1; j < N; ++j) fline0[j + 1] = Dr[j];
            }
            __syncwarp();
#pragma unroll 1
            for (int k = 0; k < N; ++k) {
                const float2* cur = fline0 + (k & 1) * LP;
                float2* nxt = fline0 + ((k + 1) & 1) * LP;
                const unsigned cur32 = (unsigned)__cvta_generic_to_shared(cur);
                const unsigned nxt32 = (unsigned)__cvta_generic_to_shared(nxt);
                const int lim = N - 1 - k;
                singular |= (wk >> 5) == 0u;
                done |= win;
                if (lim == 0) break;
                const float2 c = Dr[0];
                const float2 L1 = make_float2(-(c.x * iv.x - c.y * iv.y), -(c.x * iv.y + c.y * iv.x));   // l = -c / pivot
                // i l as a register pair of its own (the products by a run-time 1 keep ptxas from re-deriving it from L1 by operand
                // modifiers, which costs a pair copy per packed FMA)
                const float2 L2 = make_float2(-L1.y * fone, L1.x * fone);
                float4 pq[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) pq[t] = lds_pinned_f4(cur32 + 16u + 16u * (unsigned)t);       // entries 1..8
#pragma unroll
                for (int j0 = 1; j0 < N; j0 += 8) {
                    if (j0 > lim) break;                          // warp-uniform; at most seven dead entries of work
                    float4 nq[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) nq[t] = lds_pinned_f4(cur32 + 16u + 16u * (unsigned)(((j0 - 1) >> 1) + 4 + t));
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        const int j = j0 + t;
                        if (j < N) {
                            const float px = (t & 1) ? pq[t >> 1].z : pq[t >> 1].x, py = (t & 1) ? pq[t >> 1].w : pq[t >> 1].y;
                            const float2 v = ffma2s(px, L1, ffma2s(py, L2, Dr[j]));
                            Dr[j - 1] = v;
                            if (j == 1) {
                                wk = offer(v);
                                iv = make_float2(__shfl_sync(FULL, own.x, (int)(wk & 31u)), __shfl_sync(FULL, own.y, (int)(wk & 31u)));
                                win = !done && lane == (int)(wk & 31u) && (wk >> 5) != 0u;
                                if (win) { mypn = pn; myerr = (float)(k + 2) * rinv; }
                            }
                        }
                    }
                    // the next pivot row publishes this chunk of its row (predicated, no divergence): new register r (= j - 1) is
                    // entry r of the next line, slot r + 1; slots of odd r are 16-byte aligned: three pair stores + two singles
                    if (j0 > 1 && j0 - 1 < N - 1) sts_pred_f2(nxt32 + 8u * (unsigned)j0, Dr[j0 - 1], win);
#pragma unroll
                    for (int r = j0; r < j0 + 6; r += 2) {
                        if (r + 1 < N - 1) sts_pred_f2x2(nxt32 + 8u * (unsigned)(r + 1), Dr[r], Dr[r + 1], win);
                        else if (r < N - 1) sts_pred_f2(nxt32 + 8u * (unsigned)(r + 1), Dr[r], win);
                    }
                    if (j0 + 6 < N - 1) sts_pred_f2(nxt32 + 8u * (unsigned)(j0 + 7), Dr[j0 + 6], win);
                    chk ^= __float_as_uint(nq[3].w) & lanebits;       // a (per-lane) use inside this block keeps the prefetch where it is written
#pragma unroll
                    for (int t = 0; t < 4; ++t) pq[t] = nq[t];
                }
                __syncwarp();
            }
            // error bound of the FP32 factorisation (see above) and the decision to redo in float64
            float es = myerr;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) es += __shfl_xor_sync(FULL, es, o);
#ifdef CGBW_DIAG_ERRSUM                    // calibration build: never fall back on the indicator, export it instead
            need64 = singular || chk == 0x7fc0dead;
#else
            need64 = singular || !(es < p.slater32_errsum_max) || chk == 0x7fc0dead;    // (chk: never true, keeps the prefetch alive)
#endif
            aD = (rv && mypn != 1.f) ? 0.5 * log((double)mypn) : 0.0;
            aD = warp_sum_f64(aD);
#ifdef CGBW_DIAG_ERRSUM
            aJ_diag = (double)es;
#endif
        }
        if (need64) {
            aD = slater_logdet_f64<N>(sm, sK, lane, kself, rv, mmax, wx, wy);
            if (lane == 0 && p.fallback_counter) atomicAdd(p.fallback_counter + 1, 1ull);
        }
        PROF_MARK(8);
        aJ = lacc_per_lane ? warp_sum_f64((double)lacc * 0.693147180559945309417)
                           : (double)lacc * 0.693147180559945309417;
        ldD = aD - nlogL;
        lp = 2.0 * ldD + aJ;                                 // logpsi.py:30,179
        }   // evaluation

        if constexpr (SPEC) {
            // ---------------- exchange the nodes' log p, replay the decisions along the realised path ----------------
            const int w = (int)(threadIdx.x >> 5);
            double* const slot = sLp + (sp_round & 1) * 8;       // double-buffered: one block barrier per round
            if (lane == 0) slot[w] = lp;
            if (active && step == 0 && p.out_parts && lane == 0) { p.out_parts[2 * (size_t)b] = ldD; p.out_parts[2 * (size_t)b + 1] = aJ; }
            __syncthreads();
            const bool writer = w == 0 && lane == 0;
            const int extra = 2 * sp_nacc > sp_t ? 2 * SPECW - 1 : SPECW;      // the same expression the round started with
            const bool first = sp_t == 0;
            double cur = sp_logp;
            if (first) {
                cur = slot[SPECW - 1];
                if (p.out_logp_trace && writer) p.out_logp_trace[b] = cur;
            }
            int node = 1, adv = 0;
            while (sp_t + adv < p.mc_steps) {
                const int sl = node < SPECW ? node - 1 : ((!first && node == extra) ? SPECW - 1 : -1);
                if (sl < 0) break;                               // not evaluated in this round
                const int s = sp_t + adv + 1;
                const double lpn = slot[sl];
                const double u = uniform_of(b, s);
                const double ratio = exp(lpn - cur);             // MCMC.py:28
                const int acc = (u < ratio) ? 1 : 0;             // MCMC.py:29  (NaN -> reject)
                if (acc) { cur = lpn; ++sp_nacc; add_noise(sp_x, b, s); }      // MCMC.py:31
                if (writer) {
                    if (p.out_acc_trace) p.out_acc_trace[(size_t)(s - 1) * p.B + b] = (uint8_t)acc;
                    if (p.out_logp_trace) p.out_logp_trace[(size_t)s * p.B + b] = lpn;
                }
                node = 2 * node + acc;
                ++adv;
            }
            sp_logp = cur;
            sp_t += adv;
            ++sp_round;
            if (sp_t >= p.mc_steps) {                            // sweep of this walker finished
                if (w == 0) {
                    double2 xf = sp_x;
                    if (p.wrap) { xf.x -= p.L * floor(xf.x / p.L); xf.y -= p.L * floor(xf.y / p.L); }   // VMC.py:24
                    if (rv && p.mc_steps > 0) *reinterpret_cast<double2*>(p.x_out + ((size_t)b * N + lane) * 2) = xf;
                    if (lane == 0) {
                        if (p.out_logp) p.out_logp[b] = sp_logp;
                        if (p.out_naccept) p.out_naccept[b] = sp_nacc;
                        if (p.out_stats && p.mc_steps > 0) atomicAdd(p.out_stats + CGB_STATS_NACCEPT, (double)sp_nacc);
                    }
                }
                sp_b += gridDim.x;
                sp_load = true;
            }
        } else {
        // ---------------- Metropolis accept / reject (MCMC.py:28-33), warp-uniform ------------------------------
        int acc = 1;
        if (step == 0) {
            logp_cur = lp;
#ifdef CGBW_DIAG_ERRSUM
            aJ = aJ_diag;              // calibration build: the FP32 Slater LU's error indicator instead of log|det J|
#endif
            if (p.out_parts && lane == 0) { p.out_parts[2 * (size_t)b] = ldD; p.out_parts[2 * (size_t)b + 1] = aJ; }
        } else {
            const double u = uniform_of(b, step);
            const double ratio = exp(lp - logp_cur);         // MCMC.py:28
            acc = (u < ratio) ? 1 : 0;                       // MCMC.py:29  (NaN -> reject)
            if (acc) logp_cur = lp;
            if (p.out_acc_trace && lane == 0) p.out_acc_trace[(size_t)(step - 1) * p.B + b] = (uint8_t)acc;
        }
        if (p.out_logp_trace && lane == 0) p.out_logp_trace[(size_t)step * p.B + b] = lp;

        const bool last = step == p.mc_steps;
        if (p.mc_steps > 0) {
            double2* xo = reinterpret_cast<double2*>(p.x_out + ((size_t)b * N + kself) * 2);
            const bool changed = step == 0 ? (p.x_out != p.x) : (acc != 0);      // MCMC.py:31
            if (last) {
                double2 xf = xv;
                if (!changed && rv) xf = __ldcg(xo);
                if (p.wrap) { xf.x -= p.L * floor(xf.x / p.L); xf.y -= p.L * floor(xf.y / p.L); }   // VMC.py:24
                if (rv) *xo = xf;
            } else if (changed && rv) {
                *xo = xv;
            }
        }
        if (lane == 0) {
            int na = 0;
            if (step > 0) na = __ldcg(p.ws_nacc + b) + acc;
            if (last) {
                if (p.out_logp) p.out_logp[b] = logp_cur;
                if (p.out_naccept) p.out_naccept[b] = na;
                if (p.out_stats && p.mc_steps > 0) atomicAdd(p.out_stats + CGB_STATS_NACCEPT, (double)na);   // integers: exact, order-free
            } else {
                p.ws_logp[b] = logp_cur;
                p.ws_nacc[b] = na;
                p.done[b] = step + 1;
            }
        }
        if (!last) {                       // publish the walker state, then push the walker back as ready
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                const unsigned long long t = atomicAdd(p.qctr + 1, 1ull);
                unsigned long long* slot = p.queue + (t & qmask);
                while ((unsigned)(ld_acquire_u64(slot) >> 32) != (unsigned)t) __nanosleep(40);
                st_release_u64(slot, ((unsigned long long)(unsigned)(t + 1ull) << 32) | (unsigned long long)(unsigned)b);
            }
        }
        PROF_MARK(9);
        }
    }
#if CGBW_PROF
    if (lane == 0 && p.prof)
        for (int i = 0; i < 12; ++i) atomicAdd(p.prof + i, s_prof[threadIdx.x >> 5][i]);
#endif
}

// ---------------------------------------------------------------------------------------------------
// instantiated shapes: the closed shells of the 2D plane-wave lattice up to one warp (the reference's runs use 13, 29)
bool wsweep_supported(int n) { return n == 9 || n == 13 || n == 21 || n == 25 || n == 29; }

// launch configuration per (device, instantiation): function attribute + occupancy are queried once, the steady-state
// call is the launch alone
struct WLaunchCfg { int dev = -1; int wpc = 0; int occ = 0; size_t smem = 0; };

template <int N>
static int launch_wn(const SweepParams& p, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static thread_local int sms_cache[2] = {-1, 0};
    if (sms_cache[0] != dev) { cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); sms_cache[0] = dev; sms_cache[1] = sms; }
    sms = sms_cache[1];
    auto config = [&](auto kern, const int nw, const int nw_max, WLaunchCfg& c) -> int {
        if (c.dev == dev && c.wpc == nw) return 0;
        const size_t smem = (size_t)W_SHARED_BYTES + (size_t)WLay<N>::total * nw;
        // the attribute belongs to the FUNCTION, not to a launch shape: always the largest CTA this instantiation is launched with
        // (a smaller value set for a 2-warp launch would make the next 4-warp launch fail with cudaErrorInvalidValue)
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)((size_t)W_SHARED_BYTES + (size_t)WLay<N>::total * nw_max));
        if (e != cudaSuccess) return (int)e;
        int occ = 1;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, nw * 32, smem);
        c.dev = dev; c.wpc = nw; c.occ = occ < 1 ? 1 : occ; c.smem = smem;
        return 0;
    };
    if (p.spec) {          // speculative mode: one CTA of 8 (B <= #SMs) or 4 warps per walker, persistent over the sweep
        auto go = [&](auto kern, const int nw, WLaunchCfg& c) -> int {
            if (const int rc = config(kern, nw, nw, c)) return rc;
            long long grid = (long long)sms * c.occ;
            if (grid > (long long)p.B) grid = p.B;
            kern<<<(int)grid, nw * 32, c.smem, st>>>(p);
            return (int)cudaGetLastError();
        };
        static thread_local WLaunchCfg c8, c4;
        return p.spec == 8 ? go(wsweep_kernel<N, 8>, 8, c8) : go(wsweep_kernel<N, 4>, 4, c4);
    }
    // warps per CTA: 4 when the batch fills the machine, fewer for small batches so the walkers spread over all SMs
    int wpc = CGBW_WPC;
    while (wpc > 1 && (long long)p.B < (long long)sms * 3 * wpc) wpc >>= 1;
    static thread_local WLaunchCfg cq[CGBW_WPC + 1];
    WLaunchCfg& c = cq[wpc];
    if (const int rc = config(wsweep_kernel<N, 0>, wpc, CGBW_WPC, c)) return rc;
    SweepParams pq = p;
    pq.n_sms = (unsigned)sms;
    long long grid = (long long)sms * c.occ;
    const long long need = ((long long)p.B + wpc - 1) / wpc;      // never more warps than walkers
    if (grid > need) grid = need;
    wsweep_kernel<N, 0><<<(int)grid, wpc * 32, c.smem, st>>>(pq);
    return (int)cudaGetLastError();
}

// walkers the speculative mode keeps in flight at once (one CTA each); above this the ready-queue kernel is the better use
int wsweep_spec_capacity(int n) {
    (void)n;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms * 3;        // __launch_bounds__(128, 3): 168 registers per thread
}

int launch_wsweep(const SweepParams& p, void* wconst_staging, cudaStream_t st) {
    if (!wsweep_supported(p.n)) return CGB_E_UNSUPPORTED;
    if ((unsigned long long)p.B * (unsigned long long)(p.mc_steps + 2) >= (1ull << 31)) return CGB_E_UNSUPPORTED;
    if (p.wconst != wconst_staging) return CGB_E_BADARG;
    // the constant bank is per device: make sure it is the stream's device that gets refreshed and launched on (not asked of a
    // stream under graph capture: only launches and memcpy nodes are issued on those)
    int cur_dev = 0, st_dev = -1;
    cudaGetDevice(&cur_dev);
    cudaStreamCaptureStatus cap_st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cap_st) != cudaSuccess) { cap_st = cudaStreamCaptureStatusNone; (void)cudaGetLastError(); }
    if (cap_st == cudaStreamCaptureStatusNone && st != nullptr) {
        if (cudaStreamGetDevice(st, &st_dev) == cudaSuccess) { if (st_dev != cur_dev) return CGB_E_BADARG; }
        else (void)cudaGetLastError();
    }
    const unsigned cap = p.spec ? 0u : (unsigned)p.qcap;
    wsweep_prologue_kernel<<<cap ? (cap + 255) / 256 : 1, 256, 0, st>>>(p.params, reinterpret_cast<WConst*>(wconst_staging),
                                                                       const_cast<double*>(p.wshared), p.queue,
                                                                       cap, (unsigned)p.B, p.qctr, p.done, p.out_stats,
                                                                       (double)p.B * (double)p.mc_steps,
                                                                       const_cast<uint32_t*>(p.key_dev), p.advance_key);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    // Constant bank refresh, stream-ordered (device-to-device).  The block is per DEVICE: the refresh must not overtake a
    // sweep of another stream that still reads the previous content.  One event per device marks the last reader; a call on a
    // different stream waits for it before refreshing (sweeps of different streams are serialised on the device, each sees its
    // own parameters).  A stream under graph capture cannot take or give that dependency (include/cgb200.h).
    static std::mutex guard_mu;
    static struct { cudaEvent_t ev; cudaStream_t last; bool have; } guard[64] = {};
    const std::lock_guard<std::mutex> lock(guard_mu);
    const bool guarded = cap_st == cudaStreamCaptureStatusNone && cur_dev >= 0 && cur_dev < 64;
    if (guarded && guard[cur_dev].have && guard[cur_dev].last != st) {
        e = cudaStreamWaitEvent(st, guard[cur_dev].ev, 0);
        if (e != cudaSuccess) return (int)e;
    }
    e = cudaMemcpyToSymbolAsync(cW, wconst_staging, sizeof(WConst), 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return (int)e;
    int rc;
    switch (p.n) {
        case 9: rc = launch_wn<9>(p, st); break;
        case 13: rc = launch_wn<13>(p, st); break;
        case 21: rc = launch_wn<21>(p, st); break;
        case 25: rc = launch_wn<25>(p, st); break;
        case 29: rc = launch_wn<29>(p, st); break;
        default: rc = CGB_E_UNSUPPORTED;
    }
    if (guarded && rc == 0) {
        if (!guard[cur_dev].have) {
            if (cudaEventCreateWithFlags(&guard[cur_dev].ev, cudaEventDisableTiming) != cudaSuccess) return rc;
            guard[cur_dev].have = true;
        }
        cudaEventRecord(guard[cur_dev].ev, st);
        guard[cur_dev].last = st;
    }
    return rc;
}

}  // namespace cgb
