// Fused Metropolis sweep / batched log p kernel for sm_100a.
//
// One CTA per walker, persistent over all mc_steps of the sweep (src/MCMC.py:22-40): the walker's
// coordinates, log p and RNG state never leave the SM between proposals.  For every proposal the CTA
// evaluates  log p = 2 log|det D| + log|det J|  (src/logpsi.py:22-31,179):
//   flow forward (src/flow.py:39-55) -> Slater matrix (src/slater.py:14-17) -> complex LU, and the
//   analytic Jacobian of the flow (replacing jax.jacfwd, src/logpsi.py:27-28) -> real LU.
//
// Thread layout: 4 lanes per particle row ("row i, lane-in-row q"), 8 rows per warp, NW = ceil(n/8) warps.
// In the two O(n^2) pair passes lane q of row i handles partners j = q, q+4, ...; row sums are 2-stage
// shuffle butterflies inside the 4-lane group, so no cross-warp reduction buffers are needed.
// Weights live in shared memory (broadcast LDS.128); per-row contractions Q_i, R_i in registers.
// LU: both factorisations (real 2n x 2n Jacobian as unpivoted 2x2-block LU, complex n x n Slater matrix with
// partial pivoting) are register-resident and use all threads: thread (row, q) keeps columns q, q+4, ... of its
// row, the active column is broadcast inside the quad by shuffle, the pivot line travels through a
// double-buffered shared-memory line and there is ONE block barrier per elimination step (pivot search: one
// REDUX per warp on |a| bits with the row id in the low bits; every warp publishes its candidate row).
#include "cgb_common.cuh"
#include "cgb_internal.h"

namespace cgb {

// ---------------------------------------------------------------------------------------------------
// parameter packing: float64 haiku-order vector -> FP32 kernel layout (Wt, bt pre-scaled by log2 e)
// ---------------------------------------------------------------------------------------------------
__global__ void prep_params_kernel(const double* __restrict__ p, float* __restrict__ w, double* stats, double nprop,
                                   uint32_t* key_dev, int advance_key, int* seg_done, unsigned long long* ticket, int B) {
    for (int t = threadIdx.x; t < B; t += blockDim.x) seg_done[t] = 0;        // segmented sweeps (sweep_kernel, seg_len > 0)
    if (threadIdx.x == 0) {
        *ticket = 0ull;       // prologue duties of the sampling step (include/cgb200.h: CGB_STATS_*, CGB_FLAG_ADVANCE_KEY)
        if (stats) { for (int t = 0; t < CGB_STATS_LEN; ++t) stats[t] = 0.0; stats[CGB_STATS_NPROPOSED] = nprop; }
        if (key_dev && advance_key) {
            const uint64_t k = (((uint64_t)key_dev[0] << 32) | (uint64_t)key_dev[1]) + 0x9E3779B97F4A7C15ull;
            key_dev[0] = (uint32_t)(k >> 32); key_dev[1] = (uint32_t)k;
        }
    }
    // source offsets in the flat 1074 vector (include/cgb200.h)
    const int S_WT = 0, S_BT = 80, S_W0 = 96, S_B0 = 240, S_W1 = 256, S_B1 = 1024, S_WF = 1040;
    for (int t = threadIdx.x; t < W_TOTAL; t += blockDim.x) {
        float v;
        if (t < W_BT) v = (float)(p[S_WT + t] * 1.4426950408889634);
        else if (t < W_W0C) v = (float)(p[S_BT + (t - W_BT)] * 1.4426950408889634);
        else if (t < W_B0) v = (float)p[S_W0 + 4 * 16 + (t - W_W0C)];   // rows 4..8 of W0 (s0 == 0, flow.py:16-18)
        else if (t < W_W1A) v = (float)p[S_B0 + (t - W_B0)];
        else if (t < W_B1) v = (float)p[S_W1 + (t - W_W1A)];            // W1a | W1b | W1c contiguous
        else if (t < W_WF) v = (float)p[S_B1 + (t - W_B1)];
        else v = (float)p[S_WF + (t - W_WF)];                           // bf dropped: common shift of z
        w[t] = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// Shared-memory layout.  Everything whose size depends only on the row-slot count N8 = 8*NW sits at
// compile-time offsets; the two matrices (sizes depend on the run-time n) come last.
template <int N8>
struct Lay {
    static constexpr int o_xs = 0;                    // double[2*N8]   current walker
    static constexpr int o_xp = o_xs + 16 * N8;       // double[2*N8]   proposal
    static constexpr int o_ls = o_xp + 16 * N8;       // double[16]     log-det partial sums
    static constexpr int o_hb = o_ls + 128;           // double[16]     mean_k h_k
    static constexpr int o_us = o_hb + 128;           // double2[N8]    reduced coordinates u = frac(x/L)
    static constexpr int o_dz = o_us + 16 * N8;       // double2[N8]    backflow displacement s_i Wf (float64)
    static constexpr int o_w = o_dz + 16 * N8;        // float[W_TOTAL] FP32 weight pack
    static constexpr int o_k = o_w + 4 * W_TOTAL;     // int2[N8]       integer part of the occupied k rows
    static constexpr int o_pos = o_k + 8 * N8;        // float4[N8]     (cos,sin) of the half angles pi*u
    static constexpr int o_s0 = o_pos + 16 * N8;      // float[N8*16]   sigma0
    // [o_pos, o_red) = pos | sigma0 | H is dead once pass 3 is done: the Slater power tables alias it
    static constexpr int ptab_bytes = 16 * N8 + 64 * N8 + 144 * N8;
    static constexpr int o_H = o_s0 + 64 * N8;        // float[N8*36] H_k   (aliases double[N8*18] h_i: dead before H is written)
    static constexpr int o_red = o_H + 144 * N8;      // unsigned[2*8] J pivots, unsigned[2*8] D pivots, int[4] flags
    static constexpr int o_J = o_red + 160;           // float4[n*pitchB] Jacobian as 2x2 blocks; then R|T float[2*N8*36] (later the LU pivot lines)
    static_assert(o_J % 16 == 0, "alignment");
};
__host__ __device__ inline int pitch_B(int n) { return n | 1; }   // float4 blocks per block-row, odd: conflict-free LDS.128

size_t sweep_smem_bytes(int n) {
    const int n8 = 8 * sweep_num_warps(n);
    const size_t fixed = (size_t)(16 * n8 * 4 + 256 + 4 * W_TOTAL + 16 * n8 + 8 * n8 + 64 * n8 + 144 * n8 + 160);
    const size_t jb = sizeof(float4) * n * pitch_B(n);
    const size_t rt = sizeof(float) * n8 * 36 * 2;      // R|T rows; afterwards the LU pivot lines [2][NW][8 NW + 16] x 16 B
    const size_t nw = n8 / 8, ln = 2 * nw * (8 * nw + 16) * 16;
    return fixed + jb + (rt > ln ? rt : ln);
}

struct Smem {
    float* wts; double* xs; double* xp; float4* pos; double2* us; int2* kint; double2* ptab; double2* dz; float* sig0;
    double* hbuf; double* hbar; float* Rbuf; float* Tbuf; float* Hbuf; float4* J;
    unsigned* redJ; unsigned* redD; double* lsum; int* flag;
};

__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// float64 load from the same shared-memory offset in CTA `rank` of the cluster (distributed shared memory)
__device__ __forceinline__ double ld_dsmem_f64(const double* local, int rank) {
    unsigned a = (unsigned)__cvta_generic_to_shared(local), ra;
    double v;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(ra) : "memory");
    return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}

// packed FP32 FMA of sm_100 (FFMA2: two IEEE fmas per lane and instruction, scalar operand broadcast by the instruction):
// same results as two fmaf, half the issue slots
__device__ __forceinline__ float2 ffma2s(const float a, const float2 b, const float2 c) {
    return __ffma2_rn(make_float2(a, a), b, c);
}

// (a0, a1) += s * (x, y)  and  the 2x2 block update  v += L * q  of the Jacobian LU, packed where it pays: the 96-register
// instantiations (NW <= 4) spill more with register pairs than the packed instructions save (measured: n = 13, B = 512 4 % slower)
template <bool PK>
__device__ __forceinline__ void pair_fma(float& a0, float& a1, const float sc, const float x, const float y) {
    if constexpr (PK) {
        const float2 r = ffma2s(sc, make_float2(x, y), make_float2(a0, a1));
        a0 = r.x; a1 = r.y;
    } else {
        a0 = fmaf(sc, x, a0); a1 = fmaf(sc, y, a1);
    }
}
template <bool PK>
__device__ __forceinline__ void block_fma(float4& v, const float l00, const float l01, const float l10, const float l11, const float4 qv) {
    pair_fma<PK>(v.x, v.y, l01, qv.z, qv.w);
    pair_fma<PK>(v.x, v.y, l00, qv.x, qv.y);
    pair_fma<PK>(v.z, v.w, l11, qv.z, qv.w);
    pair_fma<PK>(v.z, v.w, l10, qv.x, qv.y);
}

__device__ __forceinline__ double quad_sum_f64(double v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    return v;
}
__device__ __forceinline__ int ld_acquire_s32(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_s32(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Kahan-compensated accumulation: the sums over partners feed the Slater phases, whose condition number
// amplifies FP32 summation error past the 1e-5 budget.
__device__ __forceinline__ void kahan_add(float& sum, float& comp, float x) {
    const float y = x - comp;
    const float t = sum + y;
    comp = (t - sum) - y;
    sum = t;
}
// float64 softplus (forward value) + FP32 logistic sigmoid (Jacobian side) of a float64 pre-activation
__device__ __forceinline__ void softplus_sigmoid_f64(double pre, double& sp, float& sg) {
    const double e = exp(-fabs(pre));
    sp = fmax(pre, 0.0) + log1p(e);
    const float ef = (float)e, r = rcp_approx(1.f + ef);
    sg = pre >= 0.0 ? r : ef * r;
}
__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// pair features of (row i, partner j) from the half-angle table: cos/sin of theta = 2 pi (x_i-x_j)/L and
// d^2 = sin^2(theta_x/2) + sin^2(theta_y/2)   (src/flow.py:22-25)
struct PairFeat { float cx, cy, sx, sy, d2; };
__device__ __forceinline__ PairFeat pair_features(const float4 pi_, const float4 pj) {
    const float shx = pi_.y * pj.x - pi_.x * pj.y, chx = fmaf(pi_.x, pj.x, pi_.y * pj.y);
    const float shy = pi_.w * pj.z - pi_.z * pj.w, chy = fmaf(pi_.z, pj.z, pi_.w * pj.w);
    PairFeat f;
    f.cx = fmaf(-2.f * shx, shx, 1.f);
    f.cy = fmaf(-2.f * shy, shy, 1.f);
    f.sx = 2.f * shx * chx;
    f.sy = 2.f * shy * chy;
    f.d2 = fmaf(shx, shx, shy * shy);
    return f;
}

template <int QQ>
__device__ __forceinline__ void store_quarter(float* dst, const float (&v)[32]) {
    // v is [c][d] (16x2); lane QQ stores channels 4QQ..4QQ+3 -> 8 floats
    float4* d4 = reinterpret_cast<float4*>(dst + 8 * QQ);
    d4[0] = make_float4(v[8 * QQ + 0], v[8 * QQ + 1], v[8 * QQ + 2], v[8 * QQ + 3]);
    d4[1] = make_float4(v[8 * QQ + 4], v[8 * QQ + 5], v[8 * QQ + 6], v[8 * QQ + 7]);
}

// ---------------------------------------------------------------------------------------------------
// SPECC = 4, 8: speculative ("prefetching") Metropolis for batches that leave SMs idle (SPECC = 0: one CTA per walker,
// step after step).  A thread-block CLUSTER of SPECC CTAs owns one walker; per round CTA r evaluates node r + 1 of the
// binary decision tree of the next steps (node 1: the next proposal; nodes 2 / 3: the proposal of the step after it given
// that the first one is rejected / accepted; ...), the last CTA one node of the level below (the all-rejected or the
// all-accepted path, whichever the running acceptance favours; in the first round: the initial walker).  Noise and
// uniforms are keyed by (walker, step), so every node's configuration is known up front.  One cluster barrier per round;
// then thread 0 of every CTA reads the nodes' log p through distributed shared memory and replays the decisions along
// the realised path (MCMC.py:28-33: same arithmetic, same order), every CTA applies the accepted moves to its replica of
// the walker, and the chain has advanced log2(SPECC) (+1) steps in one evaluation latency.  Bit-identical trajectories.
template <int NW, int SPECC>
__global__ void __launch_bounds__(NW * 32, (NW <= 2 ? 10 : (NW <= 4 ? 5 : 2)))
sweep_kernel(const SweepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(SPECC == 0 || SPECC == 4 || SPECC == 8, "cluster width of the speculative mode");
    constexpr bool PK = NW >= 5;             // packed FP32 FMAs (FFMA2) in the 128 / 168-register instantiations
    const int n = p.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = tid >> 2, q = tid & 3;
    const bool rv = row < n;
    const int i = rv ? row : n - 1;
    constexpr int n8 = 8 * NW;
    using L_ = Lay<n8>;
    static_assert(L_::o_J == 16 * n8 * 4 + 256 + 4 * W_TOTAL + 16 * n8 + 8 * n8 + 64 * n8 + 144 * n8 + 160, "host size formula");
    const int N2 = 2 * n, pB = pitch_B(n);
    Smem s;
    s.xs = (double*)(smem_raw + L_::o_xs); s.xp = (double*)(smem_raw + L_::o_xp); s.lsum = (double*)(smem_raw + L_::o_ls);
    s.hbar = (double*)(smem_raw + L_::o_hb); s.us = (double2*)(smem_raw + L_::o_us); s.dz = (double2*)(smem_raw + L_::o_dz);
    s.wts = (float*)(smem_raw + L_::o_w); s.pos = (float4*)(smem_raw + L_::o_pos); s.kint = (int2*)(smem_raw + L_::o_k); s.ptab = (double2*)(smem_raw + L_::o_pos);
    s.sig0 = (float*)(smem_raw + L_::o_s0); s.Hbuf = (float*)(smem_raw + L_::o_H); s.hbuf = (double*)(smem_raw + L_::o_H);
    s.redJ = (unsigned*)(smem_raw + L_::o_red); s.redD = s.redJ + 16; s.flag = (int*)(s.redJ + 32);
    s.J = (float4*)(smem_raw + L_::o_J);     // block (i,k) = (J[2i][2k], J[2i][2k+1], J[2i+1][2k], J[2i+1][2k+1])
    s.Rbuf = (float*)(s.J + n * pB); s.Tbuf = s.Rbuf + n8 * 36;
    constexpr int nJw = (NW + 1) / 2;  // warps of the pivoted fallback LU of the Jacobian (one thread per scalar row)
    const float invn = 1.f / (float)n;
    const double invL = 1.0 / p.L;
    const float alpha = (float)(2.0 * 3.14159265358979323846 * invL) * invn;   // (2 pi / L) / n

    {   // FP32 weight pack (4 kB): one TMA bulk copy per CTA
        __shared__ __align__(8) unsigned long long s_mbar;
        static_assert((4 * W_TOTAL) % 16 == 0, "bulk copies move multiples of 16 bytes");
        tma_stage_block(s.wts, p.wpack, 4 * W_TOTAL, &s_mbar);
    }
    const float4* wt4 = reinterpret_cast<const float4*>(s.wts + W_WT);     // [f*4+cq]
    const float4* bt4 = reinterpret_cast<const float4*>(s.wts + W_BT);
    const float4* w0c4 = reinterpret_cast<const float4*>(s.wts + W_W0C);   // [f*4+cq]
    const float4* b04 = reinterpret_cast<const float4*>(s.wts + W_B0);
    const float4* w1a4 = reinterpret_cast<const float4*>(s.wts + W_W1A);   // [row*4+cq]
    const float4* w1b4 = reinterpret_cast<const float4*>(s.wts + W_W1B);
    const float4* w1c4 = reinterpret_cast<const float4*>(s.wts + W_W1C);
    const float4* b14 = reinterpret_cast<const float4*>(s.wts + W_B1);
    const float* wf = s.wts + W_WF;                                        // [e*2+a]

    // proposal noise of (walker b, step st) for coordinate tid (MCMC.py:26): an input in parity mode, else Philox keyed by
    // (walker, step, particle) -- independent of who asks and when
    auto noise_of = [&](const int b, const int st) -> double {
        if (p.noise) return p.noise[((size_t)(st - 1) * p.B + b) * N2 + tid];
        const float2 g = rng_normal2(rng_seed_of(p.seed, p.key_dev), (uint32_t)b, (uint32_t)st, (uint32_t)(tid >> 1));
        return (double)((tid & 1) ? g.y : g.x);
    };
    auto uniform_of = [&](const int b, const int st) -> double {                 // MCMC.py:29
        if (p.uniforms) return p.uniforms[(size_t)(st - 1) * p.B + b];
        return rng_uniform(rng_seed_of(p.seed, p.key_dev), (uint32_t)b, (uint32_t)st);
    };
    constexpr int CW = SPECC ? SPECC : 1;                    // CTAs per walker
    const int crank = SPECC ? (int)cluster_ctarank() : 0;
    double* const splp = reinterpret_cast<double*>(smem_raw + L_::o_red + 144);     // [2] this CTA's node log p, double-buffered
    int sp_round = 0;

    // Work items.  Default: one item per walker (its whole chain), dealt statically.  Segmented mode (seg_len > 0; batches of a
    // few walkers per resident CTA, where whole chains leave a ragged last wave -- 512 walkers on 296 CTA slots are two rounds
    // for 1.73 rounds of work): a chain is cut into segments of seg_len steps, item = (segment, walker) in segment-major order,
    // taken by a global ticket; a segment waits for its walker's previous one (done[b], release / acquire), which was ticketed
    // earlier and therefore belongs to a running CTA -- no deadlock whatever else occupies the device.  State between the
    // segments of a walker travels through x_out / ws_logp / ws_nacc.
    const bool segmented = SPECC == 0 && p.seg_len > 0;
    const int nseg = segmented ? (p.mc_steps + p.seg_len - 1) / p.seg_len : 1;
    const int nitems = nseg * p.B;                       // the launcher keeps this below 2^31
    __shared__ int s_item;
    int item = (int)(blockIdx.x / CW) - (int)(gridDim.x / CW);
    for (;;) {
        if (segmented) {
            __syncthreads();
            if (tid == 0) s_item = (int)atomicAdd(p.qctr, 1ull);
            __syncthreads();
            item = s_item;
        } else {
            item += gridDim.x / CW;
        }
        if (item >= nitems) break;
        const int seg = segmented ? item / p.B : 0, b = item - seg * p.B;
        const int step_lo = seg == 0 ? 0 : seg * p.seg_len + 1;
        const int step_hi = segmented ? min((seg + 1) * p.seg_len, p.mc_steps) : p.mc_steps;
        const bool last_seg = seg == nseg - 1;
        if (seg > 0 && tid == 0) {
            while (ld_acquire_s32(p.done + b) < seg) __nanosleep(200);
        }
        __syncthreads();
        if (tid == 0) s.flag[1] = 0;
        if (tid < N2) s.xs[tid] = seg == 0 ? p.x[(size_t)b * N2 + tid] : __ldcg(p.x_out + (size_t)b * N2 + tid);
        __syncthreads();
        if (tid < n) {
            int si = p.state_idx[(size_t)b * n + tid];
            si = min(max(si, 0), p.M - 1);
            const int mx = (int)rint(p.ktable[2 * si] - p.tw0), my = (int)rint(p.ktable[2 * si + 1] - p.tw1);
            s.kint[tid] = make_int2(mx, my);
            atomicMax(&s.flag[1], max(abs(mx), abs(my)));
        }
        double logp_cur = 0.0;   // thread 0 only
        int nacc = 0;            // thread 0 only (speculative mode: every thread)
        int t_done = 0;          // speculative mode: steps decided so far
        if (seg > 0 && tid == 0) { logp_cur = __ldcg(p.ws_logp + b); nacc = __ldcg(p.ws_nacc + b); }

        for (int it = step_lo;; ++it) {
          int step = it, node = 0, depth = 0;
          bool active = true;
          if constexpr (SPECC == 0) {
              if (step > step_hi) break;
          } else {
              if (it > 0 && t_done >= p.mc_steps) break;
              node = crank + 1;                              // heap numbering: child of m is 2m + accepted
              if (crank == SPECC - 1) node = t_done == 0 ? 0 : (2 * nacc > t_done ? 2 * SPECC - 1 : SPECC);
              step = 0;
              if (node > 0) {
                  depth = 31 - __clz(node);
                  step = t_done + 1 + depth;
                  active = step <= p.mc_steps;
              }
          }
          bool use_pivot = p.force_pivot != 0;
          if (active)
          for (;;) {           // at most two passes: unpivoted block LU, then (rarely) the pivoted fallback
            __syncthreads();
            if (tid == 0) { s.flag[2] = 0; s.flag[3] = 0; }
            // ---------------- stage 0: proposal (MCMC.py:26) and reduced coordinates ----------------
            if (tid < N2) {
                double xv = s.xs[tid];
                if constexpr (SPECC != 0) {
                    for (int j = 0; j < depth; ++j)                     // accepted moves on the path above this node
                        if ((node >> (depth - 1 - j)) & 1) xv += p.sigma * noise_of(b, t_done + 1 + j);
                }
                if (step > 0) xv += p.sigma * noise_of(b, step);
                s.xp[tid] = xv;
                double u = xv * invL;
                u -= floor(u);                                   // features are L-periodic (flow.py:22-25)
                const float uh = (float)u, ul = (float)(u - (double)uh);
                float sn, cs;
                sincospif(uh, &sn, &cs);                         // half angle pi*u
                const float pl = kPi * ul;
                float* pp = reinterpret_cast<float*>(s.pos) + (tid >> 1) * 4 + (tid & 1) * 2;
                pp[0] = fmaf(-pl, sn, cs);
                pp[1] = fmaf(pl, cs, sn);
                reinterpret_cast<double*>(s.us)[tid] = u;
            }
            __syncthreads();

            const float4 pi_ = s.pos[i];
            const int iters = (n + 3) >> 2;

            // ---------------- pass 1: two-particle stream, means over partners (flow.py:34-36,46) ----
            // FP32 terms, Kahan-compensated sums, float64 from the 4-lane butterfly onwards.
            double m0d[5], md[16];
            {
                float m0[5], m0c[5], m[16], mc[16];
#pragma unroll
                for (int f = 0; f < 5; ++f) { m0[f] = 0.f; m0c[f] = 0.f; }
#pragma unroll
                for (int c = 0; c < 16; ++c) { m[c] = 0.f; mc[c] = 0.f; }
                for (int it = 0; it < iters; ++it) {
                    const int j = q + 4 * it;
                    if (j < n) {
                        const PairFeat f = pair_features(pi_, s.pos[j]);
                        const float d = sqrt_approx(f.d2);       // 0 on the diagonal (flow.py:25)
                        kahan_add(m0[0], m0c[0], f.cx); kahan_add(m0[1], m0c[1], f.cy);
                        kahan_add(m0[2], m0c[2], f.sx); kahan_add(m0[3], m0c[3], f.sy);
                        kahan_add(m0[4], m0c[4], d);
#pragma unroll
                        for (int cq = 0; cq < 4; ++cq) {
                            const float4 w0 = wt4[cq], w1 = wt4[4 + cq], w2 = wt4[8 + cq], w3 = wt4[12 + cq],
                                         w4 = wt4[16 + cq], bb = bt4[cq];
                            const float a0 = fmaf(d, w4.x, fmaf(f.sy, w3.x, fmaf(f.sx, w2.x, fmaf(f.cy, w1.x, fmaf(f.cx, w0.x, bb.x)))));
                            const float a1 = fmaf(d, w4.y, fmaf(f.sy, w3.y, fmaf(f.sx, w2.y, fmaf(f.cy, w1.y, fmaf(f.cx, w0.y, bb.y)))));
                            const float a2 = fmaf(d, w4.z, fmaf(f.sy, w3.z, fmaf(f.sx, w2.z, fmaf(f.cy, w1.z, fmaf(f.cx, w0.z, bb.z)))));
                            const float a3 = fmaf(d, w4.w, fmaf(f.sy, w3.w, fmaf(f.sx, w2.w, fmaf(f.cy, w1.w, fmaf(f.cx, w0.w, bb.w)))));
                            kahan_add(m[4 * cq + 0], mc[4 * cq + 0], softplus2(a0));
                            kahan_add(m[4 * cq + 1], mc[4 * cq + 1], softplus2(a1));
                            kahan_add(m[4 * cq + 2], mc[4 * cq + 2], softplus2(a2));
                            kahan_add(m[4 * cq + 3], mc[4 * cq + 3], softplus2(a3));
                        }
                    }
                }
                const double invnd = 1.0 / (double)n;
#pragma unroll
                for (int f = 0; f < 5; ++f) m0d[f] = quad_sum_f64((double)m0[f] - (double)m0c[f]) * invnd;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    md[c] = quad_sum_f64((double)m[c] - (double)mc[c]) * (0.693147180559945309417 * invnd);
            }

            // ---------------- single-particle stream (flow.py:45,51-55): forward values in float64 ------
            float Q[2][5];       // d z_i / d m0_i
            float s0own[4];      // sigma0_i for channels 4q..4q+3
            {
                const double* __restrict__ P = p.params;        // flat float64 parameters (include/cgb200.h)
                constexpr int S_W0 = 96, S_B0 = 240, S_W1 = 256, S_B1 = 1024, S_WF = 1040;
                double h4[4];
                {
                    double pre[4];
                    {
                        const double2 b01 = ldg2(P + S_B0 + 4 * q), b23 = ldg2(P + S_B0 + 4 * q + 2);
                        pre[0] = b01.x; pre[1] = b01.y; pre[2] = b23.x; pre[3] = b23.y;
                    }
#pragma unroll
                    for (int f = 0; f < 5; ++f) {
                        const double2 w01 = ldg2(P + S_W0 + (4 + f) * 16 + 4 * q), w23 = ldg2(P + S_W0 + (4 + f) * 16 + 4 * q + 2);
                        pre[0] = fma(m0d[f], w01.x, pre[0]); pre[1] = fma(m0d[f], w01.y, pre[1]);
                        pre[2] = fma(m0d[f], w23.x, pre[2]); pre[3] = fma(m0d[f], w23.y, pre[3]);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) softplus_sigmoid_f64(pre[k], h4[k], s0own[k]);
                }
                if (rv) {
                    *reinterpret_cast<float4*>(s.sig0 + row * 16 + 4 * q) = make_float4(s0own[0], s0own[1], s0own[2], s0own[3]);
                    double2* hd = reinterpret_cast<double2*>(s.hbuf + row * 18 + 4 * q);
                    hd[0] = make_double2(h4[0], h4[1]);
                    hd[1] = make_double2(h4[2], h4[3]);
                }
                __syncthreads();
                if (tid < 16) {
                    double acc = 0.0;
#pragma unroll 4
                    for (int r = 0; r < n; ++r) acc += s.hbuf[r * 18 + tid];
                    s.hbar[tid] = acc / (double)n;
                }
                __syncthreads();
                double g[4];
                {
                    const double2 b01 = ldg2(P + S_B1 + 4 * q), b23 = ldg2(P + S_B1 + 4 * q + 2);
                    g[0] = b01.x; g[1] = b01.y; g[2] = b23.x; g[3] = b23.y;
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const double he = s.hbuf[i * 18 + e], hb = s.hbar[e];
                    const double* wa = P + S_W1 + e * 16 + 4 * q;
                    const double2 a01 = ldg2(wa), a23 = ldg2(wa + 2);
                    const double2 b01 = ldg2(wa + 256), b23 = ldg2(wa + 258);
                    const double2 c01 = ldg2(wa + 512), c23 = ldg2(wa + 514);
                    g[0] = fma(he, a01.x, fma(hb, b01.x, fma(md[e], c01.x, g[0])));
                    g[1] = fma(he, a01.y, fma(hb, b01.y, fma(md[e], c01.y, g[1])));
                    g[2] = fma(he, a23.x, fma(hb, b23.x, fma(md[e], c23.x, g[2])));
                    g[3] = fma(he, a23.y, fma(hb, b23.y, fma(md[e], c23.y, g[3])));
                }
                float s1own[4];
                double dzx = 0.0, dzy = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double sp;
                    softplus_sigmoid_f64(g[k], sp, s1own[k]);
                    const double sv = h4[k] + sp;                 // s_i (flow.py:52)
                    const double2 wfv = ldg2(P + S_WF + (4 * q + k) * 2);
                    dzx = fma(sv, wfv.x, dzx);
                    dzy = fma(sv, wfv.y, dzy);
                }
                dzx = quad_sum_f64(dzx);
                dzy = quad_sum_f64(dzy);
                if (rv && q == 0) s.dz[row] = make_double2(dzx, dzy);
                // ---- Jacobian-side contractions, FP32:  WfS[e][a] = Wf[e][a] * sigma1_i[e]
                float wfs[16][2];
                {
                    const int gbase = lane & ~3;
#pragma unroll
                    for (int qq = 0; qq < 4; ++qq) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float sg = __shfl_sync(0xffffffffu, s1own[k], gbase + qq);
                            const int e = 4 * qq + k;
                            wfs[e][0] = wf[2 * e] * sg;
                            wfs[e][1] = wf[2 * e + 1] * sg;
                        }
                    }
                }
                // quarter (channels c = 4q+k) of R, T, M; Q partial
                float Qp[2][5];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int f = 0; f < 5; ++f) Qp[a][f] = 0.f;
                float Rq[2][4], Tq[2][4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int c = 4 * q + k;
                    float r0 = 0.f, r1 = 0.f, t0 = 0.f, t1 = 0.f, mm0 = wf[2 * c], mm1 = wf[2 * c + 1];
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const float4 wc = w1c4[c * 4 + e4], wb = w1b4[c * 4 + e4], wa = w1a4[c * 4 + e4];
                        const float wcv[4] = {wc.x, wc.y, wc.z, wc.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w},
                                    wav[4] = {wa.x, wa.y, wa.z, wa.w};
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            const int e = 4 * e4 + kk;
                            r0 = fmaf(wfs[e][0], wcv[kk], r0); r1 = fmaf(wfs[e][1], wcv[kk], r1);
                            t0 = fmaf(wfs[e][0], wbv[kk], t0); t1 = fmaf(wfs[e][1], wbv[kk], t1);
                            mm0 = fmaf(wfs[e][0], wav[kk], mm0); mm1 = fmaf(wfs[e][1], wav[kk], mm1);
                        }
                    }
                    Rq[0][k] = r0 * kLn2; Rq[1][k] = r1 * kLn2;   // ln2: the pair pass uses Wt*log2e
                    Tq[0][k] = t0; Tq[1][k] = t1;
                    mm0 *= s0own[k]; mm1 *= s0own[k];
#pragma unroll
                    for (int f = 0; f < 5; ++f) {
                        const float w = s.wts[W_W0C + f * 16 + c];
                        Qp[0][f] = fmaf(mm0, w, Qp[0][f]);
                        Qp[1][f] = fmaf(mm1, w, Qp[1][f]);
                    }
                }
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int f = 0; f < 5; ++f) Q[a][f] = quad_sum(Qp[a][f]);
                if (rv) {
                    *reinterpret_cast<float4*>(s.Rbuf + row * 36 + 4 * q) = make_float4(Rq[0][0], Rq[0][1], Rq[0][2], Rq[0][3]);
                    *reinterpret_cast<float4*>(s.Rbuf + row * 36 + 16 + 4 * q) = make_float4(Rq[1][0], Rq[1][1], Rq[1][2], Rq[1][3]);
                    *reinterpret_cast<float4*>(s.Tbuf + row * 36 + 4 * q) = make_float4(Tq[0][0], Tq[0][1], Tq[0][2], Tq[0][3]);
                    *reinterpret_cast<float4*>(s.Tbuf + row * 36 + 16 + 4 * q) = make_float4(Tq[1][0], Tq[1][1], Tq[1][2], Tq[1][3]);
                }
            }
            __syncthreads();     // h_i (aliased by H) and the R/T rows are complete before the pair passes touch them

            // ---------------- pass 2: local Jacobian blocks and the mean-field accumulators ----------
            float rs[4] = {0.f, 0.f, 0.f, 0.f};      // partial row sum of the off-diagonal blocks
            {
                float R[2][16];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 v = *reinterpret_cast<const float4*>(s.Rbuf + i * 36 + a * 16 + 4 * c4);
                        R[a][4 * c4 + 0] = v.x; R[a][4 * c4 + 1] = v.y; R[a][4 * c4 + 2] = v.z; R[a][4 * c4 + 3] = v.w;
                    }
                for (int it = 0; it < iters; ++it) {
                    const int k = q + 4 * it;
                    if (k < n && k != i) {
                        const PairFeat f = pair_features(pi_, s.pos[k]);
                        const float i4d = 0.25f * rsqrt_approx(f.d2);
                        // E_ik = (1/n) d t0_ik / d x_k   (5x2, 6 non-zeros)
                        const float e0x = alpha * f.sx, e2x = -alpha * f.cx, e4x = -e0x * i4d;
                        const float e1y = alpha * f.sy, e3y = -alpha * f.cy, e4y = -e1y * i4d;
                        const float d = f.d2 * (4.f * i4d);
                        float j00 = fmaf(Q[0][0], e0x, fmaf(Q[0][2], e2x, Q[0][4] * e4x));
                        float j01 = fmaf(Q[0][1], e1y, fmaf(Q[0][3], e3y, Q[0][4] * e4y));
                        float j10 = fmaf(Q[1][0], e0x, fmaf(Q[1][2], e2x, Q[1][4] * e4x));
                        float j11 = fmaf(Q[1][1], e1y, fmaf(Q[1][3], e3y, Q[1][4] * e4y));
#pragma unroll
                        for (int cq = 0; cq < 4; ++cq) {
                            const float4 w0 = wt4[cq], w1 = wt4[4 + cq], w2 = wt4[8 + cq], w3 = wt4[12 + cq],
                                         w4 = wt4[16 + cq], bb = bt4[cq];
                            const float w0v[4] = {w0.x, w0.y, w0.z, w0.w}, w1v[4] = {w1.x, w1.y, w1.z, w1.w},
                                        w2v[4] = {w2.x, w2.y, w2.z, w2.w}, w3v[4] = {w3.x, w3.y, w3.z, w3.w},
                                        w4v[4] = {w4.x, w4.y, w4.z, w4.w}, bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int c = 4 * cq + kk;
                                const float a = fmaf(d, w4v[kk], fmaf(f.sy, w3v[kk], fmaf(f.sx, w2v[kk], fmaf(f.cy, w1v[kk], fmaf(f.cx, w0v[kk], bv[kk])))));
                                const float sg = sigmoid2(a);
                                // F_ik[c,:] = sigma_t * Wt^T E_ik   (Wt carries log2e; R carries ln2)
                                const float fx = sg * fmaf(w0v[kk], e0x, fmaf(w2v[kk], e2x, w4v[kk] * e4x));
                                const float fy = sg * fmaf(w1v[kk], e1y, fmaf(w3v[kk], e3y, w4v[kk] * e4y));
                                pair_fma<PK>(j00, j01, R[0][c], fx, fy);            // J rows += R_i[c] (fx, fy)
                                pair_fma<PK>(j10, j11, R[1][c], fx, fy);
                            }
                        }
                        if (rv) s.J[row * pB + k] = make_float4(j00, j01, j10, j11);
                        rs[0] += j00; rs[1] += j01; rs[2] += j10; rs[3] += j11;
                    }
                }
            }
            // ---------------- pass 2b: H_i = d hbar / d x_i = (1/n) [ sum_k P_k E_ki - P_i sum_k E_ik ] ----
            {
                float Hacc[32];                      // [c][d]: sum_k sigma0_k * W0c^T (eps . E_ik)
#pragma unroll
                for (int t = 0; t < 32; ++t) Hacc[t] = 0.f;
                float sE[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // sum_k E_ik : (0x,2x,4x,1y,3y,4y)
                for (int it = 0; it < iters; ++it) {
                    const int k = q + 4 * it;
                    if (k < n && k != i) {
                        const PairFeat f = pair_features(pi_, s.pos[k]);
                        const float i4d = 0.25f * rsqrt_approx(f.d2);
                        const float e0x = alpha * f.sx, e2x = -alpha * f.cx, e4x = -e0x * i4d;
                        const float e1y = alpha * f.sy, e3y = -alpha * f.cy, e4y = -e1y * i4d;
#pragma unroll
                        for (int cq = 0; cq < 4; ++cq) {
                            const float4 v0 = w0c4[cq], v1 = w0c4[4 + cq], v2 = w0c4[8 + cq], v3 = w0c4[12 + cq],
                                         v4 = w0c4[16 + cq];
                            const float v0v[4] = {v0.x, v0.y, v0.z, v0.w}, v1v[4] = {v1.x, v1.y, v1.z, v1.w},
                                        v2v[4] = {v2.x, v2.y, v2.z, v2.w}, v3v[4] = {v3.x, v3.y, v3.z, v3.w},
                                        v4v[4] = {v4.x, v4.y, v4.z, v4.w};
                            const float4 sk4 = *reinterpret_cast<const float4*>(s.sig0 + k * 16 + 4 * cq);
                            const float skv[4] = {sk4.x, sk4.y, sk4.z, sk4.w};
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const int c = 4 * cq + kk;
                                // d h_k / d x_i  needs E_ki = eps . E_ik, eps = (-,-,+,+,-)
                                const float gx = fmaf(v2v[kk], e2x, -fmaf(v0v[kk], e0x, v4v[kk] * e4x));
                                const float gy = fmaf(v3v[kk], e3y, -fmaf(v1v[kk], e1y, v4v[kk] * e4y));
                                pair_fma<PK>(Hacc[2 * c], Hacc[2 * c + 1], skv[kk], gx, gy);
                            }
                        }
                        sE[0] += e0x; sE[1] += e2x; sE[2] += e4x; sE[3] += e1y; sE[4] += e3y; sE[5] += e4y;
                    }
                }
#pragma unroll
                for (int t = 0; t < 32; ++t) Hacc[t] = quad_sum(Hacc[t]);
#pragma unroll
                for (int t = 0; t < 6; ++t) sE[t] = quad_sum(sE[t]);
                // H_i = d hbar / d x_i = (1/n) [ sum_k P_k E_ki  -  P_i sum_k E_ik ]
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float s0c = s.sig0[i * 16 + c];
                    const float ox = fmaf(s.wts[W_W0C + c], sE[0], fmaf(s.wts[W_W0C + 32 + c], sE[1], s.wts[W_W0C + 64 + c] * sE[2]));
                    const float oy = fmaf(s.wts[W_W0C + 16 + c], sE[3], fmaf(s.wts[W_W0C + 48 + c], sE[4], s.wts[W_W0C + 64 + c] * sE[5]));
                    Hacc[2 * c] = (Hacc[2 * c] - s0c * ox) * invn;
                    Hacc[2 * c + 1] = (Hacc[2 * c + 1] - s0c * oy) * invn;
                }
                if (rv) {
                    float* hd = s.Hbuf + row * 36;
                    if (q == 0) store_quarter<0>(hd, Hacc);
                    else if (q == 1) store_quarter<1>(hd, Hacc);
                    else if (q == 2) store_quarter<2>(hd, Hacc);
                    else store_quarter<3>(hd, Hacc);
                }
            }
            __syncthreads();

            // ---------------- pass 3: mean-field coupling  J_ik += T_i H_k, then the diagonal block ----
            {
                float T[2][16];
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const float4 v = *reinterpret_cast<const float4*>(s.Tbuf + i * 36 + a * 16 + 4 * c4);
                        T[a][4 * c4 + 0] = v.x; T[a][4 * c4 + 1] = v.y; T[a][4 * c4 + 2] = v.z; T[a][4 * c4 + 3] = v.w;
                    }
                for (int it = 0; it < iters; ++it) {
                    const int k = q + 4 * it;
                    if (k < n && k != i) {
                        float t00 = 0.f, t01 = 0.f, t10 = 0.f, t11 = 0.f;
#pragma unroll
                        for (int c2 = 0; c2 < 8; ++c2) {
                            const float4 hv = *reinterpret_cast<const float4*>(s.Hbuf + k * 36 + 4 * c2);   // (c,x),(c,y),(c+1,x),(c+1,y)
                            pair_fma<PK>(t00, t01, T[0][2 * c2 + 1], hv.z, hv.w);
                            pair_fma<PK>(t00, t01, T[0][2 * c2], hv.x, hv.y);
                            pair_fma<PK>(t10, t11, T[1][2 * c2 + 1], hv.z, hv.w);
                            pair_fma<PK>(t10, t11, T[1][2 * c2], hv.x, hv.y);
                        }
                        if (rv) {
                            float4 v = s.J[row * pB + k];
                            v.x += t00; v.y += t01; v.z += t10; v.w += t11;
                            s.J[row * pB + k] = v;
                        }
                        rs[0] += t00; rs[1] += t01; rs[2] += t10; rs[3] += t11;
                    }
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) rs[t] = quad_sum(rs[t]);
                if (rv && q == 0) {          // translation equivariance: sum_k J_ik = I  (reference tests/test_flow.py:29-32)
                    s.J[row * pB + row] = make_float4(1.f - rs[0], -rs[1], -rs[2], 1.f - rs[3]);
                }
            }
            __syncthreads();
            if (p.out_jac && step == 0) {
                float* oj = p.out_jac + (size_t)b * N2 * N2;
                for (int t = tid; t < N2 * N2; t += NW * 32) {
                    const int rr = t / N2, cc = t - rr * N2;
                    oj[t] = reinterpret_cast<const float*>(s.J)[((rr >> 1) * pB + (cc >> 1)) * 4 + (rr & 1) * 2 + (cc & 1)];
                }
                __syncthreads();
            }

            // ---------------- LU phase: real 2n x 2n FP32 Jacobian, then complex n x n FP64 Slater matrix ---------
            // Both factorisations are REGISTER-RESIDENT and use every thread of the CTA: thread (row, q) keeps the
            // blocks / entries j = q, q+4, ... of its row (the layout the pair passes already produce), MB = 2 NW
            // per thread.  Step k = 4 base + s: the active column is register 0 of quad lane s (broadcast by shuffle);
            // every fourth step the update writes slot m into slot m-1, so register indices stay compile-time in a
            // rolled loop.  Only the pivot line travels through shared memory (double-buffered), one block barrier
            // per elimination step; warps whose rows are all finished skip the arithmetic and only keep the barrier.
            // Jacobian: J = I + (small) for any physical flow, so the fast path is an UNPIVOTED 2x2-block LU
            // (log|det J| = sum_k log|det P_kk|); multipliers and pivot determinants are monitored and an evaluation
            // that leaves the safe range is redone with the partially pivoted scalar LU below (use_pivot).
            constexpr int MB = 2 * NW, LP = 4 * MB + 16;   // lines are padded: tails are cut in chunks of 4 slots
            const int qbase = lane & ~3;
            if (warp == NW - 1) {        // mean backflow displacement (a pure phase of det D; keeps the phases small)
                double mx = 0.0, my = 0.0;
                for (int t = lane; t < n; t += 32) { const double2 v = s.dz[t]; mx += v.x; my += v.y; }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { mx += __shfl_xor_sync(0xffffffffu, mx, o); my += __shfl_xor_sync(0xffffffffu, my, o); }
                if (lane == 0) { s.lsum[12] = mx / (double)n; s.lsum[13] = my / (double)n; }
            }
            if (!use_pivot) {
                float4* const jline = reinterpret_cast<float4*>(s.Rbuf);        // [2][LP]  (R|T rows are dead)
                float4 Jr[MB];
#pragma unroll
                for (int m = 0; m < MB; ++m) {
                    const int j = 4 * m + q;
                    Jr[m] = (rv && j < n) ? s.J[row * pB + j] : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                if (row == 0) {
#pragma unroll
                    for (int m = 0; m < MB; ++m) jline[4 * m + q] = Jr[m];
                }
                float lacc = 0.f;
                bool bad = false;
                __syncthreads();
                for (int base = 0; 4 * base < n; ++base) {
#pragma unroll
                    for (int sst = 0; sst < 4; ++sst) {
                        const int k = 4 * base + sst;
                        if (k >= n) break;
                        const float4* cur = jline + (k & 1) * LP;
                        float4* nxt = jline + ((k + 1) & 1) * LP;
                        const float4 pk = cur[k];                                // pivot block (broadcast)
                        const float det = pk.x * pk.w - pk.y * pk.z;
                        const float idet = rcp_approx(det);
                        lacc += lg2_approx(fabsf(det));                          // identical in every thread
                        bad |= !(fabsf(det) > 0.05f);                            // also catches NaN
                        if (k == n - 1) break;
                        if (8 * warp + 7 > k) {                                  // this warp still has live rows
                            float4 a;
                            a.x = __shfl_sync(0xffffffffu, Jr[0].x, qbase + sst); a.y = __shfl_sync(0xffffffffu, Jr[0].y, qbase + sst);
                            a.z = __shfl_sync(0xffffffffu, Jr[0].z, qbase + sst); a.w = __shfl_sync(0xffffffffu, Jr[0].w, qbase + sst);
                            // -L = -A_ik P_kk^{-1},  P^{-1} = [[w,-y],[-z,x]] / det
                            const float l00 = (a.y * pk.z - a.x * pk.w) * idet, l01 = (a.x * pk.y - a.y * pk.x) * idet;
                            const float l10 = (a.w * pk.z - a.z * pk.w) * idet, l11 = (a.z * pk.y - a.w * pk.x) * idet;
                            bad |= (rv && row > k) && (fabsf(l00) + fabsf(l01) + fabsf(l10) + fabsf(l11)) > 4.f;
#pragma unroll
                            for (int m = 0; m < MB; ++m) {
                                if ((m & 3) == 0 && 4 * (base + m) >= n) break;  // uniform, per chunk of 4 slots: the live tail shrinks
                                const int j = 4 * (base + m) + q;
                                if (sst == 3) {
                                    if (m >= 1) {
                                        const float4 qv = cur[j];
                                        float4 v = Jr[m];
                                        block_fma<PK>(v, l00, l01, l10, l11, qv);
                                        Jr[m - 1] = v;                           // shift: column 4(base+1)+q moves to slot 0
                                    }
                                } else if (m > 0 || q > sst) {
                                    const float4 qv = cur[j];
                                    float4 v = Jr[m];
                                    block_fma<PK>(v, l00, l01, l10, l11, qv);
                                    Jr[m] = v;
                                }
                            }
                            if (row == k + 1) {                                  // publish the next pivot line
                                const int nb = (k + 1) >> 2;
#pragma unroll
                                for (int m = 0; m < MB; ++m) {
                                    if ((m & 3) == 0 && 4 * (nb + m) >= n) break;
                                    nxt[4 * (nb + m) + q] = Jr[m];
                                }
                            }
                        }
                        __syncthreads();
                    }
                }
                if (__any_sync(0xffffffffu, bad) && lane == 0) s.flag[2] = 1;
                if (tid == 0) s.lsum[15] = (double)lacc * 0.693147180559945309417;
            } else if (warp < nJw) {
                // ---- fallback: scalar LU with partial pivoting on the blocked storage, one thread per row.
                // |a| bits with the row id in the low 7 bits -> one REDUX per warp; per-warp maxima exchanged
                // through a double-buffered slot, one named barrier per elimination step.
                const int r = tid;
                float* Jf = reinterpret_cast<float*>(s.J);
                auto el = [&](int rr, int cc) -> float& { return Jf[((rr >> 1) * pB + (cc >> 1)) * 4 + (rr & 1) * 2 + (cc & 1)]; };
                bool done = !(r < N2);
                float mylog = 0.f;
                for (int k = 0; k < N2; ++k) {
                    const float cur = done ? 0.f : el(r, k);
                    const unsigned key = done ? 0u : ((__float_as_uint(fabsf(cur)) & ~0x7Fu) | (unsigned)r);
                    const unsigned wmax = __reduce_max_sync(0xffffffffu, key);
                    if (lane == 0) s.redJ[(k & 1) * 8 + warp] = wmax;
                    if (nJw > 1) named_bar_sync(1, nJw * 32); else __syncwarp();
                    unsigned gmax = s.redJ[(k & 1) * 8];
#pragma unroll
                    for (int w = 1; w < nJw; ++w) gmax = max(gmax, s.redJ[(k & 1) * 8 + w]);
                    const int pr = (int)(gmax & 0x7Fu);
                    const float piv = el(pr, k);
                    if (r == pr) {
                        done = true;
                        mylog = lg2_approx(fabsf(piv));
                    } else if (!done) {
                        const float l = -cur / piv;
                        for (int j = k + 1; j < N2; ++j) el(r, j) = fmaf(l, el(pr, j), el(r, j));
                    }
                    if (nJw > 1) named_bar_sync(1, nJw * 32); else __syncwarp();
                }
                double acc = (double)mylog;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) s.lsum[8 + warp] = acc * 0.693147180559945309417;
            }
            __syncthreads();
            if (use_pivot || s.flag[2] == 0) {
                // ---- Slater matrix in float64, built straight into registers: D_rj = exp(2 pi i m_j . w_r),
                // w = u + (dz - mean dz)/L  (src/slater.py:14-17; the twist and the common displacement are pure
                // row/column phases and drop out of |det|).
                const int mmax = s.flag[1];                      // max |m| over this walker's occupied orbitals
                const bool tab = mmax > 0 && (size_t)n * 2 * mmax * sizeof(double2) <= (size_t)L_::ptab_bytes;
                const double mdx = s.lsum[12], mdy = s.lsum[13];
                if (tab) {
                    // power tables  X_r^m = e^{2 pi i m w_rx},  Y_r^m,  m = 1..mmax  (one sincospi per coordinate)
                    for (int c = tid; c < 2 * n; c += NW * 32) {
                        const double wv = fma(reinterpret_cast<const double*>(s.dz)[c] - ((c & 1) ? mdy : mdx), invL,
                                              reinterpret_cast<const double*>(s.us)[c]);
                        double sn, cs;
                        sincospi(2.0 * (wv - rint(wv)), &sn, &cs);
                        double2* tb = s.ptab + (size_t)c * mmax;   // [row][comp][m-1]
                        double2 pw = make_double2(cs, sn);
                        tb[0] = pw;
                        for (int m = 1; m < mmax; ++m) {
                            pw = make_double2(pw.x * cs - pw.y * sn, pw.x * sn + pw.y * cs);
                            tb[m] = pw;
                        }
                    }
                    __syncthreads();
                }
                auto build_D = [&](double2 (&Dr)[MB]) {
                    const double2 uu = s.us[i], dd = s.dz[i];
                    const double wx = fma(dd.x - mdx, invL, uu.x), wy = fma(dd.y - mdy, invL, uu.y);
#pragma unroll
                    for (int m = 0; m < MB; ++m) {
                        const int j = 4 * m + q;
                        double2 v = make_double2(0.0, 0.0);
                        if (rv && j < n) {
                            const int2 mk = s.kint[j];
                            if (tab) {
                                const int ax = abs(mk.x), ay = abs(mk.y);
                                double2 X = make_double2(1.0, 0.0), Y = make_double2(1.0, 0.0);
                                if (ax) { X = s.ptab[(size_t)(2 * row) * mmax + ax - 1]; if (mk.x < 0) X.y = -X.y; }
                                if (ay) { Y = s.ptab[(size_t)(2 * row + 1) * mmax + ay - 1]; if (mk.y < 0) Y.y = -Y.y; }
                                v = make_double2(X.x * Y.x - X.y * Y.y, X.x * Y.y + X.y * Y.x);
                            } else {
                                double ph = fma((double)mk.x, wx, (double)mk.y * wy);
                                ph -= rint(ph);
                                double sn, cs;
                                sincospi(2.0 * ph, &sn, &cs);
                                v = make_double2(cs, sn);
                            }
                        }
                        Dr[m] = v;
                    }
                };
                // ---- Slater LU, fast path (round 2): the same scheme as the float64 LU below, in FP32 -- entries rounded from the
                // float64 build, update v += l p as two packed FMAs per entry, one FP32 reciprocal (no Newton steps) -- under the
                // a-posteriori error indicator es = sum_k k / |u_kk|^2 of the warp kernel (cgb_wsweep.cu; calibration
                // profiles/r02_slater32_calibration.txt).  Evaluations above the threshold, singular ones and the debug flag
                // take the float64 LU.
                // (n <= 16: the float64 LU is cheap and the random-occupation test states there exceed the indicator on every
                // fourth walker -- measured slower: n=13 B=512 1.43 -> 1.75 ms -- so small n keep the float64 LU)
                bool slater64 = p.force_f64_slater != 0 || n <= 16;
                if (!slater64) {
                    float2 Df[MB];
                    {
                        double2 Dr[MB];
                        build_D(Dr);
#pragma unroll
                        for (int m = 0; m < MB; ++m) Df[m] = make_float2((float)Dr[m].x, (float)Dr[m].y);
                    }
                    float2* const fslots = reinterpret_cast<float2*>(s.Rbuf);      // [2][NW][LP]
                    bool done = !rv, singular = false;
                    float mypn = 1.f, myerr = 0.f;
                    unsigned wcand = 0u;
                    auto offer = [&](int kn) {
                        const float mag = fabsf(Df[0].x) + fabsf(Df[0].y);
                        const unsigned key = (done || q != (kn & 3)) ? 0u : ((__float_as_uint(mag) & ~0x3Fu) | (unsigned)row);
                        wcand = __reduce_max_sync(0xffffffffu, key);
                        if (lane == 0) s.redD[(kn & 1) * 8 + warp] = wcand;
                    };
                    auto publish = [&](int kn) {
                        if ((wcand >> 6) != 0u && row == (int)(wcand & 0x3Fu)) {
                            float2* slot = fslots + ((kn & 1) * NW + warp) * LP;
                            const int nb = kn >> 2;
#pragma unroll
                            for (int m = 0; m < MB; ++m) {
                                if ((m & 3) == 0 && 4 * (nb + m) >= n) break;
                                slot[4 * (nb + m) + q] = Df[m];
                            }
                        }
                    };
                    if (tid < 16) s.redD[tid] = 0u;
                    __syncthreads();
                    offer(0);
                    publish(0);
                    __syncthreads();
                    for (int base = 0; 4 * base < n; ++base) {
#pragma unroll
                        for (int sst = 0; sst < 4; ++sst) {
                            const int k = 4 * base + sst;
                            if (k >= n) break;
                            unsigned gmax;
                            {
                                const uint4 ka = *reinterpret_cast<const uint4*>(s.redD + (k & 1) * 8);
                                const uint4 kb = *reinterpret_cast<const uint4*>(s.redD + (k & 1) * 8 + 4);
                                gmax = max(max(max(ka.x, ka.y), max(ka.z, ka.w)), max(max(kb.x, kb.y), max(kb.z, kb.w)));
                            }
                            singular |= (gmax >> 6) == 0u;
                            const int pr = (int)(gmax & 0x3Fu);
                            const float2* cur = fslots + ((k & 1) * NW + (pr >> 3)) * LP;
                            const float2 piv = cur[k];
                            const float pn = fmaf(piv.x, piv.x, piv.y * piv.y);
                            const float ip = rcp_approx(pn);
                            if (tid == k) { mypn = pn; myerr = (float)(k + 1) * ip; }
                            if (row == pr) done = true;
                            if (k == n - 1) break;
                            if (!__all_sync(0xffffffffu, done)) {
                                const float ivx = piv.x * ip, ivy = -piv.y * ip;     // 1 / pivot
                                const float cx = __shfl_sync(0xffffffffu, Df[0].x, qbase + sst);
                                const float cy = __shfl_sync(0xffffffffu, Df[0].y, qbase + sst);
                                const float2 L1 = make_float2(-(cx * ivx - cy * ivy), -(cx * ivy + cy * ivx));   // l = -c / pivot
                                const float2 L2 = make_float2(-L1.y, L1.x);                                      // i l
#pragma unroll
                                for (int m = 0; m < MB; ++m) {
                                    if ((m & 3) == 0 && 4 * (base + m) >= n) break;
                                    if (m == 4) offer(k + 1);
                                    const int j = 4 * (base + m) + q;
                                    if (sst == 3) {
                                        if (m >= 1) {
                                            const float2 pv = cur[j];
                                            Df[m - 1] = __ffma2_rn(make_float2(pv.x, pv.x), L1, __ffma2_rn(make_float2(pv.y, pv.y), L2, Df[m]));
                                        }
                                    } else if (m > 0 || q > sst) {
                                        const float2 pv = cur[j];
                                        Df[m] = __ffma2_rn(make_float2(pv.x, pv.x), L1, __ffma2_rn(make_float2(pv.y, pv.y), L2, Df[m]));
                                    }
                                }
                                if (4 * (base + 4) >= n || MB <= 4) offer(k + 1);
                            } else {
                                offer(k + 1);
                            }
                            publish(k + 1);
                            __syncthreads();
                        }
                    }
                    // log|det| (per-thread pivot moduli, float64 logarithm) and the error indicator, reduced over the CTA
                    double acc = (tid < n) ? 0.5 * log((double)mypn) : 0.0;
                    float es = (tid < n) ? myerr : 0.f;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { acc += __shfl_xor_sync(0xffffffffu, acc, o); es += __shfl_xor_sync(0xffffffffu, es, o); }
                    if (lane == 0) { s.lsum[warp] = acc; reinterpret_cast<float*>(s.redJ)[warp] = es; }
                    if (__any_sync(0xffffffffu, singular) && lane == 0) s.flag[3] = 1;
                    __syncthreads();
                    float est = 0.f;
#pragma unroll
                    for (int w = 0; w < NW; ++w) est += reinterpret_cast<const float*>(s.redJ)[w];
                    slater64 = !(est < p.slater32_errsum_max * (n > 32 ? 4.f : 1.f)) || s.flag[3] != 0;
                    __syncthreads();
                    if (slater64 && tid == 0) { s.flag[3] = 0; if (p.fallback_counter) atomicAdd(p.fallback_counter + 1, 1ull); }
                }
                if (slater64) {
                // ---- Slater LU: partially pivoted, implicit row exchange.  After its update every unfinished row
                // offers its entry of the next column; one REDUX per warp picks the warp's candidate row, which is
                // published right away into the warp's own line slot together with its key, so ONE block barrier per
                // step suffices: after it every thread reads the NW keys, takes the best and reads that warp's slot.
                double2 Dr[MB];
                build_D(Dr);
                double2* const dslots = reinterpret_cast<double2*>(s.Rbuf);       // [2][NW][LP]
                bool done = !rv, singular = false;
                double mypn = 1.0;
                unsigned wcand = 0u;                      // this warp's candidate key for the next pivot column
                auto offer = [&](int kn) {                // candidates for pivot column kn (held in slot 0 of quad lane kn&3)
                    const float mag = (float)(fabs(Dr[0].x) + fabs(Dr[0].y));
                    const unsigned key = (done || q != (kn & 3)) ? 0u : ((__float_as_uint(mag) & ~0x3Fu) | (unsigned)row);
                    wcand = __reduce_max_sync(0xffffffffu, key);
                    if (lane == 0) s.redD[(kn & 1) * 8 + warp] = wcand;
                };
                auto publish = [&](int kn) {              // the warp's candidate row goes into the warp's own line slot
                    if ((wcand >> 6) != 0u && row == (int)(wcand & 0x3Fu)) {
                        double2* slot = dslots + ((kn & 1) * NW + warp) * LP;
                        const int nb = kn >> 2;
#pragma unroll
                        for (int m = 0; m < MB; ++m) {
                            if ((m & 3) == 0 && 4 * (nb + m) >= n) break;
                            slot[4 * (nb + m) + q] = Dr[m];
                        }
                    }
                };
                if (tid < 16) s.redD[tid] = 0u;
                __syncthreads();
                offer(0);
                publish(0);
                __syncthreads();
                for (int base = 0; 4 * base < n; ++base) {
#pragma unroll
                    for (int sst = 0; sst < 4; ++sst) {
                        const int k = 4 * base + sst;
                        if (k >= n) break;
                        unsigned gmax;
                        {
                            const uint4 ka = *reinterpret_cast<const uint4*>(s.redD + (k & 1) * 8);      // unused slots stay 0
                            const uint4 kb = *reinterpret_cast<const uint4*>(s.redD + (k & 1) * 8 + 4);
                            gmax = max(max(max(ka.x, ka.y), max(ka.z, ka.w)), max(max(kb.x, kb.y), max(kb.z, kb.w)));
                        }
                        singular |= (gmax >> 6) == 0u;
                        const int pr = (int)(gmax & 0x3Fu);
                        const double2* cur = dslots + ((k & 1) * NW + (pr >> 3)) * LP;
                        const double2 piv = cur[k];
                        const double pn = fma(piv.x, piv.x, piv.y * piv.y);
                        if (tid == k) mypn = pn;                               // log taken once, after the loop
                        if (row == pr) done = true;
                        if (k == n - 1) break;
                        if (!__all_sync(0xffffffffu, done)) {
                            double ip = (double)rcp_approx((float)pn);          // 1/pn: FP32 seed + 2 Newton steps
                            ip = ip * (2.0 - pn * ip);
                            ip = ip * (2.0 - pn * ip);
                            const double ivx = piv.x * ip, ivy = -piv.y * ip;   // 1 / pivot
                            const double cx = __shfl_sync(0xffffffffu, Dr[0].x, qbase + sst);
                            const double cy = __shfl_sync(0xffffffffu, Dr[0].y, qbase + sst);
                            const double lx = -(cx * ivx - cy * ivy), ly = -(cx * ivy + cy * ivx);   // l = -c / pivot
#pragma unroll
                            for (int m = 0; m < MB; ++m) {
                                if ((m & 3) == 0 && 4 * (base + m) >= n) break;
                                if (m == 4) offer(k + 1);                       // column k+1 is final after the first chunk: REDUX overlaps the rest
                                const int j = 4 * (base + m) + q;
                                if (sst == 3) {
                                    if (m >= 1) {
                                        const double2 pv = cur[j];
                                        double2 v = Dr[m];
                                        v.x = fma(lx, pv.x, fma(-ly, pv.y, v.x));
                                        v.y = fma(lx, pv.y, fma(ly, pv.x, v.y));
                                        Dr[m - 1] = v;
                                    }
                                } else if (m > 0 || q > sst) {
                                    const double2 pv = cur[j];
                                    double2 v = Dr[m];
                                    v.x = fma(lx, pv.x, fma(-ly, pv.y, v.x));
                                    v.y = fma(lx, pv.y, fma(ly, pv.x, v.y));
                                    Dr[m] = v;
                                }
                            }
                            if (4 * (base + 4) >= n || MB <= 4) offer(k + 1);    // short tail: the loop left before slot 4
                        } else {
                            offer(k + 1);                                       // every row of this warp is finished: key 0
                        }
                        publish(k + 1);
                        __syncthreads();
                    }
                }
                double acc = (tid < n) ? 0.5 * log(mypn) : 0.0;
                if (singular) acc = -INFINITY;            // slogdet of a singular matrix: log|det| = -inf (slater.py:18)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) s.lsum[warp] = acc;
                }   // float64 LU
            }
            __syncthreads();
            if (use_pivot || s.flag[2] == 0) break;
            use_pivot = true;
            if (tid == 0 && p.fallback_counter) atomicAdd(p.fallback_counter, 1ull);
          }

            // ---------------- Metropolis accept / reject (MCMC.py:28-33) ------------------------------
            double lp = 0.0, ldJ = 0.0, ldD = 0.0;
            if (tid == 0 && active) {
                if (use_pivot) { for (int w = 0; w < nJw; ++w) ldJ += s.lsum[8 + w]; } else { ldJ = s.lsum[15]; }
                for (int w = 0; w < NW; ++w) ldD += s.lsum[w];
                ldD -= (double)n * log(p.L);                     // D carries L^{-dim/2} per entry (slater.py:17)
                lp = 2.0 * ldD + ldJ;                            // logpsi.py:30,179
            }
            if constexpr (SPECC == 0) {
            if (tid == 0) {
                int acc = 1;
                if (step == 0) {
                    logp_cur = lp;
                    if (p.out_parts) { p.out_parts[2 * (size_t)b] = ldD; p.out_parts[2 * (size_t)b + 1] = ldJ; }
                } else {
                    const double u = uniform_of(b, step);
                    const double ratio = exp(lp - logp_cur);     // MCMC.py:28
                    acc = (u < ratio) ? 1 : 0;                   // MCMC.py:29  (NaN -> reject)
                    if (acc) { logp_cur = lp; ++nacc; }
                    if (p.out_acc_trace) p.out_acc_trace[(size_t)(step - 1) * p.B + b] = (uint8_t)acc;
                }
                if (p.out_logp_trace) p.out_logp_trace[(size_t)step * p.B + b] = lp;
                s.flag[0] = acc;
            }
            __syncthreads();
            if (step > 0 && s.flag[0] && tid < N2) s.xs[tid] = s.xp[tid];     // MCMC.py:31
            } else {
            // speculative mode: publish this node's log p, one cluster barrier, replay the decisions of the realised path
            double* const myslot = splp + (sp_round & 1);
            if (tid == 0) {
                *myslot = lp;
                if (node == 0 && p.out_parts) { p.out_parts[2 * (size_t)b] = ldD; p.out_parts[2 * (size_t)b + 1] = ldJ; }
            }
            cluster_sync_all();
            if (tid == 0) {
                double lps[SPECC];
#pragma unroll
                for (int r = 0; r < SPECC; ++r) lps[r] = ld_dsmem_f64(myslot, r);
                const bool first = t_done == 0, writer = crank == 0;
                const int extra = 2 * nacc > t_done ? 2 * SPECC - 1 : SPECC;     // the same expression the round started with
                if (first) {
                    logp_cur = lps[SPECC - 1];
                    if (p.out_logp_trace && writer) p.out_logp_trace[b] = logp_cur;
                }
                int nd = 1, adv = 0, mask = 0;
                while (t_done + adv < p.mc_steps) {
                    const int sl = nd < SPECC ? nd - 1 : ((!first && nd == extra) ? SPECC - 1 : -1);
                    if (sl < 0) break;                           // not evaluated in this round
                    const int st = t_done + adv + 1;
                    double lpn = lps[0];
#pragma unroll
                    for (int r = 1; r < SPECC; ++r) lpn = sl == r ? lps[r] : lpn;
                    const double u = uniform_of(b, st);
                    const double ratio = exp(lpn - logp_cur);    // MCMC.py:28
                    const int acc = (u < ratio) ? 1 : 0;         // MCMC.py:29  (NaN -> reject)
                    if (acc) { logp_cur = lpn; mask |= 1 << adv; }
                    if (writer) {
                        if (p.out_acc_trace) p.out_acc_trace[(size_t)(st - 1) * p.B + b] = (uint8_t)acc;
                        if (p.out_logp_trace) p.out_logp_trace[(size_t)st * p.B + b] = lpn;
                    }
                    nd = 2 * nd + acc;
                    ++adv;
                }
                s.flag[0] = adv | (mask << 8);
            }
            __syncthreads();
            {
                const int adv = s.flag[0] & 0xff, mask = s.flag[0] >> 8;
                if (tid < N2) {
                    double xv = s.xs[tid];
                    for (int j = 0; j < adv; ++j)
                        if ((mask >> j) & 1) xv += p.sigma * noise_of(b, t_done + 1 + j);      // MCMC.py:31, in order
                    s.xs[tid] = xv;
                }
                nacc += __popc(mask);
                t_done += adv;
            }
            ++sp_round;
            }
        }   // steps

        __syncthreads();
        if (p.mc_steps > 0 && tid < N2 && crank == 0) {
            double xv = s.xs[tid];
            if (p.wrap && last_seg) xv -= p.L * floor(xv / p.L);             // VMC.py:24
            p.x_out[(size_t)b * N2 + tid] = xv;
        }
        if (tid == 0 && crank == 0) {
            if (last_seg) {
                if (p.out_logp) p.out_logp[b] = logp_cur;
                if (p.out_naccept) p.out_naccept[b] = nacc;
                if (p.out_stats && p.mc_steps > 0) atomicAdd(p.out_stats + CGB_STATS_NACCEPT, (double)nacc);   // integers: exact, order-free
            } else {
                p.ws_logp[b] = logp_cur; p.ws_nacc[b] = nacc;
            }
        }
        if (!last_seg) {                                  // hand the walker to whoever holds its next segment
            __threadfence();
            __syncthreads();
            if (tid == 0) { __threadfence(); st_release_s32(p.done + b, seg + 1); }
        }
    }
    if constexpr (SPECC != 0) cluster_sync_all();     // no CTA leaves while a neighbour may still read its log p slots
}

template <int NW, int SPECC>
static int launch_cluster(const SweepParams& p, cudaStream_t st, size_t smem, bool* fits) {
    auto kern = sweep_kernel<NW, SPECC>;
    *fits = false;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = SPECC; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3((unsigned)(SPECC * p.B), 1, 1);
    cfg.blockDim = dim3(NW * 32, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    e = cudaOccupancyMaxActiveClusters(&nclusters, kern, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); return 0; }      // cannot tell: let the caller try a narrower tree
    if (nclusters < p.B) return 0;                               // the whole batch must be in flight at once to pay off
    *fits = true;
    e = cudaLaunchKernelEx(&cfg, kern, p);
    return (int)(e != cudaSuccess ? e : cudaGetLastError());
}

template <int NW>
static int launch_nw(const SweepParams& p, cudaStream_t st) {
    const size_t smem = sweep_smem_bytes(p.n);
    // speculative mode (p.cspec: -1 = widest decision tree whose clusters are all resident at once, 0 = off, 4/8 = that width or off; a pair
    // of CTAs -- 1.5 steps per round for twice the work -- was measured not to pay)
    if (p.mc_steps > 0 && p.cspec != 0) {
        bool fits = false;
        int rc;
        if (p.cspec == 8 || p.cspec < 0) { rc = launch_cluster<NW, 8>(p, st, smem, &fits); if (fits || rc) return rc; }
        if (p.cspec == 4 || p.cspec < 0) { rc = launch_cluster<NW, 4>(p, st, smem, &fits); if (fits || rc) return rc; }
    }
    cudaError_t e = cudaFuncSetAttribute(sweep_kernel<NW, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sweep_kernel<NW, 0>, NW * 32, smem);
    if (occ < 1) occ = 1;
    int grid = p.B < sms * occ ? p.B : sms * occ;
    SweepParams q = p;
    q.seg_len = 0;
    if (NW >= 5 && p.mc_steps >= 8 && p.B > grid) {
        // whole chains: ceil(B / grid) rounds; segments of mc_steps / 16 steps level that to within a few per cent.  Only where an
        // evaluation is long against a segment's start-up (n > 32) and the ragged wave costs more than 3 %
        const long long rounds = ((long long)p.B + grid - 1) / grid;
        if ((double)(rounds * grid) > 1.03 * (double)p.B) q.seg_len = p.mc_steps / 16 > 1 ? p.mc_steps / 16 : 1;
    }
    if (p.seg_override >= 0) q.seg_len = p.seg_override < p.mc_steps ? p.seg_override : 0;
    if (q.seg_len > 0 && ((long long)p.mc_steps / q.seg_len + 2) * (long long)p.B + (long long)grid >= (1ll << 31)) q.seg_len = 0;
    sweep_kernel<NW, 0><<<grid, NW * 32, smem, st>>>(q);
    return (int)cudaGetLastError();
}

int launch_sweep(const SweepParams& p, cudaStream_t st) {
    switch (sweep_num_warps(p.n)) {
        case 2: return launch_nw<2>(p, st);
        case 3: return launch_nw<3>(p, st);
        case 4: return launch_nw<4>(p, st);
        case 5: return launch_nw<5>(p, st);
        case 6: return launch_nw<6>(p, st);
        case 7: return launch_nw<7>(p, st);
        case 8: return launch_nw<8>(p, st);
        default: return CGB_E_UNSUPPORTED;
    }
}

int launch_prep_params(const SweepParams& p, float* wpack, cudaStream_t st) {
    prep_params_kernel<<<1, 256, 0, st>>>(p.params, wpack, p.out_stats, (double)p.B * (double)p.mc_steps,
                                          const_cast<uint32_t*>(p.key_dev), p.advance_key, p.done, p.qctr, p.B);
    return (int)cudaGetLastError();
}

}  // namespace cgb
