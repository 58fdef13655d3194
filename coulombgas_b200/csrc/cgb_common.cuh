// Shared device helpers for libcgb200 (sm_100a).  Approximate transcendental wrappers map 1:1 onto the
// SFU (MUFU.EX2 / LG2 / RCP / RSQ / SQRT); everything else is plain FP32 FMA-pipe arithmetic.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cgb {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kPi = 3.14159265358979323846f;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// softplus in log2 units: log2(1 + 2^a) = max(a,0) + log2(1 + 2^-|a|).   2 MUFU.
__device__ __forceinline__ float softplus2(float a) {
    return fmaxf(a, 0.f) + lg2_approx(1.f + ex2_approx(-fabsf(a)));
}
// logistic sigmoid of a/log2e given a in log2 units: 1/(1+2^-a).  2 MUFU.  2^-a -> inf gives 0.
__device__ __forceinline__ float sigmoid2(float a) { return rcp_approx(1.f + ex2_approx(-a)); }

// softplus (natural units) and its derivative from a natural-unit pre-activation; shares the exp.
__device__ __forceinline__ void softplus_sigmoid(float pre, float& sp, float& sg) {
    const float a = pre * kLog2e;
    const float e = ex2_approx(-fabsf(a));
    const float r = rcp_approx(1.f + e);
    sp = kLn2 * (fmaxf(a, 0.f) + lg2_approx(1.f + e));
    sg = a >= 0.f ? r : e * r;
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) of a contiguous block global -> shared, completion on an mbarrier.
// Called by EVERY thread of the CTA (contains two block barriers); thread 0 arms the barrier with the byte count and issues the
// copy, everybody waits on the phase-0 parity.  dst / src 16-byte aligned, bytes a multiple of 16; `mbar` is 8 bytes of shared
// memory that nobody else touches.  The weight blocks of the sweep kernels (a few kB, staged once per CTA) come in this way.
__device__ __forceinline__ void tma_stage_block(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* mbar) {
    const unsigned mb = (unsigned)__cvta_generic_to_shared(mbar), dst = (unsigned)__cvta_generic_to_shared(dst_smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(src_gmem), "r"(bytes), "r"(mb) : "memory");
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mb) : "memory");
    }
    __syncthreads();
}

// ---- Philox4x32-10 (Salmon et al. 2011), counter-based RNG for the throughput mode --------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// uniform in the OPEN interval (0,1) from the top 23 bits: (k + 1/2) 2^-23 is exact in FP32, so neither end point can be
// produced by rounding (a 24-bit variant rounds its largest value to 1.0f: log(1) = 0 breaks -log(-log(u)))
__device__ __forceinline__ float u32_to_unit_open(uint32_t r) {
    return ((float)(r >> 9) + 0.5f) * (1.0f / 8388608.0f);
}
// the same with a second word behind it: when the top 23 bits are all zero (probability 2^-23) the next 23 bits refine the
// value, so the lower tail reaches 2^-47 instead of 2^-24 (Box-Muller radius up to 8 sigma instead of 5.8)
__device__ __forceinline__ float u32x2_to_unit_open(uint32_t hi, uint32_t lo) {
    const uint32_t k = hi >> 9;
    return k ? ((float)k + 0.5f) * (1.0f / 8388608.0f) : ((float)(lo >> 9) + 0.5f) * (1.0f / 70368744177664.0f);
}

// ---- the throughput mode's random streams (shared by every kernel and by the debug export that the statistical tests
// read): all keyed by (seed; walker, step / draw, particle / orbital, stream id), i.e. independent of the schedule ----------
// proposal noise of (walker b, step s, particle i): two independent N(0,1) (MCMC.py:26), Box-Muller with accurate logf /
// sincospif (the approximate __logf has an absolute error that distorts the radius near u = 1)
__device__ __forceinline__ float2 rng_normal2(uint64_t seed, uint32_t b, uint32_t s, uint32_t i) {
    const uint4 r = philox4x32_10(make_uint4(b, s, i, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float rad = sqrtf(-2.f * logf(u32x2_to_unit_open(r.x, r.z)));
    float sn, cs;
    sincospif(2.f * u32_to_unit_open(r.y), &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}
// Metropolis uniform of (walker b, step s) in [0, 1) with 53 random bits (MCMC.py:29)
__device__ __forceinline__ double rng_uniform(uint64_t seed, uint32_t b, uint32_t s) {
    const uint4 r = philox4x32_10(make_uint4(b, s, 0xffffffffu, 1u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return (double)((((unsigned long long)r.x << 32) | (unsigned long long)r.y) >> 11) * (1.0 / 9007199254740992.0);
}
// Gumbel noise of (walker b, draw i, orbital j) for the categorical draws of the momentum sampler (sampler.py:36-37)
__device__ __forceinline__ double rng_gumbel(uint64_t seed, uint32_t b, uint32_t i, uint32_t j) {
    const uint4 r = philox4x32_10(make_uint4(b, i, j, 2u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return (double)(-logf(-logf(u32x2_to_unit_open(r.x, r.z))));        // u strictly inside (0,1): finite
}
// the seed of a call: the host value, or -- when the caller passes a device-resident (2,) uint32 PRNG key, the form a
// traced jax key has inside pmap (src/VMC.py:20, src/MCMC.py:24) -- key[0] << 32 | key[1] read by the kernel itself
__device__ __forceinline__ uint64_t rng_seed_of(uint64_t seed_host, const uint32_t* key_dev) {
    if (key_dev) return ((uint64_t)__ldg(key_dev) << 32) | (uint64_t)__ldg(key_dev + 1);
    return seed_host;
}

}  // namespace cgb
