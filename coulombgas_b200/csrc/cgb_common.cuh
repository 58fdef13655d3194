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

}  // namespace cgb
