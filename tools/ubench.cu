// Pipe-throughput micro-benchmarks for sm_100a (B200): which instruction mixes the sweep kernel can lean on.
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/ubench.bin tools/ubench.cu
// Run  :  tools/ubench.bin            (prints ops / clk / SM for every kind, one line each)
// Each kind runs 148*R resident CTAs of 256 threads; rate = work per CTA * R / max CTA cycles (clock64).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define ITERS 4096

__device__ __forceinline__ float ex2a(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int KIND>
__global__ void __launch_bounds__(256) k(float* out, long long* cyc, float seed) {
    __shared__ float4 sm[64];
    if (threadIdx.x < 64) sm[threadIdx.x] = make_float4(seed, seed * 0.5f, seed * 0.25f, 1.f);
    __syncthreads();
    float a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const float b = 1.000001f, c = 1e-7f;
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
    float2 p4 = p0, p5 = p1, p6 = p2, p7 = p3;
    const float2 b2 = make_float2(b, b), c2 = make_float2(c, c);
    double d0 = a0, d1 = a1, d2 = a2, d3 = a3;
    float acc[4][4] = {};
    unsigned fa[4] = {__float_as_uint(a0), __float_as_uint(a1), __float_as_uint(a2), __float_as_uint(a3)};
    unsigned fb[2] = {__float_as_uint(b), __float_as_uint(c)};
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (KIND == 0) {          // FFMA x8
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        } else if (KIND == 1) {   // FFMA2 x8 (16 FMAs)
            p0 = __ffma2_rn(p0, b2, c2); p1 = __ffma2_rn(p1, b2, c2); p2 = __ffma2_rn(p2, b2, c2); p3 = __ffma2_rn(p3, b2, c2);
            p4 = __ffma2_rn(p4, b2, c2); p5 = __ffma2_rn(p5, b2, c2); p6 = __ffma2_rn(p6, b2, c2); p7 = __ffma2_rn(p7, b2, c2);
        } else if (KIND == 2) {   // 8 FFMA + 2 MUFU
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = ex2a(a6); a7 = ex2a(a7);
            a6 = fmaf(a6, c, c); a7 = fmaf(a7, c, c);
        } else if (KIND == 3) {   // 4 FFMA2 (8 FMAs) + 2 MUFU + 2 FFMA
            p0 = __ffma2_rn(p0, b2, c2); p1 = __ffma2_rn(p1, b2, c2); p2 = __ffma2_rn(p2, b2, c2); p3 = __ffma2_rn(p3, b2, c2);
            a6 = ex2a(a6); a7 = ex2a(a7);
            a6 = fmaf(a6, c, c); a7 = fmaf(a7, c, c);
        } else if (KIND == 4) {   // DFMA x4
            d0 = fma(d0, 1.0000001, 1e-9); d1 = fma(d1, 1.0000001, 1e-9); d2 = fma(d2, 1.0000001, 1e-9); d3 = fma(d3, 1.0000001, 1e-9);
        } else if (KIND == 5) {   // mma.sync m16n8k8 tf32 x4 independent accumulators
#pragma unroll
            for (int t = 0; t < 4; ++t)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(acc[t][0]), "+f"(acc[t][1]), "+f"(acc[t][2]), "+f"(acc[t][3])
                             : "r"(fa[0]), "r"(fa[1]), "r"(fa[2]), "r"(fa[3]), "r"(fb[0]), "r"(fb[1]));
        } else if (KIND == 6) {   // 8 FFMA + 1 broadcast LDS.128
            const float4 w = sm[it & 63];
            a0 = fmaf(a0, w.x, c); a1 = fmaf(a1, w.y, c); a2 = fmaf(a2, w.z, c); a3 = fmaf(a3, w.w, c);
            a4 = fmaf(a4, w.x, c); a5 = fmaf(a5, w.y, c); a6 = fmaf(a6, w.z, c); a7 = fmaf(a7, w.w, c);
        } else if (KIND == 7) {   // 4 FFMA2 + 1 broadcast LDS.128
            const float4 w = sm[it & 63];
            const float2 wl = make_float2(w.x, w.y), wh = make_float2(w.z, w.w);
            p0 = __ffma2_rn(p0, wl, c2); p1 = __ffma2_rn(p1, wh, c2); p2 = __ffma2_rn(p2, wl, c2); p3 = __ffma2_rn(p3, wh, c2);
        } else if (KIND == 8) {   // SHFL x4
            a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2);
            a2 = __shfl_xor_sync(0xffffffffu, a2, 1); a3 = __shfl_xor_sync(0xffffffffu, a3, 2);
        } else if (KIND == 9) {   // 4 FFMA2 + 4 FADD2 (packed add)
            p0 = __ffma2_rn(p0, b2, c2); p1 = __ffma2_rn(p1, b2, c2); p2 = __ffma2_rn(p2, b2, c2); p3 = __ffma2_rn(p3, b2, c2);
            p4 = __fadd2_rn(p4, c2); p5 = __fadd2_rn(p5, c2); p6 = __fadd2_rn(p6, c2); p7 = __fadd2_rn(p7, c2);
        } else if (KIND == 10) {  // 4 FFMA2 + 4 integer ALU ops (LOP3) -> do packed FMAs leave issue slots for the ALU pipe?
            p0 = __ffma2_rn(p0, b2, c2); p1 = __ffma2_rn(p1, b2, c2); p2 = __ffma2_rn(p2, b2, c2); p3 = __ffma2_rn(p3, b2, c2);
            fa[0] = (fa[0] ^ fa[1]) & 0x7fffffffu; fa[1] = (fa[1] ^ fa[2]) | 1u; fa[2] = (fa[2] ^ fa[3]) & 0x3fffffffu; fa[3] = (fa[3] ^ fa[0]) | 2u;
        } else if (KIND == 11) {  // 8 FFMA + 4 integer ALU ops
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
            fa[0] = (fa[0] ^ fa[1]) & 0x7fffffffu; fa[1] = (fa[1] ^ fa[2]) | 1u; fa[2] = (fa[2] ^ fa[3]) & 0x3fffffffu; fa[3] = (fa[3] ^ fa[0]) | 2u;
        } else if (KIND == 12) {  // MUFU x4 only
            a0 = ex2a(a0); a1 = ex2a(a1); a2 = ex2a(a2); a3 = ex2a(a3);
        }
    }
    const long long t1 = clock64();
    float r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y + p4.x + p4.y + p5.x +
              p5.y + p6.x + p6.y + p7.x + p7.y + (float)(d0 + d1 + d2 + d3);
    for (int t = 0; t < 4; ++t) r += acc[t][0] + acc[t][1] + acc[t][2] + acc[t][3] + __uint_as_float(fa[t]);
    if (r == 123.456f) out[0] = r;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

struct Kind { const char* name; double work_per_thread_iter; const char* unit; };

template <int KIND>
static void run(const Kind& kd, int R, float* out, long long* cyc, int sms) {
    const int grid = sms * R;
    k<KIND><<<grid, 256>>>(out, cyc, 0.5f);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND><<<grid, 256>>>(out, cyc, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
    const long long mx = *std::max_element(h.begin(), h.end());
    const double per_sm_clk = kd.work_per_thread_iter * 256.0 * ITERS * R / (double)mx;
    const double total = kd.work_per_thread_iter * 256.0 * ITERS * grid / (ms * 1e-3);
    printf("%-44s R=%d  %8.1f %s/clk/SM   %8.3f T%s/s  (%.3f ms, %lld cyc)\n", kd.name, R, per_sm_clk, kd.unit, total * 1e-12, kd.unit, ms, mx);
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; long long* cyc;
    cudaMalloc(&out, 1024); cudaMalloc(&cyc, sizeof(long long) * sms * 8);
    const Kind kinds[] = {
        {"FFMA x8", 8, "FMA"}, {"FFMA2 x8 (16 FMA)", 16, "FMA"}, {"6 FFMA + 2 MUFU + 2 FFMA", 10, "inst"},
        {"4 FFMA2 + 2 MUFU + 2 FFMA (thread-inst 8)", 8, "inst"}, {"DFMA x4", 4, "DFMA"},
        {"mma.sync m16n8k8 tf32 x4 (MACs)", 4 * 1024.0 / 32.0, "MAC"}, {"8 FFMA + LDS.128 bcast", 8, "FMA"},
        {"4 FFMA2 + LDS.128 bcast", 8, "FMA"}, {"SHFL x4", 4, "SHFL"}, {"4 FFMA2 + 4 FADD2 (flop-pairs 16)", 16, "op"},
        {"4 FFMA2 + 4 LOP3 (FMA count 8)", 8, "FMA"}, {"8 FFMA + 4 LOP3 (FMA count 8)", 8, "FMA"}, {"MUFU.EX2 x4", 4, "MUFU"}};
    for (int R : {2, 4}) {
        run<0>(kinds[0], R, out, cyc, sms); run<1>(kinds[1], R, out, cyc, sms); run<2>(kinds[2], R, out, cyc, sms);
        run<3>(kinds[3], R, out, cyc, sms); run<4>(kinds[4], R, out, cyc, sms); run<5>(kinds[5], R, out, cyc, sms);
        run<6>(kinds[6], R, out, cyc, sms); run<7>(kinds[7], R, out, cyc, sms); run<8>(kinds[8], R, out, cyc, sms);
        run<9>(kinds[9], R, out, cyc, sms); run<10>(kinds[10], R, out, cyc, sms); run<11>(kinds[11], R, out, cyc, sms);
        run<12>(kinds[12], R, out, cyc, sms);
    }
    return 0;
}
