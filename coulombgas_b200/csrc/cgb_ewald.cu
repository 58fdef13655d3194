// Ewald Coulomb energy of a walker batch for sm_100a  (src/potential.py:36-77).
//
// The reference evaluates  sum_G g_G sum_{i<j} cos(2 pi G.r_ij)  pairwise (O(|G| n^2)); here the same sum
// is taken in structure-factor form  1/2 (|S(G)|^2 - n),  S(G) = sum_i e^{2 pi i G.u_i}  (O(|G| n)), with
// e^{2 pi i G.u} = px_i[|Gx|] * py_i[|Gy|]^(*) from per-particle power tables kept in shared memory, and the
// (Gx,Gy)/(Gx,-Gy)/(-G) quadruple sharing one set of four real accumulators (host-side merge in
// cgb_ewald_prepare_host).  Real space: nearest image only, erfc(kappa d)/d (potential.py:45-50).
// One CTA (128 threads) per walker, persistent over the batch.
#include "cgb_common.cuh"
#include "cgb_internal.h"

namespace cgb {

constexpr int EW_THREADS = 128;

__global__ void __launch_bounds__(EW_THREADS, 8) ewald_kernel(const EwaldParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = p.n, tid = threadIdx.x, np1 = p.gmax + 1;
    double* u = reinterpret_cast<double*>(smem_raw);                    // [n][2] reduced coordinates
    float2* px = reinterpret_cast<float2*>(u + 2 * n);                  // [n][np1]
    float2* py = px + n * np1;                                          // [n][np1]
    __shared__ double red[EW_THREADS / 32];
    const double invL = 1.0 / p.L;
    const float kap = (float)p.kappa;

    for (int b = blockIdx.x; b < p.B; b += gridDim.x) {
        __syncthreads();
        for (int t = tid; t < 2 * n; t += EW_THREADS) {
            double v = p.x[(size_t)b * 2 * n + t] * invL;               // x / L  (potential.py:77)
            u[t] = v - floor(v);
        }
        __syncthreads();
        // power tables e^{2 pi i g u}, g = 0..gmax, phases reduced exactly in float64
        for (int t = tid; t < 2 * n * np1; t += EW_THREADS) {
            const int g = t % np1, c = t / np1;                          // c = particle*2 + comp
            double ph = (double)g * u[c];
            ph -= floor(ph);
            float sn, cs;
            sincospif(2.f * (float)ph, &sn, &cs);
            ((c & 1) ? py : px)[(c >> 1) * np1 + g] = make_float2(cs, sn);
        }
        __syncthreads();
        float acc = 0.f;
        // reciprocal space
        for (int t = tid; t < p.n_items; t += EW_THREADS) {
            const float4 it = p.items[t];
            const int gx = (int)it.x, gy = (int)it.y;
            float ac = 0.f, bd = 0.f, ad = 0.f, bc = 0.f;
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
                const float2 a = px[i * np1 + gx], c = py[i * np1 + gy];
                ac = fmaf(a.x, c.x, ac); bd = fmaf(a.y, c.y, bd);
                ad = fmaf(a.x, c.y, ad); bc = fmaf(a.y, c.x, bc);
            }
            const float spr = ac - bd, spi = ad + bc, smr = ac + bd, smi = bc - ad;
            const float fn = (float)n;
            acc = fmaf(0.5f * it.z, fmaf(spr, spr, spi * spi) - fn, acc);
            acc = fmaf(0.5f * it.w, fmaf(smr, smr, smi * smi) - fn, acc);
        }
        // real space, nearest image (potential.py:45-50)
        for (int t = tid; t < n * n; t += EW_THREADS) {
            const int i = t / n, j = t % n;
            if (i < j) {
                double dx = u[2 * i] - u[2 * j], dy = u[2 * i + 1] - u[2 * j + 1];
                dx -= rint(dx);
                dy -= rint(dy);
                const float fx = (float)dx, fy = (float)dy;
                const float d2 = fmaf(fx, fx, fy * fy);
                const float rd = rsqrtf(d2);
                acc = fmaf(erfcf(kap * (d2 * rd)), rd, acc);
            }
        }
        double a = (double)acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
        if ((tid & 31) == 0) red[tid >> 5] = a;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int w = 0; w < EW_THREADS / 32; ++w) tot += red[w];
            const double g0 = -2.0 * sqrt(3.14159265358979323846) / p.kappa;       // potential.py:59
            tot += g0 * (0.5 * n * (n - 1));                                      // potential.py:62
            p.out[b] = 2.0 * p.rs * invL * tot;                                   // potential.py:77
        }
    }
}

int launch_ewald(const EwaldParams& p, cudaStream_t st) {
    const size_t smem = sizeof(double) * 2 * p.n + sizeof(float2) * 2 * p.n * (p.gmax + 1);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(ewald_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    int dev = 0, sms = 148, occ = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ewald_kernel, EW_THREADS, smem);
    if (occ < 1) occ = 1;
    const int grid = p.B < sms * occ ? p.B : sms * occ;
    ewald_kernel<<<grid, EW_THREADS, smem, st>>>(p);
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

}  // namespace cgb
