// tools/ubench.cu — instruction-throughput microbenchmarks on B200 (sm_100a) that decide the
// rollout kernel's design: scalar FFMA vs packed FFMA2 (fma.rn.f32x2), min3, select chains, MUFU.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
// Run:   tools/ubench   (prints warp-instructions per clock per SM for each pattern)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 2048
#define UNROLL 8

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    ra = *reinterpret_cast<unsigned long long*>(&a);
    rb = *reinterpret_cast<unsigned long long*>(&b);
    rc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}

// 1. scalar FFMA, 8 independent chains, loop-invariant multiplicand/addend
__global__ void k_ffma_inv(float* out, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], a, b);
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 2. scalar FFMA with three distinct, per-chain registers
__global__ void k_ffma_3reg(float* out, float a, float b) {
    float x[8], y[8], z[8];
    for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x + i; y[i] = a + i * 1e-3f; z[i] = b + i * 1e-3f; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fmaf(x[i], y[i], z[(i + u) & 7]);
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 3. packed FFMA2, 8 independent chains
__global__ void k_ffma2(float* out, float a, float b) {
    float2 x[8], y[8], z[8];
    for (int i = 0; i < 8; ++i) { x[i] = make_float2(threadIdx.x + i, i); y[i] = make_float2(a, a + i * 1e-3f); z[i] = make_float2(b, b + i * 1e-3f); }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = ffma2(x[i], y[i], z[(i + u) & 7]);
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i].x + x[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 4. FMNMX3 chains
__global__ void k_min3(float* out, float a, float b) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) x[i] = fminf(fminf(x[i], a + u), x[(i + 1) & 7] + b);   // FADD + FMNMX3
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 5. compare+select pairs (FSETP + FSEL + SEL), the exact arg-min node
__global__ void k_argmin_node(float* out, int* outi, float a) {
    float d[8]; int id[8];
    for (int i = 0; i < 8; ++i) { d[i] = threadIdx.x * a + i; id[i] = i; }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = (i + 1 + u) & 7;
                const bool lt = d[j] < d[i];
                d[i] = lt ? d[j] + a : d[i];        // keeps values moving so nothing folds away
                id[i] = lt ? id[j] : id[i];
            }
    float s = 0; int t = 0; for (int i = 0; i < 8; ++i) { s += d[i]; t += id[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s; outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
// 6. LOP3 + FMNMX packed-key node
__global__ void k_packed_node(float* out, float a) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * a + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float k = __int_as_float((__float_as_int(x[(i + 1) & 7]) & ~31) | u);
                x[i] = fminf(x[i] * a, k);
            }
    float s = 0; for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// 7. MUFU (ex2) chains
__global__ void k_mufu(float* out) {
    float x[4];
    for (int i = 0; i < 4; ++i) x[i] = 0.001f * (threadIdx.x + i);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}
// 8. mix: 2 FFMA + argmin node per candidate (the exact search inner pattern), 30 candidates
__global__ void k_search_exact(float* out, const float4* __restrict__ tab, float xs, float ys) {
    float wa[30], wb[30], wc[30];
    for (int j = 0; j < 30; ++j) { float4 t = tab[j]; wa[j] = t.x; wb[j] = t.y; wc[j] = t.z; }
    float xl = threadIdx.x * xs, yl = threadIdx.x * ys; int acc = 0;
    for (int it = 0; it < ITERS; ++it) {
        float d[32]; int id[32];
#pragma unroll
        for (int j = 0; j < 30; ++j) { d[j] = fmaf(wa[j], xl, fmaf(wb[j], yl, wc[j])); id[j] = j; }
        d[30] = d[31] = 3e38f; id[30] = 30; id[31] = 31;
#pragma unroll
        for (int w = 1; w < 32; w *= 2)
#pragma unroll
            for (int j = 0; j + w < 32; j += 2 * w) { bool lt = d[j + w] < d[j]; d[j] = lt ? d[j + w] : d[j]; id[j] = lt ? id[j + w] : id[j]; }
        acc += id[0]; xl += 1e-3f * id[0]; yl -= 1e-3f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = xl + yl + acc;
}
// 9. same with FFMA2 over candidate pairs
__global__ void k_search_exact_f2(float* out, const float4* __restrict__ tab, float xs, float ys) {
    float2 wa[15], wb[15], wc[15];
    for (int j = 0; j < 15; ++j) { float4 t = tab[2 * j], v = tab[2 * j + 1]; wa[j] = make_float2(t.x, v.x); wb[j] = make_float2(t.y, v.y); wc[j] = make_float2(t.z, v.z); }
    float xl = threadIdx.x * xs, yl = threadIdx.x * ys; int acc = 0;
    for (int it = 0; it < ITERS; ++it) {
        float d[32]; int id[32];
        const float2 x2 = make_float2(xl, xl), y2 = make_float2(yl, yl);
#pragma unroll
        for (int j = 0; j < 15; ++j) { float2 r = ffma2(wa[j], x2, ffma2(wb[j], y2, wc[j])); d[2 * j] = r.x; d[2 * j + 1] = r.y; id[2 * j] = 2 * j; id[2 * j + 1] = 2 * j + 1; }
        d[30] = d[31] = 3e38f; id[30] = 30; id[31] = 31;
#pragma unroll
        for (int w = 1; w < 32; w *= 2)
#pragma unroll
            for (int j = 0; j + w < 32; j += 2 * w) { bool lt = d[j + w] < d[j]; d[j] = lt ? d[j + w] : d[j]; id[j] = lt ? id[j + w] : id[j]; }
        acc += id[0]; xl += 1e-3f * id[0]; yl -= 1e-3f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = xl + yl + acc;
}
// 10. packed-key search (quantised, not exact) for reference
__global__ void k_search_packed(float* out, const float4* __restrict__ tab, float xs, float ys) {
    float wa[30], wb[30], wc[30];
    for (int j = 0; j < 30; ++j) { float4 t = tab[j]; wa[j] = t.x; wb[j] = t.y; wc[j] = t.z; }
    float xl = threadIdx.x * xs, yl = threadIdx.x * ys; int acc = 0;
    for (int it = 0; it < ITERS; ++it) {
        float key[30];
#pragma unroll
        for (int j = 0; j < 30; ++j) key[j] = __int_as_float((__float_as_int(fmaf(wa[j], xl, fmaf(wb[j], yl, wc[j]))) & ~31) | j);
        float m[10];
#pragma unroll
        for (int i = 0; i < 10; ++i) m[i] = fminf(fminf(key[3 * i], key[3 * i + 1]), key[3 * i + 2]);
        float best = fminf(fminf(fminf(fminf(m[0], m[1]), m[2]), fminf(fminf(m[3], m[4]), m[5])), fminf(fminf(fminf(m[6], m[7]), m[8]), m[9]));
        int id = __float_as_int(best) & 31;
        acc += id; xl += 1e-3f * id; yl -= 1e-3f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = xl + yl + acc;
}

template <class F>
double time_ms(F launch) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    launch(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) { cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float t; cudaEventElapsedTime(&t, a, b); if (t < best) best = t; }
    return best;
}

int main() {
    int dev = 0, sm = 0, khz = 0;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    printf("SMs %d, nominal clock %.0f MHz\n", sm, khz / 1e3);
    const int threads = 256;
    float *out; int* outi; float4* tab;
    cudaMalloc(&out, (size_t)sm * 16 * threads * 4); cudaMalloc(&outi, (size_t)sm * 16 * threads * 4);
    float4 h[32]; for (int j = 0; j < 32; ++j) h[j] = make_float4(-2e-3f * j, 1e-3f * j, 4e-6f * j * j, 0);
    cudaMalloc(&tab, sizeof(h)); cudaMemcpy(tab, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int bps : {1, 2, 4, 8}) {
        const int blocks = sm * bps;
        const double warps = (double)blocks * threads / 32;
        auto rep = [&](const char* name, double ms, double instr_per_thread, double flops_per_instr) {
            const double winstr = warps * instr_per_thread;
            printf("  %-34s %8.3f ms  %7.2f Gwarp-instr/s  %6.2f TFLOP/s\n", name, ms, winstr / ms / 1e6, winstr * 32 * flops_per_instr / ms / 1e9);
        };
        printf("blocks/SM = %d (warps/SM = %d)\n", bps, bps * threads / 32);
        const double n = (double)ITERS * UNROLL * 8;
        rep("FFMA x=x*a+b (invariant a,b)", time_ms([&] { k_ffma_inv<<<blocks, threads>>>(out, 0.999f, 1e-3f); }), n, 2);
        rep("FFMA 3 distinct regs", time_ms([&] { k_ffma_3reg<<<blocks, threads>>>(out, 0.999f, 1e-3f); }), n, 2);
        rep("FFMA2 (fma.rn.f32x2)", time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 0.999f, 1e-3f); }), n, 4);
        rep("FADD+FMNMX3 pairs (2 instr)", time_ms([&] { k_min3<<<blocks, threads>>>(out, 1.f, 1e-3f); }), 2 * n, 0);
        rep("argmin node FSETP+FSEL+FADD+SEL (4)", time_ms([&] { k_argmin_node<<<blocks, threads>>>(out, outi, 1.0001f); }), 4 * n, 0);
        rep("packed node LOP3+FMUL+FMNMX (3)", time_ms([&] { k_packed_node<<<blocks, threads>>>(out, 1.0001f); }), 3 * n, 0);
        rep("MUFU.EX2", time_ms([&] { k_mufu<<<blocks, threads>>>(out); }), (double)ITERS * UNROLL * 4, 0);
        const double s = (double)ITERS;
        double t1 = time_ms([&] { k_search_exact<<<blocks, threads>>>(out, tab, 1e-4f, 2e-4f); });
        double t2 = time_ms([&] { k_search_exact_f2<<<blocks, threads>>>(out, tab, 1e-4f, 2e-4f); });
        double t3 = time_ms([&] { k_search_packed<<<blocks, threads>>>(out, tab, 1e-4f, 2e-4f); });
        printf("  search exact scalar : %8.3f ms -> %6.2f G lookups/s\n", t1, warps * 32 * s / t1 / 1e6);
        printf("  search exact FFMA2  : %8.3f ms -> %6.2f G lookups/s\n", t2, warps * 32 * s / t2 / 1e6);
        printf("  search packed keys  : %8.3f ms -> %6.2f G lookups/s\n", t3, warps * 32 * s / t3 / 1e6);
    }
    return 0;
}
