// Issue-rate microbenchmarks behind the depthwise kernel's design: FFMA vs FFMA2 (fma.rn.f32x2) vs
// FHFMA (fma.rn.f32.f16) vs half->float converts.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
constexpr int ITERS = 4096, NACC = 16;

__global__ void k_ffma(float *out, float a, float b) {
    float acc[NACC];
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float *out, float a, float b) {
    unsigned long long acc[NACC];
    float2 av = make_float2(a, a), bv = make_float2(b, b);
    unsigned long long A = *reinterpret_cast<unsigned long long *>(&av), B = *reinterpret_cast<unsigned long long *>(&bv);
    for (int i = 0; i < NACC; ++i) { float2 t = make_float2(threadIdx.x + i, i); acc[i] = *reinterpret_cast<unsigned long long *>(&t); }
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(A), "l"(B));
    float s = 0; for (int i = 0; i < NACC; ++i) { float2 t = *reinterpret_cast<float2 *>(&acc[i]); s += t.x + t.y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int HI>
__global__ void k_fhfma(float *out, uint32_t xa, uint32_t xb) {
    float acc[NACC];
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x + i;
    __half2 a = *reinterpret_cast<__half2 *>(&xa), b = *reinterpret_cast<__half2 *>(&xb);
    unsigned short al = __half_as_ushort(HI ? a.y : a.x), bl = __half_as_ushort(HI ? b.y : b.x);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) asm volatile("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(acc[i]) : "h"(al), "h"(bl));
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_cvt(float *out, const uint32_t *in) {
    uint32_t h[NACC];
    for (int i = 0; i < NACC; ++i) h[i] = in[threadIdx.x + i];
    float s = 0;
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            float2 f = __half22float2(*reinterpret_cast<__half2 *>(&h[i]));
            s += f.x; s += f.y;        // 2 cvt + 2 fadd
            h[i] += 0x00010001u;
        }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_hfma2(float *out, uint32_t xa, uint32_t xb) {
    __half2 acc[NACC];
    __half2 a = *reinterpret_cast<__half2 *>(&xa), b = *reinterpret_cast<__half2 *>(&xb);
    for (int i = 0; i < NACC; ++i) acc[i] = __floats2half2_rn(threadIdx.x * 0.001f, i);
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = __hfma2(acc[i], a, b);
    float s = 0; for (int i = 0; i < NACC; ++i) s += __low2float(acc[i]) + __high2float(acc[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float *out; cudaMalloc(&out, 148 * 8 * 1024 * 4); uint32_t *in; cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0x3c, 4096 * 4);
    const int grid = 148 * 4, block = 512;
    const double lane_ops = (double)grid * block * ITERS * NACC;
    __half2 ha = __floats2half2_rn(1.0001f, 0.9999f); uint32_t ua = *reinterpret_cast<uint32_t *>(&ha);
    auto rep = [&](const char *n, float ms, double per) {
        printf("%-28s %8.3f ms  %7.2f T lane-instr/s  (%.1f per clk per SM @1.965 GHz)  %7.2f T elem-ops/s\n", n, ms,
               lane_ops / ms / 1e9, lane_ops / ms / 1e9 * 1e12 / 148 / 1.965e9 / 1e0 / 1e0 * 1e-0 / 1.0 / 1e0 / 1.0 / 1.0 / 1.0 * 1.0 / 1.0, lane_ops * per / ms / 1e9);
    };
    rep("FFMA", timeit([&] { k_ffma<<<grid, block>>>(out, 1.0001f, 0.5f); }), 1);
    rep("FFMA2 (f32x2)", timeit([&] { k_ffma2<<<grid, block>>>(out, 1.0001f, 0.5f); }), 2);
    rep("FHFMA lo (f32 += f16*f16)", timeit([&] { k_fhfma<0><<<grid, block>>>(out, ua, ua); }), 1);
    rep("FHFMA hi", timeit([&] { k_fhfma<1><<<grid, block>>>(out, ua, ua); }), 1);
    rep("HFMA2", timeit([&] { k_hfma2<<<grid, block>>>(out, ua, ua); }), 2);
    rep("cvt h2->f2 + 2 FADD + IADD", timeit([&] { k_cvt<<<grid, block>>>(out, in); }), 1);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
