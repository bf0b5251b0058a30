// ubench_f32x2.cu -- issue / pipe throughput of Blackwell's packed FP32 instructions
// (fma.rn.f32x2 -> FFMA2, add -> FADD2, mul -> FMUL2) against the scalar forms, full chip.
// Answers: does a kernel bound by instruction issue (k_corr_fft) gain from packing (re, im)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_f32x2 ubench_f32x2.cu
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi)
{
    u64 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float2 up(u64 v)
{
    float2 r;
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
    return r;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c)
{
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)), "l"(pk(c.x, c.y)));
    return up(r);
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b)
{
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return up(r);
}
__device__ __forceinline__ float2 add2(float2 a, float2 b)
{
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(pk(a.x, a.y)), "l"(pk(b.x, b.y)));
    return up(r);
}

#define NV 8
// MODE 0: scalar FFMA, 2*NV independent chains; 1: FFMA2, NV chains (same flops);
// 2: scalar radix-2 DIT butterflies (cmul in the volk fma form + add/sub); 3: the same packed
template <int MODE> __global__ void k(float2 *out, float2 w, int iters)
{
    float2 v[NV];
#pragma unroll
    for (int i = 0; i < NV; i++)
        v[i] = make_float2(threadIdx.x * 1e-3f + i, blockIdx.x * 1e-4f - i);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
            if (MODE == 0) {
#pragma unroll
                for (int i = 0; i < NV; i++) {
                    v[i].x = __fmaf_rn(v[i].x, w.x, w.y);
                    v[i].y = __fmaf_rn(v[i].y, w.x, w.y);
                }
            } else if (MODE == 1) {
#pragma unroll
                for (int i = 0; i < NV; i++)
                    v[i] = fma2(v[i], make_float2(w.x, w.x), make_float2(w.y, w.y));
            } else if (MODE == 2) {
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    float2 a = v[i], b = v[i + 1], t;
                    t.x = __fmaf_rn(w.x, b.x, -(w.y * b.y));
                    t.y = __fmaf_rn(w.x, b.y, w.y * b.x);
                    v[i] = make_float2(a.x + t.x, a.y + t.y);
                    v[i + 1] = make_float2(a.x - t.x, a.y - t.y);
                }
            } else {
#pragma unroll
                for (int i = 0; i < NV; i += 2) {
                    float2 a = v[i], b = v[i + 1];
                    float2 p = mul2(make_float2(-w.y, w.y), make_float2(b.y, b.x));
                    float2 t = fma2(make_float2(w.x, w.x), b, p);
                    v[i] = add2(a, t);
                    v[i + 1] = add2(a, make_float2(-t.x, -t.y));
                }
            }
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < NV; i++) {
        s.x += v[i].x;
        s.y += v[i].y;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> static void run(const char *name, double flop_per_iter_thread)
{
    const int blocks = 148 * 8, threads = 256, iters = 2000;
    float2 *out;
    cudaMalloc(&out, sizeof(float2) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const float2 w = make_float2(0.99999f, 1e-6f);
    k<MODE><<<blocks, threads>>>(out, w, 10);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, w, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = flop_per_iter_thread * iters * (double)blocks * threads;
    printf("%-28s %8.3f ms  %7.2f TFLOP/s\n", name, ms, flops / ms / 1e9);
    cudaFree(out);
}

int main()
{
    run<0>("FFMA  (scalar)", 16.0 * 2 * NV * 2);
    run<1>("FFMA2 (packed)", 16.0 * NV * 4);
    run<2>("butterfly scalar (8 instr)", 16.0 * (NV / 2) * 10);
    run<3>("butterfly packed (4 instr)", 16.0 * (NV / 2) * 10);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
