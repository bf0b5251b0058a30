// ubench_latency.cu -- dependent-issue latency of the instructions on the per-channel
// recurrences' critical paths (k_msk, k_nco_phase), one warp alone on an SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench_latency ubench_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256

template <int OP>
__global__ void chain(float *out, int *iout, float seed, int iters, long long *cycles)
{
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 32)
        sm[i] = (float)((i * 7 + 3) & 1023);
    __syncwarp();
    float x = seed + threadIdx.x * 1e-3f, y = seed * 0.5f;
    int k = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < REP; r++) {
            if (OP == 0)
                x = x + y; // FADD
            else if (OP == 1)
                x = x * y; // FMUL
            else if (OP == 2)
                x = __fmaf_rn(x, y, y); // FFMA
            else if (OP == 3)
                x = (float)__float2int_rn(x) + 0.25f; // F2I + I2F + FADD
            else if (OP == 4)
                x = floorf(x) + 0.75f; // FRND + FADD
            else if (OP == 5)
                k = ((int)sm[k & 1023]) ; // LDS + F2I
            else if (OP == 6)
                k = __reduce_or_sync(0xffffffffu, k) + 1; // REDUX
            else if (OP == 7)
                k = __shfl_xor_sync(0xffffffffu, k, 1) + 1; // SHFL
            else if (OP == 8)
                k = k * 3 + 1; // IMAD
            else if (OP == 9)
                x = fabsf(x + 3.0f) - fabsf(x - 3.0f); // 2 FADD parallel + FADD
            else if (OP == 10)
                k = __any_sync(0xffffffffu, k > r) ? k + 1 : k - 1; // VOTE + select
            else if (OP == 11)
                x = (x > y) ? x - y : x + y; // FSETP + FSEL-ish
        }
    }
    long long t1 = clock64();
    out[threadIdx.x] = x;
    iout[threadIdx.x] = k;
    if (threadIdx.x == 0)
        cycles[0] = t1 - t0;
}

int main()
{
    float *out;
    int *iout;
    long long *cyc;
    cudaMalloc(&out, 128);
    cudaMalloc(&iout, 128);
    cudaMalloc(&cyc, 8);
    const char *names[] = { "FADD", "FMUL", "FFMA", "F2I+I2F+FADD", "FRND.FLOOR+FADD", "LDS+F2I", "REDUX.OR+IADD",
                            "SHFL+IADD", "IMAD", "clip core (2 FADD || + FADD)", "VOTE.ANY+SEL", "FSETP+select" };
    const int iters = 64;
#define RUN(OP)                                                                   \
    {                                                                             \
        chain<OP><<<1, 32>>>(out, iout, 1.0001f, iters, cyc);                     \
        chain<OP><<<1, 32>>>(out, iout, 1.0001f, iters, cyc);                     \
        cudaDeviceSynchronize();                                                  \
        long long h;                                                              \
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                           \
        printf("%-32s %.2f cycles per link\n", names[OP], (double)h / (iters * REP)); \
    }
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
