// int_rate.cu -- issue rate of the integer instructions the moment-form vertical pass could be built from, per SM and
// clock on this GPU: IDP.2A / IDP.4A (dot products), IMAD, IADD3, PRMT, LOP3, FFMA2.  Each kernel runs ILP independent
// chains per thread, 8 warps per SMSP, long enough to hide launch cost.  nvcc -arch=sm_100a -O3 -o int_rate int_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum Op { IDP2A, IDP4A, IMAD, IADD, PRMT, LOP, FFMA2, N_OPS };
static const char *names[N_OPS] = {"IDP.2A", "IDP.4A", "IMAD", "IADD3", "PRMT", "LOP3", "FFMA2"};

template <int OP> __global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t a, uint32_t b, int iters)
{
    constexpr int ILP = 8;
    uint32_t x[ILP];
    float2 f[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = threadIdx.x * 7 + i; f[i] = make_float2((float)i, (float)threadIdx.x); }
    const float2 fa = make_float2(__uint_as_float(a), __uint_as_float(b));
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (OP == IDP2A) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == IDP4A) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == IMAD) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(a));
                if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x2103;" : "+r"(x[i]) : "r"(a));
                if (OP == LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(*(unsigned long long *)&f[i]) : "l"(*(const unsigned long long *)&fa));
            }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i] + __float_as_uint(f[i].x) + __float_as_uint(f[i].y);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> static void run(uint32_t *out, int sms, int mhz)
{
    const int iters = 4096, blocks = sms, threads = 1024;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, threads>>>(out, 3, 5, 16);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(out, 3, 5, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)iters * 4 * 8 * (threads / 32); // warp instructions per SM
    const double cycles = ms * 1e-3 * mhz * 1e6;
    printf("{\"op\": \"%s\", \"warp_inst_per_clk_per_sm\": %.3f, \"ms\": %.3f}\n", names[OP], inst / cycles, ms);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int mhz = p.clockRate / 1000;
    uint32_t *out;
    cudaMalloc(&out, (size_t)p.multiProcessorCount * 1024 * 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz_assumed\": %d}\n", p.name, p.multiProcessorCount, mhz);
    run<IDP2A>(out, p.multiProcessorCount, mhz);
    run<IDP4A>(out, p.multiProcessorCount, mhz);
    run<IMAD>(out, p.multiProcessorCount, mhz);
    run<IADD>(out, p.multiProcessorCount, mhz);
    run<PRMT>(out, p.multiProcessorCount, mhz);
    run<LOP>(out, p.multiProcessorCount, mhz);
    run<FFMA2>(out, p.multiProcessorCount, mhz);
    return cudaGetLastError() != cudaSuccess;
}
