// Issue rate of the integer instructions the DP kernels are built from, per SM sub-partition (warp instructions per
// cycle), measured with eight independent chains per thread and enough warps to hide the latency.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx_rate dpx_rate.cu && ./dpx_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ unsigned op(unsigned a, unsigned b, unsigned c)
{
    unsigned r;
    if (OP == 0) asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                                  // IADD3
    if (OP == 1) asm volatile("max.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                                  // VIMNMX.U32
    if (OP == 2) asm volatile("max.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                                // VIMNMX.U16x2
    if (OP == 3) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; max.u32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));        // VIADDMNMX.U32
    if (OP == 4) asm volatile("{.reg .b32 t; add.u16x2 t, %1, %2; max.u16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));    // VIADDMNMX.U16x2
    if (OP == 5) asm volatile("{.reg .u32 t; max.u32 t, %1, %2; max.u32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));        // VIMNMX3.U32
    if (OP == 6) asm volatile("{.reg .b32 t; max.u16x2 t, %1, %2; max.u16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));    // VIMNMX3.U16x2
    if (OP == 7) asm volatile("prmt.b32 %0, %1, %2, 0xFDB9;" : "=r"(r) : "r"(a), "r"(b));                          // PRMT
    if (OP == 8) asm volatile("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));                // LOP3
    if (OP == 9) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));                   // IMAD
    if (OP == 10) asm volatile("add.u16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                               // add.u16x2
    if (OP == 11) asm volatile("{.reg .pred p; setp.gt.u32 p, %1, %3; selp.b32 %0, %1, %2, p;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));   // ISETP + SEL
    return r;
}

template <int OP>
__global__ void rate(unsigned* out, int iters, unsigned seed)
{
    unsigned x[8];
    for (int k = 0; k < 8; k++) x[k] = seed + threadIdx.x * 8 + k;
    const unsigned b = seed ^ 0x10203u, c = seed + 77u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = op<OP>(x[k], b, c);
    }
    unsigned s = 0;
    for (int k = 0; k < 8; k++) s ^= x[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
static void run(const char* name, unsigned* d, int sms, double ghz)
{
    const int iters = 4096, blocks = sms * 4, threads = 512;
    rate<OP><<<blocks, threads>>>(d, 16, 1);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate<OP><<<blocks, threads>>>(d, iters, 3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winst = (double)blocks * (threads / 32) * iters * 32.0;
    const double per_clk_smsp = winst / (ms * 1e-3 * ghz * 1e9) / (sms * 4);
    printf("{\"op\": \"%s\", \"ms\": %.3f, \"warp_inst_per_clk_per_smsp\": %.3f, \"thread_ops_T_per_s\": %.2f}\n", name, ms, per_clk_smsp, winst * 32 / (ms * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    unsigned* d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 4 * 512 * 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
    run<0>("add.u32 (IADD3)", d, p.multiProcessorCount, ghz);
    run<1>("max.u32 (VIMNMX.U32)", d, p.multiProcessorCount, ghz);
    run<2>("max.u16x2 (VIMNMX.U16x2)", d, p.multiProcessorCount, ghz);
    run<3>("add+max u32 (VIADDMNMX.U32)", d, p.multiProcessorCount, ghz);
    run<4>("add+max u16x2 (VIADDMNMX.U16x2)", d, p.multiProcessorCount, ghz);
    run<5>("max3 u32 (VIMNMX3.U32)", d, p.multiProcessorCount, ghz);
    run<6>("max3 u16x2 (VIMNMX3.U16x2)", d, p.multiProcessorCount, ghz);
    run<7>("prmt (PRMT)", d, p.multiProcessorCount, ghz);
    run<8>("lop3 (LOP3)", d, p.multiProcessorCount, ghz);
    run<9>("mad.lo (IMAD)", d, p.multiProcessorCount, ghz);
    run<10>("add.u16x2", d, p.multiProcessorCount, ghz);
    run<11>("selp (ISETP+SEL)", d, p.multiProcessorCount, ghz);
    return 0;
}
