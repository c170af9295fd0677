// imad_peak.cu -- measures the B200 integer multiply-add issue rates that bound the Montgomery kernels
// ("No integer-pipe peak is provided -- the build must measure one").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_peak imad_peak.cu ; run on the GPU box.
// Each kernel runs ILP independent chains per thread so that latency is hidden and the pipe rate shows.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[ILP], y[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = threadIdx.x + i;
        y[i] = blockIdx.x + i;
        w[i] = x[i];
    }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
            if (MODE == 2) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(a));
            if (MODE == 3)  // the carry-chain pair used by mont32.cuh
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
            if (MODE == 4) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
            if (MODE == 5) asm volatile("mad.lo.u64 %0, %0, %1, %2;" : "+l"(w[i]) : "l"((uint64_t)a | 1ull << 40), "l"((uint64_t)b));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(const char* name, int ops_per_stmt, int sms, uint32_t* d) {
    const int blocks = sms * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 3, 5);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, 3, 5);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256 * ITERS * ILP * ops_per_stmt;
    const double tops = ops / (ms * 1e-3) / 1e12;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"op\": \"%s\", \"Tops_per_s\": %.3f, \"lanes_per_clk_per_SM_at_%dMHz\": %.1f, \"ms\": %.3f}\n", name, tops, clk / 1000,
           tops * 1e12 / sms / (clk * 1e3), ms);
    return tops;
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* d;
    cudaMalloc(&d, (size_t)sms * 8 * 256 * 4);
    run<0>("mad.lo.u32 (IMAD)", 1, sms, d);
    run<1>("mad.hi.u32 (IMAD.HI)", 1, sms, d);
    run<2>("mad.wide.u32 (IMAD.WIDE)", 1, sms, d);
    run<3>("mad.lo.cc + madc.hi.cc pair (counted as 2 ops)", 2, sms, d);
    run<4>("add.cc + addc pair (counted as 2 ops)", 2, sms, d);
    run<5>("mad.lo.u64 (counted as 1 op)", 1, sms, d);
    return 0;
}
