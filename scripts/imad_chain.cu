// imad_chain.cu -- issue rates of the multiply-add forms a multi-limb Montgomery product can be built from (B200).
// Round 1 measured IMAD.WIDE.U32 at full rate (17.9 T/s) and built the 4-limb product from mad.lo.cc / madc.hi.cc
// pairs, which ptxas fuses into IMAD.WIDE.U32.X with a predicate carry in and out.  This bench separates:
//   wide        mad.wide.u32 with a 64-bit addend and LOOP-INVARIANT operands: ptxas hoists the product out of the loop and
//               what is timed is IADD3 / IADD3.X pairs -- this line is NOT an IMAD.WIDE rate (round 2 misread it as one)
//   wide_var    IMAD.WIDE.U32 with a 64-bit addend, no carry, multiplicand changing every iteration (the low word of the
//               neighbouring accumulator): the real issue rate of the instruction -- one warp instruction per 4 cycles per
//               SM sub-partition, the same as the carry forms below (ncu: fmaheavy pipe 91-95 % active, profiles/r02_mont29_ncu.md)
//   wide_cout   mad.lo.cc + madc.hi (carry out of the low half only)
//   wide_x      madc.lo.cc + madc.hi.cc chains of 4 pairs, carry in AND out (round 1's chain8)
//   add64       add.cc + addc (64-bit add on the ALU pipe), shf (funnel shift), lop3
//   mix         wide + add64 + shf interleaved 2:1:1 (do the two pipes overlap?)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imad_chain imad_chain.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 2048;
constexpr int ILP = 8;

template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t x[ILP], y[ILP], z[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
        x[i] = threadIdx.x + i;
        y[i] = blockIdx.x + i;
        z[i] = x[i] ^ y[i];
        w[i] = x[i];
    }
    for (int it = 0; it < ITERS; ++it) {
        if (MODE == 2) {  // 2 chains of 4 carry-linked pairs per statement group: 8 wide products
#pragma unroll
            for (int i = 0; i < ILP; i += 4)
                asm volatile(
                    "mad.lo.cc.u32 %0, %8, %9, %0;\n\tmadc.hi.cc.u32 %1, %8, %9, %1;\n\t"
                    "madc.lo.cc.u32 %2, %8, %10, %2;\n\tmadc.hi.cc.u32 %3, %8, %10, %3;\n\t"
                    "madc.lo.cc.u32 %4, %8, %9, %4;\n\tmadc.hi.cc.u32 %5, %8, %9, %5;\n\t"
                    "madc.lo.cc.u32 %6, %8, %10, %6;\n\tmadc.hi.u32 %7, %8, %10, %7;"
                    : "+r"(x[i]), "+r"(y[i]), "+r"(x[i + 1]), "+r"(y[i + 1]), "+r"(x[i + 2]), "+r"(y[i + 2]), "+r"(x[i + 3]), "+r"(y[i + 3])
                    : "r"(z[i]), "r"(a), "r"(b));
        } else {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (MODE == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(a));
                if (MODE == 7) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 1) % ILP]), "r"(a));
                if (MODE == 8) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 1) % ILP]), "r"((uint32_t)w[(i + 3) % ILP]));
                if (MODE == 1) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(x[i]), "+r"(y[i]) : "r"(z[i]), "r"(a));
                if (MODE == 3) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
                if (MODE == 4) asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(x[i]) : "r"(y[i]));
                if (MODE == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(a), "r"(b));
                if (MODE == 6) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(z[i]), "r"(a));
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(z[i]), "r"(b));
                    asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a), "r"(b));
                    asm volatile("shf.r.wrap.b32 %0, %0, %1, 29;" : "+r"(z[i]) : "r"(y[i]));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i] + y[i] + z[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_inner, int sms, uint32_t* d) {
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
    const double ops = (double)blocks * 256 * ITERS * ILP * ops_per_inner;
    const double tops = ops / (ms * 1e-3) / 1e12;
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"op\": \"%s\", \"T_per_s\": %.3f, \"per_clk_per_SM_at_%dMHz\": %.1f, \"ms\": %.3f}\n", name, tops, clk / 1000, tops * 1e12 / sms / (clk * 1e3), ms);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t* d;
    cudaMalloc(&d, (size_t)sms * 8 * 256 * 4);
    run<0>("wide: mad.wide.u32 with loop-invariant operands -- hoisted by ptxas, times IADD3 pairs, NOT a multiply rate", 1, sms, d);
    run<7>("wide_var: IMAD.WIDE.U32 64-bit addend, no carry, register x constant, multiplicand changes every iteration (per product)", 1, sms, d);
    run<8>("wide_var_rr: the same, register x register (per product)", 1, sms, d);
    run<1>("wide_cout: mad.lo.cc + madc.hi (per product)", 1, sms, d);
    run<2>("wide_x: carry chain of 4 mad pairs, carry in and out (per product)", 1, sms, d);
    run<3>("add64: add.cc + addc (per 64-bit add)", 1, sms, d);
    run<4>("shf.r.wrap (per shift)", 1, sms, d);
    run<5>("lop3 (per op)", 1, sms, d);
    run<6>("mix: 2 wide + 1 add64 + 1 shf (per group of 4 statements)", 1, sms, d);
    return 0;
}
