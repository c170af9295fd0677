// mont29_bench.cu -- throughput of whole 4-limb Montgomery products on B200, by formulation.
//   g4      32-bit limbs, even/odd carry chains (g4.cuh: IMAD.WIDE.U32.X, half rate)                 -- the kernel of record
//   cios29  radix 2^29, interleaved CIOS (lazy29.hpp::mont, round 2's g29.cuh)
//   ps29    radix 2^29, product scanning (81 independent IMAD.WIDE) then column-serial reduction      (lazy29.hpp::mont_ps)
//   ps29_1  the same for moduli with p = 1 (mod 2^29): m = -c_k, no multiplication by n0 or p_0       (BLS12-381 Fr)
// Each thread runs ILP independent chains x <- mont(x, y) for ITERS steps; MINB = resident CTAs per SM the kernel is
// compiled for (register cap).  Prints warp-level cycles per product per SM sub-partition next to the fma-pipe floor
// (one IMAD-class warp instruction per 2 cycles).  Results are checked against the host build of lazy29.hpp.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../thaler_study_b200/csrc -I../include -o mont29_bench mont29_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#include "g4.cuh"
#include "lazy29.hpp"

using namespace scb;
using l29::Desc29;
using l29::L9;

constexpr int ITERS = 512;

template <int MODE>
__device__ __forceinline__ L9 mont_v(const Desc29& d, const L9& a, const L9& b) {
    if constexpr (MODE == 1) return l29::mont(d, a, b);
    if constexpr (MODE == 2) return l29::mont_ps<false>(d, a, b);
    if constexpr (MODE == 3) return l29::mont_ps<true>(d, a, b);
    if constexpr (MODE == 4) return l29::mont_ps_par<false>(d, a, b);
    if constexpr (MODE == 5) return l29::mont_ps_par<true>(d, a, b);
    if constexpr (MODE == 6 || MODE == 7) {  // probe: the 81 partial products alone (6: register x register, 7: register x constant), columns xor-folded
        uint64_t c[17];
#pragma unroll
        for (int k = 0; k < 17; ++k) c[k] = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i)
#pragma unroll
            for (int j = 0; j < 9; ++j) c[i + j] += (uint64_t)a.l[j] * (MODE == 6 ? b.l[i] : d.p[i]);
        L9 r;
#pragma unroll
        for (int j = 0; j < 9; ++j) r.l[j] = ((uint32_t)c[j] ^ (uint32_t)(c[j] >> 32) ^ (j < 8 ? (uint32_t)c[j + 9] ^ (uint32_t)(c[j + 9] >> 32) : 0u)) & l29::M29;
        return r;
    }
    return a;
}

template <int MODE, int ILP, int MINB>
__global__ void __launch_bounds__(256, MINB) k29(Desc29 d, const uint32_t* in, uint32_t* out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    L9 x[ILP], y[ILP];
#pragma unroll
    for (int c = 0; c < ILP; ++c)
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            x[c].l[j] = in[j] + (j == 0 ? (uint32_t)(tid & 1023) + c : 0);
            y[c].l[j] = in[9 + j] + (j == 1 ? (uint32_t)c : 0);
        }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < ILP; ++c) x[c] = mont_v<MODE>(d, x[c], y[c]);
    }
#pragma unroll
    for (int c = 0; c < ILP; ++c)
#pragma unroll
        for (int j = 0; j < 9; ++j) out[((size_t)tid * ILP + c) * 9 + j] = x[c].l[j];
}

template <int ILP, int MINB>
__global__ void __launch_bounds__(256, MINB) kg4(FieldDesc f, const uint32_t* in, uint32_t* out) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    const g4::Arith ar(f);
    g4::W8 x[ILP], y[ILP];
#pragma unroll
    for (int c = 0; c < ILP; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            x[c].w[j] = in[18 + j] + (j == 0 ? (uint32_t)(tid & 1023) + c : 0);
            y[c].w[j] = in[26 + j] + (j == 1 ? (uint32_t)c : 0);
        }
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int c = 0; c < ILP; ++c) x[c] = ar.mul(x[c], y[c]);
    }
#pragma unroll
    for (int c = 0; c < ILP; ++c)
#pragma unroll
        for (int j = 0; j < 8; ++j) out[((size_t)tid * ILP + c) * 9 + j] = x[c].w[j];
}

// BLS12-381 Fr
static const uint64_t P64[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};

static int g_sms = 148, g_clk_khz = 1965000;
static uint32_t *d_in, *d_out;
static std::vector<uint32_t> h_in(34), h_out;

// canonical value of lazy limbs on the host: to_words + subtract p while >= p
static void canon(const L9& a, uint32_t (&w)[8]) {
    uint32_t top;
    l29::to_words(a, w, top);
    auto geq = [&]() {
        if (top) return true;
        for (int i = 7; i >= 0; --i) {
            const uint32_t pi = (uint32_t)(P64[i / 2] >> (32 * (i & 1)));
            if (w[i] != pi) return w[i] > pi;
        }
        return true;
    };
    while (geq()) {
        uint64_t borrow = 0;
        for (int i = 0; i < 8; ++i) {
            const uint64_t pi = (uint32_t)(P64[i / 2] >> (32 * (i & 1)));
            const uint64_t v = (uint64_t)w[i] - pi - borrow;
            w[i] = (uint32_t)v;
            borrow = (v >> 63) & 1;
        }
        top -= (uint32_t)borrow;
    }
}

template <int MODE, int ILP, int MINB>
static void run29(const char* name, const Desc29& d, double floor_cyc) {
    auto kern = k29<MODE, ILP, MINB>;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const int blocks = g_sms * MINB;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<blocks, 256>>>(d, d_in, d_out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<blocks, 256>>>(d, d_in, d_out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // check thread 5, chain ILP-1 against the host
    h_out.resize((size_t)blocks * 256 * ILP * 9);
    cudaMemcpy(h_out.data(), d_out, h_out.size() * 4, cudaMemcpyDeviceToHost);
    const int tid = 5, c = ILP - 1;
    L9 x, y, got;
    for (int j = 0; j < 9; ++j) {
        x.l[j] = h_in[j] + (j == 0 ? (uint32_t)(tid & 1023) + c : 0);
        y.l[j] = h_in[9 + j] + (j == 1 ? (uint32_t)c : 0);
        got.l[j] = h_out[((size_t)tid * ILP + c) * 9 + j];
    }
    for (int it = 0; it < ITERS; ++it) x = l29::mont(d, x, y);
    uint32_t wa[8], wb[8];
    canon(x, wa);
    canon(got, wb);
    bool ok = true;
    for (int i = 0; i < 8; ++i) ok &= wa[i] == wb[i];
    if (MODE >= 6) ok = true;  // probes, not products
    const double prods = (double)blocks * 256 * ILP * ITERS;
    const double warp_prods_per_smsp = prods / 32 / (g_sms * 4.0);
    const double cyc = ms * 1e-3 * g_clk_khz * 1e3 / warp_prods_per_smsp;
    printf("{\"variant\": \"%s\", \"ilp\": %d, \"ctas_per_sm\": %d, \"regs\": %d, \"ms\": %.3f, \"Gprod_per_s\": %.2f, \"cycles_per_warp_product_per_smsp\": %.0f, "
           "\"fma_pipe_floor_cycles\": %.0f, \"frac_of_floor\": %.3f, \"ok\": %s}\n",
           name, ILP, MINB, fa.numRegs, ms, prods / (ms * 1e-3) / 1e9, cyc, floor_cyc, floor_cyc / cyc, ok ? "true" : "false");
    fflush(stdout);
}

template <int ILP, int MINB>
static void rung4(const FieldDesc& f) {
    auto kern = kg4<ILP, MINB>;
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    const int blocks = g_sms * MINB;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<blocks, 256>>>(f, d_in, d_out);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<blocks, 256>>>(f, d_in, d_out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double prods = (double)blocks * 256 * ILP * ITERS;
    const double warp_prods_per_smsp = prods / 32 / (g_sms * 4.0);
    const double cyc = ms * 1e-3 * g_clk_khz * 1e3 / warp_prods_per_smsp;
    printf("{\"variant\": \"g4 (carry chains, with final conditional subtraction)\", \"ilp\": %d, \"ctas_per_sm\": %d, \"regs\": %d, \"ms\": %.3f, \"Gprod_per_s\": %.2f, "
           "\"cycles_per_warp_product_per_smsp\": %.0f, \"fma_pipe_floor_cycles\": 512, \"frac_of_floor\": %.3f}\n",
           ILP, MINB, fa.numRegs, ms, prods / (ms * 1e-3) / 1e9, cyc, 512.0 / cyc);
    fflush(stdout);
}

int main(int argc, char** argv) {
    const bool prof = argc > 1;  // any argument: one launch pair of a few variants (for ncu)
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&g_clk_khz, cudaDevAttrClockRate, 0);
    Desc29 d;
    if (!l29::make_desc(P64, 255, &d)) return 1;
    FieldDesc f{};
    for (int i = 0; i < 4; ++i) f.p[i] = P64[i];
    f.n = 4;
    f.bits = 255;
    {
        uint64_t inv = 1;
        for (int i = 0; i < 7; ++i) inv *= 2 - P64[0] * inv;
        f.inv = 0 - inv;
    }
    srand(7);
    for (int j = 0; j < 9; ++j) {
        h_in[j] = (uint32_t)rand() & (j == 8 ? 0xffffff : l29::M29);
        h_in[9 + j] = (uint32_t)rand() & (j == 8 ? 0x3fffff : l29::M29);
    }
    for (int j = 0; j < 16; ++j) h_in[18 + j] = (uint32_t)rand() * 2654435761u;
    h_in[25] &= 0x3fffffff;
    h_in[33] &= 0x3fffffff;
    cudaMalloc(&d_in, 34 * 4);
    cudaMalloc(&d_out, (size_t)g_sms * 4 * 256 * 4 * 9 * 4);
    cudaMemcpy(d_in, h_in.data(), 34 * 4, cudaMemcpyHostToDevice);
    printf("{\"sms\": %d, \"clock_khz\": %d, \"p0_is_one\": %s}\n", g_sms, g_clk_khz, d.p[0] == 1 ? "true" : "false");

    if (prof) {
        rung4<2, 2>(f);
        run29<1, 1, 4>("cios29", d, 342.0);
        run29<2, 2, 2>("ps29", d, 342.0);
        run29<6, 2, 2>("probe: 81 products R x R", d, 162.0);
        run29<7, 2, 2>("probe: 81 products R x const", d, 162.0);
        return 0;
    }
    rung4<1, 2>(f);
    rung4<2, 2>(f);
    rung4<3, 2>(f);
    rung4<2, 3>(f);
#define RUN29(MODE, NAME, FLOOR)          \
    run29<MODE, 1, 2>(NAME, d, FLOOR);    \
    run29<MODE, 2, 2>(NAME, d, FLOOR);    \
    run29<MODE, 3, 2>(NAME, d, FLOOR);    \
    run29<MODE, 1, 3>(NAME, d, FLOOR);    \
    run29<MODE, 2, 3>(NAME, d, FLOOR);    \
    run29<MODE, 1, 4>(NAME, d, FLOOR);    \
    run29<MODE, 2, 4>(NAME, d, FLOOR);
    RUN29(1, "cios29 (lazy29::mont)", 342.0)
    RUN29(2, "ps29 (product scanning, serial carries)", 342.0)
    RUN29(3, "ps29_1 (p = 1 mod 2^29)", 306.0)
    RUN29(4, "ps29 parallel carries", 342.0)
    RUN29(5, "ps29_1 parallel carries", 306.0)
    RUN29(6, "probe: 81 products R x R, xor-folded", 162.0)
    RUN29(7, "probe: 81 products R x const, xor-folded", 162.0)
    return 0;
}
