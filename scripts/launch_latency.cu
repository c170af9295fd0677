// launch_latency.cu -- what one synchronous kernel call costs on this box: launch + completion wait for an empty kernel,
// by parameter size, wait method (cudaStreamSynchronize under the default / spin / blocking-sync device flags, or the
// host polling a flag the kernel writes into mapped pinned memory) -- the fixed cost under every one-launch C-ABI call
// (scb_mle_evaluate, scb_prover_round, ...).  Build: nvcc -O3 -o launch_latency launch_latency.cu
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

struct Big { uint64_t w[160]; };  // 1280 bytes, the size of PointArg
__global__ void k_small(volatile uint64_t* flag, uint64_t seq) { if (flag && threadIdx.x == 0) { *flag = seq; __threadfence_system(); } }
__global__ void k_big(Big b, volatile uint64_t* flag, uint64_t seq) { if (flag && threadIdx.x == 0) { *flag = seq + b.w[7]; __threadfence_system(); } }

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;  // 0 default flags, 1 spin, 2 blocking sync
    if (mode == 1) cudaSetDeviceFlags(cudaDeviceScheduleSpin | cudaDeviceMapHost);
    else if (mode == 2) cudaSetDeviceFlags(cudaDeviceScheduleBlockingSync | cudaDeviceMapHost);
    else cudaSetDeviceFlags(cudaDeviceMapHost);
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    volatile uint64_t* flag;
    cudaHostAlloc((void**)&flag, 64, cudaHostAllocMapped);
    *flag = 0;
    Big b{};
    const int reps = 2000;
    auto bench = [&](const char* name, auto&& fn) {
        for (int i = 0; i < 100; ++i) fn(i);
        cudaStreamSynchronize(s);
        const double t0 = now_us();
        for (int i = 0; i < reps; ++i) fn(1000 + i);
        const double t1 = now_us();
        printf("{\"flags\": %d, \"case\": \"%s\", \"us_per_call\": %.2f}\n", mode, name, (t1 - t0) / reps);
    };
    bench("small params, cudaStreamSynchronize", [&](int i) { k_small<<<1, 32, 0, s>>>(nullptr, i); cudaStreamSynchronize(s); });
    bench("1280-byte params, cudaStreamSynchronize", [&](int i) { k_big<<<1, 32, 0, s>>>(b, nullptr, i); cudaStreamSynchronize(s); });
    bench("small params, spin on cudaStreamQuery", [&](int i) { k_small<<<1, 32, 0, s>>>(nullptr, i); while (cudaStreamQuery(s) == cudaErrorNotReady) {} });
    bench("small params, host polls a mapped flag", [&](int i) { k_small<<<1, 32, 0, s>>>(flag, (uint64_t)i + 5); while (*flag != (uint64_t)i + 5) {} });
    bench("1280-byte params, host polls a mapped flag", [&](int i) { k_big<<<1, 32, 0, s>>>(b, flag, (uint64_t)i + 5); while (*flag != (uint64_t)i + 5) {} });
    bench("grid of 1184 CTAs x 256, cudaStreamSynchronize", [&](int i) { k_small<<<1184, 256, 0, s>>>(nullptr, i); cudaStreamSynchronize(s); });
    return 0;
}
