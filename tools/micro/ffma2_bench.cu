// Microbenchmark: FFMA vs FFMA2 (packed FP32x2, sm_100) latency and throughput.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, bool PACKED>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
    float2 x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 16; r++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (PACKED) x[i] = __ffma2_rn(x[i], aa, bb);
                else { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i].x + x[i].y;
    if (s == 123.456f) out[0] = s;
}

template <int ILP, bool PACKED>
void run(const char *name, int blocks, int threads) {
    float *d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2048;
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<ILP, PACKED><<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double fmas = 2.0 * ILP * 16.0 * iters * (double)blocks * threads;       // scalar FMAs
    printf("{\"kernel\": \"%s\", \"ilp\": %d, \"blocks\": %d, \"threads\": %d, \"ms\": %.4f, \"tflops\": %.2f, "
           "\"cycles_per_warp_instr_at_1.9GHz\": %.2f}\n", name, ILP, blocks, threads, best, 2.0 * fmas / best / 1e9,
           best * 1e-3 * 1.9e9 / ((PACKED ? 1.0 : 2.0) * ILP * 16.0 * iters));
    cudaFree(d);
}

int main() {
    int sms = 148;
    // latency: one warp per SM, dependent chain (ILP 1) and ILP 4
    run<1, false>("ffma  1 warp/SM", sms, 32);  run<1, true>("ffma2 1 warp/SM", sms, 32);
    run<4, false>("ffma  1 warp/SM", sms, 32);  run<4, true>("ffma2 1 warp/SM", sms, 32);
    // 8 warps per SM (2 per SMSP), like the gradient kernels
    run<1, false>("ffma  8 warps/SM", sms, 256); run<1, true>("ffma2 8 warps/SM", sms, 256);
    run<4, false>("ffma  8 warps/SM", sms, 256); run<4, true>("ffma2 8 warps/SM", sms, 256);
    // throughput: 64 warps per SM
    run<4, false>("ffma  64 warps/SM", sms * 8, 256); run<4, true>("ffma2 64 warps/SM", sms * 8, 256);
    return 0;
}
