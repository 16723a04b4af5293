// ifetch_bench.cu - how fast can warps stream STRAIGHT-LINE code on sm_100a?
//
// The traced kernels of this repository are thousands of instructions of branch-free FP32 code per
// warp (iiwa14 FD gradient: 8.7 k instructions = 140 KB), and ncu reports them bound by
// `no_instruction` stalls.  This microbenchmark isolates that: a body of B independent-chain FFMAs
// (8 accumulators, so dependencies never bind), executed `iters` times, with W warps per SM that
// either start together or are de-phased by a busy-wait so that they sit at different places of
// the body (dephase > 0), or whole SMs are de-phased against each other (dephase < 0).  Output: JSON lines with the achieved instructions/cycle per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ifetch_bench ifetch_bench.cu && ./ifetch_bench
#include <cuda_runtime.h>
#include <cstdio>

#define F1(k) x[(k) & 7] = fmaf(x[(k) & 7], a, b);
#define F8(k) F1(k) F1(k + 1) F1(k + 2) F1(k + 3) F1(k + 4) F1(k + 5) F1(k + 6) F1(k + 7)
#define F64(k) F8(k) F8(k + 8) F8(k + 16) F8(k + 24) F8(k + 32) F8(k + 40) F8(k + 48) F8(k + 56)
#define F256(k) F64(k) F64(k + 64) F64(k + 128) F64(k + 192)
#define F1K(k) F256(k) F256(k + 256) F256(k + 512) F256(k + 768)
#define F4K(k) F1K(k) F1K(k + 1024) F1K(k + 2048) F1K(k + 3072)

template <int B>
struct Body;
template <> struct Body<256> { static __device__ __forceinline__ void run(float *x, float a, float b) { F256(0) } };
template <> struct Body<1024> { static __device__ __forceinline__ void run(float *x, float a, float b) { F1K(0) } };
template <> struct Body<2048> { static __device__ __forceinline__ void run(float *x, float a, float b) { F1K(0) F1K(0) } };
template <> struct Body<4096> { static __device__ __forceinline__ void run(float *x, float a, float b) { F4K(0) } };
template <> struct Body<8192> { static __device__ __forceinline__ void run(float *x, float a, float b) { F4K(0) F4K(0) } };
template <> struct Body<16384> { static __device__ __forceinline__ void run(float *x, float a, float b) { F4K(0) F4K(0) F4K(0) F4K(0) } };

template <int B>
__global__ void __launch_bounds__(512) k(float *out, int iters, int dephase, float a, float b) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 1e-3f + i;
    const int warp = threadIdx.x >> 5;
    if (dephase > 0) {             // de-phase the warps of an SM
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)dephase * warp) { }
    } else if (dephase < 0) {      // warps of an SM start together, SMs are de-phased against each other
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)(-dephase) * (blockIdx.x % 37)) { }
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) Body<B>::run(x, a, b);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}

template <int B>
void bench(int warps, int dephase, float *d) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const int iters = (int)(2000000LL / B) + 1;             // ~2 M instructions per warp
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        k<B><<<sms, 32 * warps>>>(d, iters, dephase, 1.0001f, 1e-7f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)iters * B * warps;          // per SM
    const double cycles = ms * 1e-3 * clk_khz * 1e3;
    printf("{\"body_instr\": %d, \"body_kb\": %d, \"warps_per_sm\": %d, \"dephase_cycles_per_warp\": %d, \"ms\": %.3f, "
           "\"ipc_per_sm\": %.3f, \"ipc_per_warp\": %.3f}\n",
           B, B * 16 / 1024, warps, dephase, ms, instr / cycles, instr / cycles / warps);
    fflush(stdout);
}

int main() {
    float *d;
    cudaMalloc(&d, 4);
    const int warps[] = {1, 4, 8, 16};
    const int deph[] = {0, 3000, -3000};
    for (int w : warps)
        for (int dp : deph) {
            if (w == 1 && dp) continue;
            bench<256>(w, dp, d);
            bench<1024>(w, dp, d);
            bench<2048>(w, dp, d);
            bench<4096>(w, dp, d);
            bench<8192>(w, dp, d);
            bench<16384>(w, dp, d);
        }
    return 0;
}
