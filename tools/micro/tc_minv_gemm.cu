// tc_minv_gemm.cu - the tensor-core question of BASELINE.json north_star, measured.
//
// "uses tensor cores only if a batched product in the gradient path measurably wins at the stated
// tolerance".  The only GEMM-shaped step of the path is df_du = -Minv * dc_du
// (reference algorithms/_forward_dynamics_gradient.py:48-57): per state (n x n) * (n x 2n), 4 n^3
// flops = 31 % of the dense FD-gradient count at n = 30 and 30 % at n = 64.  This
// program times that product, batched over states, three ways and reports the error of each against
// a float64 product of the same float32 inputs:
//   simt   FP32 FFMA: one thread per output column, the column in registers, Minv rows from shared
//          memory as 128-bit broadcast loads (what the wide kernel does, csrc/grid_wps.cuh);
//   tf32   mma.sync.m16n8k8 TF32, one pass (10-bit mantissa inputs);
//   tf32x3 the same with the 3xTF32 split (hi*hi + hi*lo + lo*hi): FP32-class accuracy;
// each with the operands streamed from HBM once per product ("hbm") and with the product repeated
// REPS times on operands that stay in shared memory ("onchip": what a fused kernel would see, where
// Minv and dc_du never leave the SM).
// Sizes: n = 32 (Atlas padded from 30; its Minv is block-diagonal 18 + 6 + 6, which the traced kernels
// exploit and a dense GEMM cannot) and n = 64 (the 64-link chain).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tc_minv_gemm tc_minv_gemm.cu
//   ./tc_minv_gemm [states]            -> one JSON line per (n, variant, mode)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ unsigned f2tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// ---- SIMT FP32: thread = output column -------------------------------------------------------
template <int N, int REPS>
__global__ void __launch_bounds__(2 * N) simt_kernel(const float *__restrict__ Minv, const float *__restrict__ dc,
                                                     float *__restrict__ out, int states) {
    __shared__ __align__(16) float sM[N * N];                    // row-major [row][k]
    const int col = threadIdx.x;                                 // 2N columns
    for (int s = blockIdx.x; s < states; s += gridDim.x) {
        for (int e = threadIdx.x; e < N * N; e += 2 * N) sM[e] = Minv[(size_t)s * N * N + e];
        float d[N];
#pragma unroll
        for (int k = 0; k < N; k++) d[k] = dc[(size_t)s * 2 * N * N + col * N + k];
        __syncthreads();
        float keep = 0.f;
        for (int rep = 0; rep < REPS; rep++) {
            for (int r = 0; r < N; r++) {
                float acc = 0.f;
                const float4 *row = reinterpret_cast<const float4 *>(sM + r * N);
#pragma unroll
                for (int k4 = 0; k4 < N / 4; k4++) {
                    const float4 m = row[k4];
                    acc = fmaf(m.x, d[4 * k4], acc);
                    acc = fmaf(m.y, d[4 * k4 + 1], acc);
                    acc = fmaf(m.z, d[4 * k4 + 2], acc);
                    acc = fmaf(m.w, d[4 * k4 + 3], acc);
                }
                if (rep == REPS - 1) out[(size_t)s * 2 * N * N + col * N + r] = -acc;
                else keep += acc;
            }
            if (REPS > 1) d[0] += keep * 1e-30f;           // keep the repeats live and dependent
        }
        __syncthreads();
    }
}

// ---- tensor cores: mma.sync m16n8k8 TF32, SPLIT = 1 (one pass) or 3 (3xTF32) -----------------------
// CTA = N/16 warps; warp w owns rows [16w, 16w+16) of the N x 2N product.  A = -Minv (row-major, symmetric),
// B = dc (col-major n x 2n == "col" operand layout: element (k, col) at col*N + k).
template <int N, int SPLIT, int REPS>
__global__ void __launch_bounds__(32 * (N / 16)) tc_kernel(const float *__restrict__ Minv, const float *__restrict__ dc,
                                                           float *__restrict__ out, int states) {
    constexpr int PADK = N + 4;                                  // conflict-free fragment loads
    extern __shared__ float smem[];
    float *sA = smem;                                            // [N][PADK]   A(row, k)
    float *sB = smem + N * PADK;                                 // [2N][PADK]  B(k, col) at col*PADK + k
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int nthr = blockDim.x;
    for (int s = blockIdx.x; s < states; s += gridDim.x) {
        for (int e = threadIdx.x; e < N * N; e += nthr) sA[(e / N) * PADK + e % N] = -Minv[(size_t)s * N * N + e];
        for (int e = threadIdx.x; e < 2 * N * N; e += nthr) sB[(e / N) * PADK + e % N] = dc[(size_t)s * 2 * N * N + e];
        __syncthreads();
        float acc[2 * N / 8][4];
        for (int rep = 0; rep < REPS; rep++) {
#pragma unroll
            for (int j = 0; j < 2 * N / 8; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < N; k0 += 8) {
                float af[4] = {sA[(16 * warp + g) * PADK + k0 + t], sA[(16 * warp + g + 8) * PADK + k0 + t],
                               sA[(16 * warp + g) * PADK + k0 + t + 4], sA[(16 * warp + g + 8) * PADK + k0 + t + 4]};
                unsigned ah[4], al[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    ah[i] = f2tf32(af[i]);
                    al[i] = f2tf32(af[i] - __uint_as_float(ah[i]));
                }
#pragma unroll
                for (int j = 0; j < 2 * N / 8; j++) {
                    float bf[2] = {sB[(8 * j + g) * PADK + k0 + t], sB[(8 * j + g) * PADK + k0 + t + 4]};
                    unsigned bh[2] = {f2tf32(bf[0]), f2tf32(bf[1])};
                    if (SPLIT == 3) {
                        unsigned bl[2] = {f2tf32(bf[0] - __uint_as_float(bh[0])), f2tf32(bf[1] - __uint_as_float(bh[1]))};
                        mma_tf32(acc[j], al, bh);                // small terms first
                        mma_tf32(acc[j], ah, bl);
                    }
                    mma_tf32(acc[j], ah, bh);
                }
            }
            if (REPS > 1 && rep < REPS - 1) {                    // keep the repeats live and dependent
                float k = 0.f;
#pragma unroll
                for (int j = 0; j < 2 * N / 8; j++) k += acc[j][0];
                if (lane == 0) sA[(16 * warp) * PADK] += k * 1e-30f;
                __syncwarp();
            }
        }
        float *o = out + (size_t)s * 2 * N * N;
#pragma unroll
        for (int j = 0; j < 2 * N / 8; j++) {
            const int c0 = 8 * j + 2 * t, r0 = 16 * warp + g;
            o[c0 * N + r0] = acc[j][0];
            o[(c0 + 1) * N + r0] = acc[j][1];
            o[c0 * N + r0 + 8] = acc[j][2];
            o[(c0 + 1) * N + r0 + 8] = acc[j][3];
        }
        __syncthreads();
    }
}

template <int N>
__global__ void ref_kernel(const float *Minv, const float *dc, double *out, int states) {
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (idx >= (size_t)states * 2 * N * N) return;
    const size_t s = idx / (2 * N * N);
    const int e = idx % (2 * N * N), col = e / N, r = e % N;
    double acc = 0.0;
    for (int k = 0; k < N; k++) acc += (double)Minv[s * N * N + r * N + k] * (double)dc[s * 2 * N * N + col * N + k];
    out[idx] = -acc;
}

__global__ void err_kernel(const float *x, const double *ref, size_t n, double *maxabs, double *maxref) {
    double e = 0.0, m = 0.0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        e = fmax(e, fabs((double)x[i] - ref[i]));
        m = fmax(m, fabs(ref[i]));
    }
    atomicMax((unsigned long long *)maxabs, __double_as_longlong(e));     // non-negative doubles order like integers
    atomicMax((unsigned long long *)maxref, __double_as_longlong(m));
}

template <class F>
static float time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch();
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

template <int N>
static void run(int states) {
    constexpr int REPS = 8;
    const size_t nm = (size_t)states * N * N, nd = 2 * nm;
    std::vector<float> hM(nm), hD(nd);
    srand(1234 + N);
    // Minv-like: symmetric, diagonally dominant with entries spanning 3 decades (joint-space inertia
    // inverses of a chain do); dc-like: entries up to a few hundred
    for (int s = 0; s < states; s++)
        for (int r = 0; r < N; r++)
            for (int c = r; c < N; c++) {
                float v = (rand() / (float)RAND_MAX - 0.5f) * (r == c ? 0.f : 2.f) * powf(10.f, -1.5f * fabsf(r - c) / N);
                if (r == c) v = 1.f + 30.f * rand() / (float)RAND_MAX;
                hM[(size_t)s * N * N + r * N + c] = hM[(size_t)s * N * N + c * N + r] = v;
            }
    for (size_t i = 0; i < nd; i++) hD[i] = (rand() / (float)RAND_MAX - 0.5f) * 400.f;
    float *dM, *dD, *dO;
    double *dR, *dE;
    CK(cudaMalloc(&dM, nm * 4));
    CK(cudaMalloc(&dD, nd * 4));
    CK(cudaMalloc(&dO, nd * 4));
    CK(cudaMalloc(&dR, nd * 8));
    CK(cudaMalloc(&dE, 16));
    CK(cudaMemcpy(dM, hM.data(), nm * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dD, hD.data(), nd * 4, cudaMemcpyHostToDevice));
    ref_kernel<N><<<(unsigned)((nd + 255) / 256), 256>>>(dM, dD, dR, states);
    CK(cudaDeviceSynchronize());
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    constexpr size_t tc_smem = sizeof(float) * 3 * N * (N + 4);
    CK(cudaFuncSetAttribute(tc_kernel<N, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
    CK(cudaFuncSetAttribute(tc_kernel<N, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
    CK(cudaFuncSetAttribute(tc_kernel<N, 1, REPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
    CK(cudaFuncSetAttribute(tc_kernel<N, 3, REPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
    const double flops = 4.0 * N * N * N * states;
    auto report = [&](const char *variant, const char *mode, int reps_in_kernel, float ms) {
        double h[2] = {0, 0};
        CK(cudaMemset(dE, 0, 16));
        err_kernel<<<1024, 256>>>(dO, dR, nd, dE, dE + 1);
        CK(cudaMemcpy(h, dE, 16, cudaMemcpyDeviceToHost));
        printf("{\"n\": %d, \"states\": %d, \"variant\": \"%s\", \"mode\": \"%s\", \"products_per_launch\": %d, "
               "\"us_per_launch\": %.2f, \"us_per_product_batch\": %.2f, \"tflops_fp32_equiv\": %.2f, "
               "\"max_rel_err\": %.3e}\n",
               N, states, variant, mode, reps_in_kernel, ms * 1e3, ms * 1e3 / reps_in_kernel,
               flops * reps_in_kernel / (ms * 1e-3) / 1e12, h[0] / h[1]);
    };
    const int grid = states < sms * 8 ? states : sms * 8;
    CK(cudaMemset(dO, 0, nd * 4));
    report("simt_fp32", "hbm", 1, time_ms([&] { simt_kernel<N, 1><<<grid, 2 * N>>>(dM, dD, dO, states); }, 20));
    report("simt_fp32", "onchip", REPS, time_ms([&] { simt_kernel<N, REPS><<<grid, 2 * N>>>(dM, dD, dO, states); }, 10));
    CK(cudaMemset(dO, 0, nd * 4));
    report("tf32_x1", "hbm", 1, time_ms([&] { tc_kernel<N, 1, 1><<<grid, 32 * (N / 16), tc_smem>>>(dM, dD, dO, states); }, 20));
    report("tf32_x1", "onchip", REPS, time_ms([&] { tc_kernel<N, 1, REPS><<<grid, 32 * (N / 16), tc_smem>>>(dM, dD, dO, states); }, 10));
    CK(cudaMemset(dO, 0, nd * 4));
    report("tf32_x3", "hbm", 1, time_ms([&] { tc_kernel<N, 3, 1><<<grid, 32 * (N / 16), tc_smem>>>(dM, dD, dO, states); }, 20));
    report("tf32_x3", "onchip", REPS, time_ms([&] { tc_kernel<N, 3, REPS><<<grid, 32 * (N / 16), tc_smem>>>(dM, dD, dO, states); }, 10));
    cudaFree(dM); cudaFree(dD); cudaFree(dO); cudaFree(dR); cudaFree(dE);
}

int main(int argc, char **argv) {
    const int states = argc > 1 ? atoi(argv[1]) : 16384;
    run<32>(states);
    run<64>(states);
    return 0;
}
