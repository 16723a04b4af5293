// grid_cps.cuh - latency kernels: one lane per du-column, G lanes per state, straight-line.
//
// For small batches (the N = 128 knot-point case of BASELINE.json) one thread per state leaves
// the critical path at the full program length (~8.5k instructions for the iiwa14 FD gradient).
// Here the 2n gradient columns of a state are spread over G = 16 or 32 lanes that all run the
// SAME traced program (algorithms.trace_column_program): the column-independent part (RNEA,
// Minv, qdd) is recomputed by every lane - redundant work, zero communication - and the column
// part is written lane-uniformly with 0/1 masks, so there is no divergence, no shuffle, no
// shared memory and no barrier.  Critical path: ~3.1k instructions.  The reference spends ~111
// block barriers per state on the same computation.
#pragma once
#include <cuda_runtime.h>

namespace GRID_NS {

template <class C, int G>
__global__ void __launch_bounds__(32)
cps_kernel(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0, const float *__restrict__ d_in1,
           int num_states, float gravity) {
    constexpr int SPW = 32 / G;                      // states per warp
    const int lane = threadIdx.x & 31;
    const int sub = lane / G, col = lane % G;
    const long long groups = ((long long)num_states + SPW - 1) / SPW;
    for (long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const long long st = grp * SPW + sub;
        const bool valid = st < num_states && col < C::COLS;
        const long long s2 = st < num_states ? st : num_states - 1;     // keep loads in range
        C::eval(d_in0 + s2 * stride0, d_in1 ? d_in1 + s2 * C::IN1 : nullptr,
                d_out + s2 * (long long)(C::COLS * C::ROWS) + (valid ? col : 0) * C::ROWS, valid ? col : C::COLS,
                valid, gravity);
    }
}

template <class C, int G>
cudaError_t cps_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, int num_states,
                       float gravity, cudaStream_t stream) {
    if (num_states <= 0) return cudaSuccess;
    constexpr int SPW = 32 / G;
    long long groups = ((long long)num_states + SPW - 1) / SPW;
    static int caps[kMaxDevices];
    int dev = 0;
    if (cudaError_t e = current_device(dev)) return e;
    int &cap = caps[dev];
    if (cap == 0) {
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cps_kernel<C, G>, 32, 0);
        cap = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int blocks = groups < cap ? (int)groups : cap;
    cps_kernel<C, G><<<blocks, 32, 0, stream>>>(d_out, d_in0, stride0, d_in1, num_states, gravity);
    return cudaGetLastError();
}

}  // namespace GRID_NS
