// grid_pipe.cuh - launch shell of the phase-split ("pipe") kernels (gridcodegenerator_b200/pipeline.py).
//
// For robots whose whole algorithm does not fit one thread (Atlas: 30 joints) the traced program
// is cut along the two independence structures of the reference algorithms: forest components
// (block-diagonal M; helpers/_topology_helpers.py:193-215 encodes the same zeros) and du-columns
// of the gradient (algorithms/_inverse_dynamics_gradient.py:189-541).  A task is one traced
// straight-line program; one WARP runs one task for 32 consecutive states (lane = state), so the
// code is divergence-free, barrier-free and shuffle-free like the thread-per-state kernels.
//   stage 0: per-component state programs (read the caller's state-major rows through a shared-
//            memory tile; write final outputs and/or the scratch words the columns need)
//   stage 1: per-column-group programs (read scratch, write output columns)
// Scratch is [tile][word][32 lanes]: every warp access is one full 128-byte line and the word
// offset is an immediate.  Output runs (a column = n contiguous floats of one state) are staged
// in the warp's shared-memory rows and flushed with consecutive lanes on consecutive words.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include "grid_tps.cuh"
#define GRID_HAS_PIPE 1

namespace GRID_NS { namespace pipe {

static std::atomic<long long> g_kernel_launches{0};     // kernels launched (an ABI call launches one per stage)
static std::atomic<long long> g_calls{0};

// Flush of one or two runs of LEN contiguous output words per state from the warp's staging rows
// (row pitch PAD).  Element e of the tile is (state e / W, word e % W); lane handles e = lane,
// lane + 32, ... so consecutive lanes store consecutive words.  (s, c) are advanced incrementally
// (no division in the loop); even LEN / PAD / offsets move float2.
template <int OUT, int LEN, int PAD, int NRUN>
__device__ __forceinline__ void flush_runs(float *__restrict__ g_tile, const float *s_warp, int off0, int off1, int cnt,
                                           int lane) {
    constexpr int W = NRUN * LEN;
    constexpr bool EVEN = (LEN % 2 == 0) && (PAD % 2 == 0) && (OUT % 2 == 0);
    if (EVEN && ((off0 | off1) & 1) == 0 && (reinterpret_cast<unsigned long long>(g_tile) & 7ull) == 0) {
        constexpr int W2 = W / 2, L2 = LEN / 2;                     // in float2 units
        constexpr int DS = 32 / W2, DC = 32 % W2;
        int s = lane / W2, c = lane - s * W2;
        const float2 *src = reinterpret_cast<const float2 *>(s_warp);
        float2 *dst = reinterpret_cast<float2 *>(g_tile);
#pragma unroll 4
        for (int k = 0; k < W2; k++) {
            if (s < cnt) dst[(long long)s * (OUT / 2) + (c < L2 ? off0 / 2 + c : off1 / 2 + c - L2)] = src[s * (PAD / 2) + c];
            s += DS;
            c += DC;
            if (c >= W2) { c -= W2; s++; }
        }
    } else {
        constexpr int DS = 32 / W, DC = 32 % W;
        int s = lane / W, c = lane - s * W;
#pragma unroll 4
        for (int k = 0; k < W; k++) {
            if (s < cnt) g_tile[(long long)s * OUT + (c < LEN ? off0 + c : off1 + c - LEN)] = s_warp[s * PAD + c];
            s += DS;
            c += DC;
            if (c >= W) { c -= W; s++; }
        }
    }
}
template <int OUT, int LEN, int PAD>
__device__ __forceinline__ void flush1(float *__restrict__ g_tile, const float *s_warp, int off, int cnt, int lane) {
    flush_runs<OUT, LEN, PAD, 1>(g_tile, s_warp, off, off, cnt, lane);
}
template <int OUT, int LEN, int PAD>
__device__ __forceinline__ void flush2(float *__restrict__ g_tile, const float *s_warp, int off0, int off1, int cnt,
                                       int lane) {
    flush_runs<OUT, LEN, PAD, 2>(g_tile, s_warp, off0, off1, cnt, lane);
}

template <class P>
struct PipeShape {
    static constexpr int IN = P::IN0 + P::IN1;
    static constexpr bool IN_LINEAR = (P::IN1 == 0) && cgcd(IN, 32) <= 2;
    static constexpr int IN_PAD = IN_LINEAR ? IN : odd_pad(IN);
    static constexpr int TILE_WORDS = (32 * IN_PAD + 3) / 4 * 4;
    static constexpr int STAGE_WORDS = (32 * P::STAGE_PAD + 3) / 4 * 4;
    // stage 0 stages the input tile and output runs; stage 1 only output runs
    static constexpr int smem_words(int stage) { return (stage == 0 ? TILE_WORDS : 0) + STAGE_WORDS; }
};

// grid = ntasks * ceil(ntiles / WARPS) CTAs; blockIdx.x = task * nblk + blk (tasks sorted by
// decreasing cost so the long ones start first); warp w of a CTA runs tile blk * WARPS + w.
// Two instruction-supply facts shape this (tools/micro/ifetch_bench.cu, profiles/r1_micro_ifetch.jsonl):
// straight-line code beyond the 32 KB SM instruction cache streams at ~0.3 instructions/cycle per
// warp, and an SM whose warps sit at DIFFERENT places of such a program is capped near 1.0 IPC
// (each warp pulls its own stream from the GPC cache / L2), whereas warps that run the SAME lines
// at about the same time share every fetched line (2.7 IPC with 8 warps).  Hence (a) task-major
// order - all resident warps run one program (tile-major order is 3.2x slower: eight ~100 KB
// programs thrash the cache, no_instruction stalls 37 per issue); (b) the WARPS warps of a CTA
// start the same program together and, having no data-dependent control flow, stay in step.
template <class P, int STAGE>
__global__ void __launch_bounds__(32 * P::WARPS, STAGE == 0 ? P::MINB0 : P::MINB1)
pipe_kernel(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0, const float *__restrict__ d_in1,
            float *__restrict__ scratch, int num_states, int ntiles, int nblk, float gravity) {
    using S = PipeShape<P>;
    extern __shared__ float smem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int task = blockIdx.x / nblk, blk = blockIdx.x - task * nblk;
    const int tile = blk * P::WARPS + warp;
    if (tile >= ntiles) return;                      // no CTA-wide barriers anywhere below
    float *smem = smem_all + warp * S::smem_words(STAGE);
    const long long first = (long long)tile * 32;
    const int cnt = min(32, num_states - (int)first);
    float *s_warp = smem + (STAGE == 0 ? S::TILE_WORDS : 0);
    if (STAGE == 0) {
        const float *src0 = d_in0 + first * (long long)stride0;
        if (S::IN_LINEAR && stride0 == P::IN0 && aligned16(src0)) {
            warp_copy_g2s(smem, src0, cnt * P::IN0, lane);
        } else {
            tile_load<P::IN0, S::IN_PAD>(smem, 0, d_in0, first, stride0, cnt, lane);
            tile_load<P::IN1, S::IN_PAD>(smem, P::IN0, d_in1, first, P::IN1, cnt, lane);
        }
        __syncwarp();
    }
    // lanes past the end of a ragged tile recompute the last valid state; their scratch lane is
    // private (scratch is allocated in whole tiles) and the flushes only write cnt states
    const int src = min(lane, cnt - 1);
    float *sc = scratch + (long long)tile * (P::SCRATCH_WORDS * 32) + lane;
    P::template run<STAGE>(task, smem + src * S::IN_PAD, sc, sc, s_warp + lane * P::STAGE_PAD,
                           d_out + first * P::OUT, cnt, lane, s_warp, gravity);
}

template <class P, int STAGE>
cudaError_t pipe_stage_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, float *scratch,
                              int num_states, float gravity, cudaStream_t stream) {
    using S = PipeShape<P>;
    constexpr int ntasks = STAGE == 0 ? P::NTASKS0 : P::NTASKS1;
    if (ntasks == 0) return cudaSuccess;
    auto kern = pipe_kernel<P, STAGE>;
    constexpr size_t smem_bytes = sizeof(float) * S::smem_words(STAGE) * P::WARPS;
    static bool configured = false;             // benign race: idempotent attribute set
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    const int ntiles = (num_states + 31) / 32;
    const int nblk = (ntiles + P::WARPS - 1) / P::WARPS;
    kern<<<(unsigned)(ntasks * (long long)nblk), 32 * P::WARPS, smem_bytes, stream>>>(
        d_out, d_in0, stride0, d_in1, scratch, num_states, ntiles, nblk, gravity);
    g_kernel_launches.fetch_add(1);
    return cudaGetLastError();
}

// Both stages on `stream`.  The scratch array comes from the stream-ordered allocator (no
// library state, safe with concurrent callers on different streams); its pool keeps the memory
// between calls, so the allocation is a free-list hit after the first launch.
template <class P>
cudaError_t pipe_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, int num_states,
                        float gravity, cudaStream_t stream) {
    if (num_states <= 0) return cudaSuccess;
    g_calls.fetch_add(1);
    float *scratch = nullptr;
    cudaError_t e;
    if (P::SCRATCH_WORDS > 0) {
        static bool pool_ready = false;
        if (!pool_ready) {
            int dev = 0;
            cudaGetDevice(&dev);
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            pool_ready = true;
        }
        const size_t ntiles = (size_t)(num_states + 31) / 32;
        e = cudaMallocAsync((void **)&scratch, ntiles * P::SCRATCH_WORDS * 32 * sizeof(float), stream);
        if (e != cudaSuccess) return e;
    }
    e = pipe_stage_launch<P, 0>(d_out, d_in0, stride0, d_in1, scratch, num_states, gravity, stream);
    if (e == cudaSuccess) e = pipe_stage_launch<P, 1>(d_out, d_in0, stride0, d_in1, scratch, num_states, gravity, stream);
    if (scratch) {
        cudaError_t e2 = cudaFreeAsync(scratch, stream);
        if (e == cudaSuccess) e = e2;
    }
    return e;
}

}}  // namespace GRID_NS::pipe
