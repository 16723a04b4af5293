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
#include <cstdlib>
#include <cstring>
#include "grid_tps.cuh"
#define GRID_HAS_PIPE 1

namespace GRID_NS { namespace pipe {

static std::atomic<long long> g_kernel_launches{0};     // kernels launched (an ABI call launches one per stage)
static std::atomic<long long> g_calls{0};

// Scratch words are written and read inside one kernel (by different SMs) in the fused kernel, so
// they must not go through the non-coherent path: ld.global.cg reads them at L2.
__device__ __forceinline__ float ldsc(const float *p) { return __ldcg(p); }
// two-states-per-lane programs (P::X2): the lane's two consecutive states of a 64-state scratch row in one load
__device__ __forceinline__ float2 ldsc2(const float *p) { return __ldcg(reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }   // folds into an operand modifier
__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Flush of one or two runs of LEN contiguous output words per state from the warp's staging rows
// (row pitch PAD).  Each lane owns one (float or float2) word position c of the run pair and, when
// the pair is short, one of K = 32 / Wp states per step: all index arithmetic is hoisted out of the
// loop, whose body is one shared load, one global store and one predicate with immediate offsets
// (the element-wise mapping it replaces cost 9 instructions per word: 19 % of the iiwa14 kernel).
// Cache policy of the output stores: ".cs" (streaming, evict-first).  The outputs are written once and never read by
// the kernels, while the scratch words and the INSTRUCTION STREAM live in the same L2: with default write-back stores
// the 472 MB of an Atlas FD gradient push them out.  Measured (profiles/r2_atlas_output_store_policy.jsonl, Atlas FD
// gradient): 735 vs 749 us at 65 536 states, 212.5 vs 222.8 at 16 384, 110.7 vs 128.0 at 8 192 (the per-GPU shard of
// a strongly scaled job: -14 %); ".wt" changes nothing.  -DGRID_PIPE_ST_POLICY=\"\" restores write-back.
#ifndef GRID_PIPE_ST_POLICY
#define GRID_PIPE_ST_POLICY ".cs"
#endif
// predicated global store without a branch: `if (k < left) *p = v`
__device__ __forceinline__ void st_if_lt(float *p, float v, int k, int left) {
    asm volatile("{ .reg .pred q; setp.lt.s32 q, %2, %3; @q st.global" GRID_PIPE_ST_POLICY ".f32 [%0], %1; }" ::"l"(p), "f"(v), "r"(k), "r"(left)
                 : "memory");
}
__device__ __forceinline__ void st_if_lt(float2 *p, float2 v, int k, int left) {
    asm volatile("{ .reg .pred q; setp.lt.s32 q, %3, %4; @q st.global" GRID_PIPE_ST_POLICY ".v2.f32 [%0], {%1, %2}; }" ::"l"(p), "f"(v.x),
                 "f"(v.y), "r"(k), "r"(left)
                 : "memory");
}

template <typename T, int OUT, int LEN, int PAD, int NRUN>      // OUT, LEN, PAD, offsets in units of T
__device__ __forceinline__ void flush_rows(T *__restrict__ g_tile, const T *s_warp, int off0, int off1, int cnt, int lane) {
    constexpr int W = NRUN * LEN;
    constexpr int Wp = W <= 1 ? 1 : W <= 2 ? 2 : W <= 4 ? 4 : W <= 8 ? 8 : W <= 16 ? 16 : 32;
    constexpr int K = 32 / Wp;                       // states per step
    static_assert(W <= 32, "flush_rows handles at most 32 words per state");
    const int ds = lane / Wp, c0 = lane - ds * Wp;
    const bool lane_on = c0 < W;
    const int c = lane_on ? c0 : 0;                  // idle lanes read word 0 (in bounds) and store nothing
    const T *src = s_warp + ds * PAD + c;
    T *dst = g_tile + (long long)ds * OUT + (c < LEN ? off0 + c : off1 + c - LEN);
    const int left = lane_on ? cnt - ds : 0;         // this lane stores while s0 < left
#pragma unroll
    for (int s0 = 0; s0 < 32; s0 += K) st_if_lt(dst + (long long)s0 * OUT, src[s0 * PAD], s0, left);
}

template <int OUT, int LEN, int PAD, int NRUN>
__device__ __forceinline__ void flush_runs(float *__restrict__ g_tile, const float *s_warp, int off0, int off1, int cnt,
                                           int lane) {
    __builtin_assume(__isShared(s_warp));
    constexpr int W = NRUN * LEN;
    constexpr bool EVEN = (LEN % 2 == 0) && (PAD % 2 == 0) && (OUT % 2 == 0);
    if constexpr (EVEN && W / 2 <= 32) {
        if (((off0 | off1) & 1) == 0 && (reinterpret_cast<unsigned long long>(g_tile) & 7ull) == 0) {
            flush_rows<float2, OUT / 2, LEN / 2, PAD / 2, NRUN>(reinterpret_cast<float2 *>(g_tile),
                                                               reinterpret_cast<const float2 *>(s_warp), off0 / 2,
                                                               off1 / 2, cnt, lane);
            return;
        }
    }
    if constexpr (W <= 32) {
        flush_rows<float, OUT, LEN, PAD, NRUN>(g_tile, s_warp, off0, off1, cnt, lane);
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
    static __host__ __device__ constexpr int smem_words(int stage) { return (stage == 0 ? TILE_WORDS : 0) + STAGE_WORDS; }
};

// Work items are (task, block of WARPS consecutive tiles), numbered task-major (tasks sorted by
// decreasing cost); warp w of a CTA runs tile blk * WARPS + w.  Instruction supply shapes all of
// this (tools/micro/ifetch_bench.cu, profiles/r1_micro_ifetch.jsonl): straight-line code beyond the
// 32 KB SM instruction cache streams at ~0.3 instructions/cycle per warp; an SM whose warps sit at
// DIFFERENT places of such a program is capped near 1.0 IPC (each warp pulls its own stream from
// the GPC cache / L2), whereas warps that run the SAME lines at about the same time share every
// fetched line (2.7 IPC with 8 warps), and SMs that are in step share lines in the GPC cache
// (2.7 vs 1.6-2.1 IPC).  Hence task-major order (tile-major is 3.2x slower: eight ~100 KB programs
// thrash the caches), CTAs of WARPS warps that start a program together and re-align at a
// __syncthreads() every P::SYNC_EVERY operations, and persistent CTAs with a uniform stride.
template <class P, int STAGE>
__global__ void __launch_bounds__(32 * P::WARPS, STAGE == 0 ? P::MINB0 : P::MINB1)
pipe_kernel(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0, const float *__restrict__ d_in1,
            float *__restrict__ scratch, unsigned int *__restrict__ ticket, int num_states, int ntiles, int nblk,
            float gravity, float dt, const float *__restrict__ d_in2, int stagger_ns, int task0, int ntasks, int order_blk, int no_store) {
    using S = PipeShape<P>;
    // states per tile: a stage-1 warp of a P::X2 variant runs 64 states (two per lane, packed FP32 instructions);
    // `ntiles` counts tiles of that size.  Scratch rows then hold 64 states: [tile of 64][word][64].
    constexpr int SPT = (STAGE == 1 && P::X2) ? 64 : 32;
    constexpr int SC_ROW = P::X2 ? 64 : 32;
    // experiment (GRID_PIPE_STAGGER_NS): CTAs are placed round-robin over the SMs, so CTAs b, b + #SMs, ... share an
    // SM; starting them apart de-phases their instruction streams
    if (stagger_ns > 0) {
        unsigned nsm;
        asm("mov.u32 %0, %%nsmid;" : "=r"(nsm));
        for (unsigned w = (blockIdx.x / nsm) * (unsigned)stagger_ns; w > 0;) {
            const unsigned step = w < 500000u ? w : 500000u;
            __nanosleep(step);
            w -= step;
        }
    }
    extern __shared__ float smem_all[];
    __shared__ int s_item;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *smem = smem_all + warp * S::smem_words(STAGE);
    float *s_warp = smem + (STAGE == 0 ? S::TILE_WORDS : 0);
    // PERSISTENT CTAs walk the (task, block of WARPS tiles) items in task-major order, all with the
    // same stride: every SM runs the same program at the same time and with the same timing, so
    // the SMs of a GPC share the instruction stream in the GPC-level cache as well (dynamically
    // dispatched CTAs drift apart: the identical iiwa14 program runs 1.5-1.7x slower that way).
    // Items are handed out in task-major order: with a ticket counter when there are more items than
    // CTAs (tasks differ in length - Atlas stage 0: 27 k vs 12 k cost units - and a fixed stride leaves
    // some CTAs with two long items and others with one), with a fixed stride otherwise.
    for (int iter = 0;; iter++) {
        int item;
        if (ticket) {
            if (threadIdx.x == 0) s_item = (int)atomicAdd(ticket, 1u);
            __syncthreads();
            item = s_item;
            __syncthreads();
        } else {
            item = blockIdx.x + iter * gridDim.x;
        }
        if (item >= ntasks * nblk) break;
        int t_rel, blk;
        if (order_blk > 0) {
            // chunk-major order (stage 1): all tasks of a chunk of order_blk tile blocks before the next chunk, so that
            // a scratch line read by several column programs comes from DRAM once and from L2 afterwards; chunks in
            // DESCENDING order - the end of the batch is what stage 0 wrote last and is still in L2.  The (shorter)
            // last chunk goes first.
            const int nch = (nblk + order_blk - 1) / order_blk;
            const int nb_last = nblk - (nch - 1) * order_blk;
            int ch, nb, r = item;
            if (r < ntasks * nb_last) {
                ch = nch - 1;
                nb = nb_last;
            } else {
                r -= ntasks * nb_last;
                const int per = ntasks * order_blk;
                ch = nch - 2 - r / per;
                r -= (r / per) * per;
                nb = order_blk;
            }
            t_rel = r / nb;
            blk = ch * order_blk + (r - t_rel * nb);
        } else {
            t_rel = item / nblk;
            blk = item - t_rel * nblk;
        }
        const int task = task0 + t_rel;
        // The task programs contain CTA-wide barriers (they keep the warps of the CTA on the same
        // lines of the program), so a warp past the last tile cannot sit out: it recomputes the
        // last tile and stores nothing (its scratch writes duplicate the owner's values).
        const int my_tile = blk * (int)(blockDim.x >> 5) + warp;
        const bool owner = my_tile < ntiles && !no_store;      // no_store (profiling): every output store predicated off
        const int tile = my_tile < ntiles ? my_tile : ntiles - 1;
        const long long first = (long long)tile * SPT;
        const int cnt = min(SPT, num_states - (int)first);
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
        // third input (the caller's Minv of the USE_QDD_MINV_FLAG overload): read by the lane straight from global
        const float *g2 = P::IN2 > 0 ? d_in2 + (first + src) * (long long)P::IN2 : nullptr;
        // the lane's scratch column: state (first + lane) of a 32-state tile; garbage in the columns of states past
        // the end is computed on and never stored
        float *sc_row = scratch + (first / SC_ROW) * (long long)(P::SCRATCH_WORDS * SC_ROW) + (first % SC_ROW);
        if (SPT == 64) {
            if ((P::X2_MASK >> task) & 1ull) {
                // packed program: states first + 2 lane and first + 2 lane + 1
                P::template run<STAGE>(task, smem, sc_row + 2 * lane, sc_row, s_warp + lane * P::STAGE_PAD,
                                       d_out + first * P::OUT, owner ? cnt : 0, lane, s_warp, gravity, dt, g2);
            } else {
                // scalar program on a 64-state tile: two halves (every warp of the CTA runs the same task, so the
                // barriers inside the programs stay matched)
#pragma unroll 1
                for (int h = 0; h < 2; h++) {
                    const int cnt_h = max(0, min(32, cnt - 32 * h));
                    P::template run<STAGE>(task, smem, sc_row + 32 * h + lane, sc_row, s_warp + lane * P::STAGE_PAD,
                                           d_out + (first + 32 * h) * P::OUT, owner ? cnt_h : 0, lane, s_warp, gravity,
                                           dt, g2);
                    __syncwarp();
                }
            }
        } else {
            float *sc = sc_row + lane;
            P::template run<STAGE>(task, smem + src * S::IN_PAD, sc, sc, s_warp + lane * P::STAGE_PAD,
                                   d_out + first * P::OUT, owner ? cnt : 0, lane, s_warp, gravity, dt, g2);
        }
        __syncwarp();
    }
}

// Warps per CTA are a launch-time choice (the kernel is compiled for up to P::WARPS): 8 when there is
// enough work, fewer for mid-size batches so that the (task, tiles) items still cover every SM
// (Atlas FD gradient at 8 192 states: 3 x 32 items of 8 tiles would occupy 96 of 148 SMs).
template <class P, int STAGE>
cudaError_t pipe_stage_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, float *scratch,
                              unsigned int *ticket, int num_states, float gravity, cudaStream_t stream, float dt,
                              const float *d_in2) {
    using S = PipeShape<P>;
    constexpr int ntasks_all = STAGE == 0 ? P::NTASKS0 : P::NTASKS1;
    if (ntasks_all == 0) return cudaSuccess;
    // profiling (GRID_PIPE_ONLY_TASK = 100 * stage + task): only that task program runs, the other stage is skipped
    int task0 = 0, ntasks = ntasks_all;
    // GRID_PIPE_ONLY_TASK = 1000: all tasks, but no output stores (what the writes cost; results are not produced)
    const int no_store = options().pipe_only_task == 1000;
    if (const int only = options().pipe_only_task; only >= 0 && only < 1000) {
        if (only / 100 != STAGE || only % 100 >= ntasks_all) return cudaSuccess;
        task0 = only % 100;
        ntasks = 1;
    }
    auto kern = pipe_kernel<P, STAGE>;
    constexpr size_t warp_smem = sizeof(float) * S::smem_words(STAGE);
    // candidate CTA sizes: 1, 2, 4, 8 warps and the size the kernel was compiled for
    constexpr int NOPT = 5;
    static const int opt_w[NOPT] = {1, 2, 4, 8, P::WARPS};
    struct DevCfg { int sms, per_sm[NOPT]; };               // resident CTAs per SM per option
    static DevCfg cfgs[kMaxDevices];                        // per device (benign race: idempotent)
    int dev = 0;
    if (cudaError_t e = current_device(dev)) return e;
    int &sms = cfgs[dev].sms;
    int *per_sm = cfgs[dev].per_sm;
    if (sms == 0) {
        int n = 0, smem_max = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        size_t want = warp_smem * P::WARPS;
        if (want > (size_t)smem_max) want = (size_t)smem_max / warp_smem * warp_smem;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)want);
        if (e != cudaSuccess) return e;
        for (int i = 0; i < NOPT; i++) {
            const int w = opt_w[i];
            if (w > P::WARPS || (i < NOPT - 1 && w == P::WARPS)) continue;      // too big / listed last
            if (warp_smem * w > want) continue;                                    // staging does not fit
            e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[i], kern, 32 * w, warp_smem * w);
            if (e != cudaSuccess) return e;
        }
        if (per_sm[NOPT - 1] < 1 && per_sm[0] < 1) return cudaErrorLaunchOutOfResources;
        sms = n;
    }
    constexpr int spt = (STAGE == 1 && P::X2) ? 64 : 32;
    const int ntiles = (num_states + spt - 1) / spt;
    // largest CTA whose items still give every resident CTA slot at least two items, else the smallest
    int wi = -1;
    for (int i = NOPT - 1; i >= 0; i--) {
        if (per_sm[i] < 1) continue;
        wi = i;
        const long long items_w = (long long)ntasks * ((ntiles + opt_w[i] - 1) / opt_w[i]);
        if (items_w >= 2LL * sms * per_sm[i]) break;
    }
    if (const int forced = options().pipe_warps) {        // experiments: GRID_PIPE_WARPS pins the CTA width
        for (int i = 0; i < NOPT; i++)
            if (opt_w[i] == forced && per_sm[i] >= 1) wi = i;
    }
    if (wi < 0) return cudaErrorLaunchOutOfResources;
    const int w = opt_w[wi];
    const int nblk = (ntiles + w - 1) / w;
    const long long items = (long long)ntasks * nblk;
    const long long cap = (long long)sms * per_sm[wi];
    const int blocks = (int)(items < cap ? items : cap);
    // stage 1 of a two-stage variant: chunk-major item order when the batch is several chunks long
    int order_blk = 0;
    if (STAGE == 1 && P::SCRATCH_WORDS > 0 && ntasks > 1) {
        const int oc = options().pipe_order_chunk >= 0 ? options().pipe_order_chunk : P::ORDER_CHUNK_STATES;
        if (oc > 0 && num_states >= 2 * oc) order_blk = (oc / spt + w - 1) / w;
    }
    kern<<<blocks, 32 * w, warp_smem * w, stream>>>(d_out, d_in0, stride0, d_in1, scratch,
                                                    items > cap ? ticket : nullptr, num_states, ntiles, nblk, gravity,
                                                    dt, d_in2, options().pipe_stagger_ns, task0, ntasks, order_blk, no_store);
    g_kernel_launches.fetch_add(1);
    return cudaGetLastError();
}

// ---- fused kernel (experimental, GRID_PIPE_MODE=fused): every SM runs ONE task for the whole launch --
// The staged kernels walk the tasks one after the other, so an SM changes program every 1-2 items
// and the stage-1 kernel re-reads the scratch array from HBM once per task.  Here the CTAs (one per
// SM) are PARTITIONED among the tasks of both stages in proportion to the tasks' instruction counts:
// CTA `rank` of the m CTAs of a task runs tiles rank, rank + m, ... in ascending order.  Each SM
// then executes one <= ~100 KB program over and over (its first 64 KB stay in the instruction
// cache, profiles/r1_ifetch_regions.md), all task groups advance through the tiles at the same
// rate, so a tile's scratch lines are consumed from L2 right after stage 0 produced them and the
// output columns of a tile complete their sectors in L2.  Stage-1 warps wait for the stage-0 task
// of their component on a per-(task, tile) flag (release/acquire at gpu scope).  Dependencies only
// point to lower CTA indices, which the hardware schedules first, so the kernel cannot deadlock
// even when not all CTAs are resident at once.
struct PipePart {
    short first[64], count[64];
};

template <class P>
__global__ void __launch_bounds__(32 * P::WARPS, 1)
pipe_fused_kernel(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0,
                  const float *__restrict__ d_in1, float *__restrict__ scratch, int *__restrict__ flags, int num_states,
                  int ntiles, int nblk, float gravity, float dt, const PipePart part, const float *__restrict__ d_in2) {
    using S = PipeShape<P>;
    extern __shared__ float smem_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *smem = smem_all + warp * S::smem_words(0);
    float *s_warp = smem + S::TILE_WORDS;
    int k = 0;
    while (k < P::NT - 1 && (int)blockIdx.x >= part.first[k] + part.count[k]) k++;
    const int rank = blockIdx.x - part.first[k], m = part.count[k];
    const bool stage0 = k < P::NTASKS0;
    const int dep = P::dep(k);
    for (int blk = rank; blk < nblk; blk += m) {
        const int my_tile = blk * P::WARPS + warp;
        const bool owner = my_tile < ntiles;
        const int tile = owner ? my_tile : ntiles - 1;
        const long long first = (long long)tile * 32;
        const int cnt = min(32, num_states - (int)first);
        if (stage0) {
            const float *src0 = d_in0 + first * (long long)stride0;
            if (S::IN_LINEAR && stride0 == P::IN0 && aligned16(src0)) {
                warp_copy_g2s(smem, src0, cnt * P::IN0, lane);
            } else {
                tile_load<P::IN0, S::IN_PAD>(smem, 0, d_in0, first, stride0, cnt, lane);
                tile_load<P::IN1, S::IN_PAD>(smem, P::IN0, d_in1, first, P::IN1, cnt, lane);
            }
        } else if (dep >= 0) {
            if (lane == 0) {
                const int *f = flags + (long long)dep * ntiles + tile;
                while (ld_acquire(f) == 0) __nanosleep(200);
            }
        }
        __syncwarp();
        const int src = min(lane, cnt - 1);
        float *sc = scratch + (long long)tile * (P::SCRATCH_WORDS * 32) + lane;
        const float *g2 = P::IN2 > 0 ? d_in2 + (first + src) * (long long)P::IN2 : nullptr;
        if (stage0)
            P::template run<0>(k, smem + src * S::IN_PAD, sc, sc, s_warp + lane * P::STAGE_PAD, d_out + first * P::OUT,
                               owner ? cnt : 0, lane, s_warp, gravity, dt, g2);
        else
            P::template run<1>(k - P::NTASKS0, smem, sc, sc, s_warp + lane * P::STAGE_PAD, d_out + first * P::OUT,
                               owner ? cnt : 0, lane, s_warp, gravity, dt, g2);
        if (stage0 && P::SCRATCH_WORDS > 0) {
            __threadfence();
            __syncwarp();
            if (lane == 0 && owner) st_release(flags + (long long)k * ntiles + tile, 1);
        }
        __syncwarp();
    }
}

// CTAs per task: start with one each, then hand the remaining CTAs one at a time to the task with
// the largest remaining time ceil(nblk / m) * cost (never more CTAs than items).
template <class P>
static int pipe_partition(int G, int nblk, PipePart &part) {
    int m[64];
    for (int k = 0; k < P::NT; k++) m[k] = 1;
    for (int used = P::NT; used < G; used++) {
        int best = -1;
        long long worst = -1;
        for (int k = 0; k < P::NT; k++) {
            if (m[k] >= nblk) continue;
            const long long load = (long long)((nblk + m[k] - 1) / m[k]) * P::cost(k);
            if (load > worst) { worst = load; best = k; }
        }
        if (best < 0) break;
        m[best]++;
    }
    int at = 0;
    for (int k = 0; k < P::NT; k++) {
        part.first[k] = (short)at;
        part.count[k] = (short)m[k];
        at += m[k];
    }
    return at;
}

template <class P>
cudaError_t pipe_fused_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, int num_states,
                              float gravity, cudaStream_t stream, bool &handled, float dt, const float *d_in2) {
    using S = PipeShape<P>;
    handled = false;
    auto kern = pipe_fused_kernel<P>;
    constexpr size_t smem_bytes = sizeof(float) * S::smem_words(0) * P::WARPS;
    static int caps[kMaxDevices];
    int dev = 0;
    if (cudaError_t e = current_device(dev)) return e;
    int &cap = caps[dev];
    if (cap == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * P::WARPS, smem_bytes);
        if (e != cudaSuccess) return e;
        cap = per_sm < 1 ? -1 : sms;             // one CTA per SM: an SM never mixes programs
    }
    if (cap < P::NT || P::NT > 64) return cudaSuccess;           // not handled: the staged kernels take over
    const int ntiles = (num_states + 31) / 32;
    const int nblk = (ntiles + P::WARPS - 1) / P::WARPS;
    PipePart part;
    const int blocks = pipe_partition<P>(cap, nblk, part);
    const size_t sc_bytes = (size_t)ntiles * P::SCRATCH_WORDS * 32 * sizeof(float);
    const size_t fl_bytes = P::SCRATCH_WORDS > 0 ? (size_t)ntiles * P::NTASKS0 * sizeof(int) : 0;
    char *buf = nullptr;
    cudaError_t e = cudaSuccess;
    if (sc_bytes + fl_bytes > 0) {
        keep_pool_memory();
        e = cudaMallocAsync((void **)&buf, sc_bytes + fl_bytes, stream);
        if (e != cudaSuccess) return e;
        if (fl_bytes) e = cudaMemsetAsync(buf + sc_bytes, 0, fl_bytes, stream);
    }
    if (e == cudaSuccess) {
        kern<<<blocks, 32 * P::WARPS, smem_bytes, stream>>>(d_out, d_in0, stride0, d_in1, (float *)buf,
                                                            (int *)(buf + sc_bytes), num_states, ntiles, nblk, gravity,
                                                            dt, part, d_in2);
        g_kernel_launches.fetch_add(1);
        e = cudaGetLastError();
    }
    if (buf) {
        cudaError_t e2 = cudaFreeAsync(buf, stream);
        if (e == cudaSuccess) e = e2;
    }
    handled = true;
    return e;
}

// Entry point: both stages on `stream` (or the experimental fused kernel, see below).
template <class P>
cudaError_t pipe_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, int num_states,
                        float gravity, cudaStream_t stream, float dt = 0.f, const float *d_in2 = nullptr) {
    if (num_states <= 0) return cudaSuccess;
    g_calls.fetch_add(1);
    float *scratch = nullptr;
    cudaError_t e;
    // GRID_PIPE_MODE=fused selects the SM-partitioned single-kernel variant.  It is correct (tests run
    // it) but measured SLOWER than the staged kernels on every robot (Atlas FD gradient 975 vs 760 us,
    // HyQ 56 vs 43 us, profiles/r1_pipe_fused_vs_staged.md): SMs of one GPC running different programs
    // lose the sharing of the instruction stream in the GPC-level cache.
    // (not instantiated for P::X2 variants: their stage-1 programs take 64-state tiles)
    if constexpr (!P::X2) {
        const bool fused = options().pipe_fused != 0;
        if (fused && P::NT > 1) {
            bool handled = false;
            e = pipe_fused_launch<P>(d_out, d_in0, stride0, d_in1, num_states, gravity, stream, handled, dt, d_in2);
            if (handled || e != cudaSuccess) return e;
        }
    }
    // Two-stage variants can run in chunks of P::CHUNK_STATES states (a multiple of 32; GRID_PIPE_CHUNK
    // overrides; 0 = one chunk, the default): the scratch words of a chunk would then still be in L2
    // when stage 1 reads them.  Measured (profiles/r1_pipe_order.md): chunking LOSES (Atlas FD gradient
    // 769 us unchunked, 789 / 979 / 1083 us with 32 k / 16 k / 8 k chunks) - fewer tiles per task per SM
    // means less re-execution of a program while it is in the instruction caches.
    int chunk = num_states;
    const int chunk_cfg = options().pipe_chunk >= 0 ? options().pipe_chunk : P::CHUNK_STATES;
    if (P::SCRATCH_WORDS > 0 && chunk_cfg > 0 && chunk_cfg < num_states) chunk = (chunk_cfg + 63) / 64 * 64;
    const int nchunks = (num_states + chunk - 1) / chunk;
    // one allocation: [ticket counters: 2 per chunk, padded to 256 B | scratch words of one chunk];
    // small batches need no tickets, single-stage variants no scratch
    static int sms_of[kMaxDevices];
    int dev = 0;
    if (cudaError_t e0 = current_device(dev)) return e0;
    int &sms = sms_of[dev];
    if (sms == 0) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    constexpr int max_tasks = P::NTASKS0 > P::NTASKS1 ? P::NTASKS0 : P::NTASKS1;
    const long long max_items = (long long)max_tasks * (((chunk + 31) / 32 + P::WARPS - 1) / P::WARPS);
    // (the counters cost a memset node: ~1.5 us, visible on 10 us launches such as HyQ Minv at 16 384
    // states - only worth it when every CTA takes several items; the stage launcher ignores them when
    // every item has its own CTA)
    const bool use_tickets = max_items > 4LL * sms;
    const size_t tk_bytes = use_tickets ? ((size_t)nchunks * 2 * sizeof(unsigned int) + 255) / 256 * 256 : 0;
    constexpr int sc_row = P::X2 ? 64 : 32;
    const size_t sc_bytes = (size_t)((chunk + sc_row - 1) / sc_row) * P::SCRATCH_WORDS * sc_row * sizeof(float);
    unsigned int *tickets = nullptr;
    float *sc = nullptr;
    e = cudaSuccess;
    if (tk_bytes + sc_bytes > 0) {
        keep_pool_memory();
        e = cudaMallocAsync((void **)&scratch, tk_bytes + sc_bytes, stream);
        if (e != cudaSuccess) return e;
        sc = reinterpret_cast<float *>(reinterpret_cast<char *>(scratch) + tk_bytes);
        if (use_tickets) {
            tickets = reinterpret_cast<unsigned int *>(scratch);
            e = cudaMemsetAsync(tickets, 0, tk_bytes, stream);
        }
    }
    for (int c = 0; c < nchunks && e == cudaSuccess; c++) {
        const int first = c * chunk;
        const int n = num_states - first < chunk ? num_states - first : chunk;
        float *o = d_out + (long long)first * P::OUT;
        const float *i0 = d_in0 + (long long)first * stride0;
        const float *i1 = d_in1 ? d_in1 + (long long)first * P::IN1 : nullptr;
        const float *i2 = d_in2 ? d_in2 + (long long)first * P::IN2 : nullptr;
        e = pipe_stage_launch<P, 0>(o, i0, stride0, i1, sc, tickets ? tickets + 2 * c : nullptr, n, gravity, stream, dt, i2);
        if (e == cudaSuccess)
            e = pipe_stage_launch<P, 1>(o, i0, stride0, i1, sc, tickets ? tickets + 2 * c + 1 : nullptr, n, gravity, stream,
                                        dt, i2);
    }
    if (scratch) {
        cudaError_t e2 = cudaFreeAsync(scratch, stream);
        if (e == cudaSuccess) e = e2;
    }
    return e;
}

}}  // namespace GRID_NS::pipe
