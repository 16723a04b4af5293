// grid_tps.cuh - thread-per-state launch shell around the traced straight-line programs.
//
// Replaces the reference's block-per-state kernels (e.g. algorithms/
// _forward_dynamics_gradient.py:107-177: one CTA per knot point, ~111 __syncthreads per
// state) for robots whose traced program fits one thread: every lane owns one state and
// runs the robot-specialised straight-line code, so there are no barriers, shuffles,
// atomics or topology lookups on the critical path.  A warp stages its 32 states through
// shared memory so that global loads and stores are coalesced although the caller's layout
// is state-major.
#pragma once
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>

namespace GRID_NS {

// ---- process-wide options --------------------------------------------------------------------
// Read from the environment ONCE (first use), changed afterwards only through grid_set_option()
// (include/grid_b200.h): the launch path never calls getenv.
enum ForceKernel { kAuto = 0, kTps = 1, kWps = 2, kCps = 3, kPipe = 4, kLps = 5 };
struct Options {
    int force_kernel;       // ForceKernel
    int pipe_fused;         // 1 = experimental single-launch variant of the phase-split kernels
    int pipe_chunk;         // states per chunk of a two-stage launch; -1 = the compiled default
    int pipe_warps;         // warps per CTA of the phase-split kernels; 0 = chosen from the batch size
    int pipe_stagger_ns;    // experiment: CTAs that share an SM start this many ns apart (de-phased instruction streams)
    int pipe_order_chunk;   // stage-1 items ordered chunk-major: states per chunk (0 = task-major over the whole batch; -1 = default)
    int pipe_only_task;     // profiling: 100 * stage + task = launch only that task program (results incomplete); -1 = all
};
inline int parse_force(const char *f) {
    if (!f || !*f) return kAuto;
    if (!strcmp(f, "tps")) return kTps;
    if (!strcmp(f, "wps")) return kWps;
    if (!strcmp(f, "cps")) return kCps;
    if (!strcmp(f, "pipe")) return kPipe;
    if (!strcmp(f, "lps")) return kLps;
    return -1;
}
inline Options &options() {
    static Options o = [] {
        Options x;
        const int f = parse_force(getenv("GRID_FORCE_KERNEL"));
        x.force_kernel = f < 0 ? kAuto : f;
        const char *m = getenv("GRID_PIPE_MODE");
        x.pipe_fused = (m && !strcmp(m, "fused")) ? 1 : 0;
        const char *c = getenv("GRID_PIPE_CHUNK");
        x.pipe_chunk = c ? atoi(c) / 32 * 32 : -1;
        const char *w = getenv("GRID_PIPE_WARPS");
        x.pipe_warps = w ? atoi(w) : 0;
        const char *st = getenv("GRID_PIPE_STAGGER_NS");
        x.pipe_stagger_ns = st ? atoi(st) : 0;
        const char *oc = getenv("GRID_PIPE_ORDER_CHUNK");
        x.pipe_order_chunk = oc ? atoi(oc) : -1;
        const char *ot = getenv("GRID_PIPE_ONLY_TASK");
        x.pipe_only_task = ot ? atoi(ot) : -1;
        return x;
    }();
    return o;
}

// ---- per-device launch state -----------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize), the SM count and the occupancy caps belong to a
// device, and one process may drive several (a torch user calling cuda:0 then cuda:1): every
// launcher keeps its cached state in an array indexed by the CURRENT device.
constexpr int kMaxDevices = 64;
inline cudaError_t current_device(int &dev) {
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    return (dev < 0 || dev >= kMaxDevices) ? cudaErrorInvalidDevice : cudaSuccess;
}

// The scratch arrays of the phase-split and chain kernels come from the stream-ordered allocator (no library
// state, safe with concurrent callers on different streams); keeping the pool's memory between calls makes the
// allocation a free-list hit after the first launch.
static void keep_pool_memory() {
    static bool dones[kMaxDevices];
    int dev = 0;
    if (current_device(dev) != cudaSuccess) return;
    bool &done = dones[dev];
    if (done) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done = true;
}

constexpr int odd_pad(int words) { return words | 1; }   // odd stride => conflict-free lane access
constexpr int cmax(int a, int b) { return a > b ? a : b; }

constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }

// Row strides of a warp's staging tile.  Lane-private accesses (lane L touches row L) hit
// gcd(stride, 32) lanes per bank.  When the exact row length already gives <= 2-way conflicts the
// tile is stored UNPADDED: it is then bit-identical to the caller's state-major layout, so the
// global <-> shared copies are plain 128-bit linear copies instead of per-element index arithmetic
// (9 % of the iiwa14 FD-gradient instruction stream).  Otherwise rows are padded to an odd length.
template <class A>
struct TpsShape {
    static constexpr int IN = A::IN0 + A::IN1 + A::IN2;
    static constexpr bool IN_LINEAR = (A::IN1 == 0) && (A::IN2 == 0) && cgcd(IN, 32) <= 2;
    static constexpr bool OUT_LINEAR = cgcd(A::OUT, 32) <= 2;
    static constexpr int IN_PAD = IN_LINEAR ? IN : odd_pad(IN);
    static constexpr int OUT_PAD = OUT_LINEAR ? A::OUT : odd_pad(A::OUT);
    static constexpr int WARP_WORDS = (32 * cmax(IN_PAD, OUT_PAD) + 3) / 4 * 4;   // tiles alias; 16 B aligned
};

__device__ __forceinline__ bool aligned16(const void *p) { return (reinterpret_cast<unsigned long long>(p) & 15ull) == 0; }

// general path: any stride, padded rows, per-element index arithmetic
template <int WORDS, int PAD>
__device__ __forceinline__ void tile_load(float *sw, int word_off, const float *__restrict__ g,
                                          long long first_state, int stride, int cnt, int lane) {
    if (WORDS == 0) return;
    const float *src = g + first_state * (long long)stride;
    for (int e = lane; e < cnt * WORDS; e += 32) {
        int s = e / WORDS, k = e - s * WORDS;
        sw[s * PAD + word_off + k] = __ldg(src + (long long)s * stride + k);
    }
}

// words floats, both pointers 16-byte aligned: 128-bit body + scalar tail
__device__ __forceinline__ void warp_copy_g2s(float *dst, const float *__restrict__ src, int words, int lane) {
    const int n4 = words >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int e = lane; e < n4; e += 32) d4[e] = __ldg(s4 + e);
    for (int e = (n4 << 2) + lane; e < words; e += 32) dst[e] = __ldg(src + e);
}
__device__ __forceinline__ void warp_copy_s2g(float *__restrict__ dst, const float *src, int words, int lane) {
    const int n4 = words >> 2;
    const float4 *s4 = reinterpret_cast<const float4 *>(src);
    float4 *d4 = reinterpret_cast<float4 *>(dst);
    for (int e = lane; e < n4; e += 32) d4[e] = s4[e];
    for (int e = (n4 << 2) + lane; e < words; e += 32) dst[e] = src[e];
}

// One warp = one tile of 32 consecutive states.  Works for any 1-D/2-D launch shape: warps
// beyond the dynamic shared memory that was provided (32*max(IN_PAD,OUT_PAD) floats per warp)
// and partial warps take no tiles.  All warps of a CTA run the loop the same number of times
// (a warp without a tile computes on stale data and stores nothing), because the traced
// program may contain CTA-wide barriers.
template <class A>
__device__ __forceinline__ void tps_body(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0,
                                         const float *__restrict__ d_in1, const float *__restrict__ d_in2,
                                         int num_states, float gravity, float dt) {
    using S = TpsShape<A>;
    extern __shared__ float smem[];
    unsigned smem_bytes;
    asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(smem_bytes));
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int full_warps = (blockDim.x * blockDim.y * blockDim.z) >> 5;
    const int warps = min(full_warps, (int)(smem_bytes / (sizeof(float) * S::WARP_WORDS)));
    const bool worker = warp < warps;
    float *sw = smem + (worker ? warp : 0) * S::WARP_WORDS;
    const int ntiles = (num_states + 31) >> 5;
    const int nblocks = gridDim.x * gridDim.y, block = blockIdx.x + blockIdx.y * gridDim.x;
    if (warps == 0) return;
    for (int tile0 = block * warps; tile0 < ntiles; tile0 += nblocks * warps) {
        const int tile = tile0 + warp;
        const long long first = (long long)tile * 32;
        const int cnt = worker ? max(0, min(32, num_states - (int)first)) : 0;
        const float *src0 = d_in0 + first * (long long)stride0;
        if (S::IN_LINEAR && stride0 == A::IN0 && aligned16(src0)) {
            warp_copy_g2s(sw, src0, cnt * A::IN0, lane);
        } else {
            tile_load<A::IN0, S::IN_PAD>(sw, 0, d_in0, first, stride0, cnt, lane);
            tile_load<A::IN1, S::IN_PAD>(sw, A::IN0, d_in1, first, A::IN1, cnt, lane);
            tile_load<A::IN2, S::IN_PAD>(sw, A::IN0 + A::IN1, d_in2, first, A::IN2, cnt, lane);
        }
        __syncwarp();
        // lanes past the end of a ragged tile recompute the last valid state (results dropped)
        const int src = max(0, min(lane, cnt - 1));
        if (worker) A::eval(sw + src * S::IN_PAD, sw + lane * S::OUT_PAD, gravity, dt);
        __syncwarp();
        float *dst = d_out + first * A::OUT;
        if (S::OUT_LINEAR && aligned16(dst)) {
            warp_copy_s2g(dst, sw, cnt * A::OUT, lane);
        } else {
            for (int e = lane; e < cnt * A::OUT; e += 32) {
                int s = e / A::OUT, k = e - s * A::OUT;
                dst[e] = sw[s * S::OUT_PAD + k];
            }
        }
        __syncwarp();
    }
}

template <class A, int WARPS, int MIN_BLOCKS>
__global__ void __launch_bounds__(32 * WARPS, MIN_BLOCKS)
tps_kernel(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0,
           const float *__restrict__ d_in1, const float *__restrict__ d_in2, int num_states, float gravity, float dt) {
    tps_body<A>(d_out, d_in0, stride0, d_in1, d_in2, num_states, gravity, dt);
}

// ---- mid-size batches: (tile, half) per warp ---------------------------------------------------------------
// A0 / A1 compute the d/dq / d/dqd block of the FD gradient (OUT = n*n words each, the column-independent part
// duplicated).  Work item = (tile, half); the output row of a state is 2 * OUT words, a half owns OUT of them.
template <class A0, class A1, int MIN_BLOCKS>
__global__ void __launch_bounds__(32, MIN_BLOCKS)
tps_half_kernel(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0, int num_states, float gravity) {
    using S = TpsShape<A0>;
    static_assert(A0::IN0 == A1::IN0 && A0::OUT == A1::OUT && A0::IN1 == 0 && A1::IN1 == 0, "halves must match");
    extern __shared__ float smem[];
    float *sw = smem;
    const int lane = threadIdx.x & 31;
    const int ntiles = (num_states + 31) >> 5;
    for (int item = blockIdx.x; item < 2 * ntiles; item += gridDim.x) {
        const int tile = item >> 1, half = item & 1;
        const long long first = (long long)tile * 32;
        const int cnt = max(0, min(32, num_states - (int)first));
        const float *src0 = d_in0 + first * (long long)stride0;
        if (S::IN_LINEAR && stride0 == A0::IN0 && aligned16(src0)) warp_copy_g2s(sw, src0, cnt * A0::IN0, lane);
        else tile_load<A0::IN0, S::IN_PAD>(sw, 0, d_in0, first, stride0, cnt, lane);
        __syncwarp();
        const int src = max(0, min(lane, cnt - 1));
        if (half == 0) A0::eval(sw + src * S::IN_PAD, sw + lane * S::OUT_PAD, gravity, 0.f);
        else A1::eval(sw + src * S::IN_PAD, sw + lane * S::OUT_PAD, gravity, 0.f);
        __syncwarp();
        float *dst = d_out + first * (2 * A0::OUT) + half * A0::OUT;
        for (int e = lane; e < cnt * A0::OUT; e += 32) {
            const int st = e / A0::OUT, k = e - st * A0::OUT;
            dst[(long long)st * (2 * A0::OUT) + k] = sw[st * S::OUT_PAD + k];
        }
        __syncwarp();
    }
}

// resident single-warp CTAs of the half kernel on the current device (0 on error)
template <class A0, class A1, int MIN_BLOCKS>
static int tps_half_cap() {
    using S = TpsShape<A0>;
    auto kern = tps_half_kernel<A0, A1, MIN_BLOCKS>;
    constexpr size_t smem_bytes = sizeof(float) * S::WARP_WORDS;
    static int caps[kMaxDevices];
    int dev = 0;
    if (current_device(dev) != cudaSuccess) return 0;
    if (caps[dev] == 0) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes) != cudaSuccess) return 0;
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem_bytes);
        caps[dev] = sms * (per_sm > 0 ? per_sm : 1);
    }
    return caps[dev];
}
// the split pays while every (tile, half) item has its own resident warp
template <class A0, class A1, int MIN_BLOCKS>
static bool tps_half_fits(int num_states) {
    const int cap = tps_half_cap<A0, A1, MIN_BLOCKS>();
    return cap > 0 && 2 * ((num_states + 31) / 32) <= cap;
}
template <class A0, class A1, int MIN_BLOCKS>
cudaError_t tps_half_launch(float *d_out, const float *d_in0, int stride0, int num_states, float gravity, cudaStream_t stream) {
    using S = TpsShape<A0>;
    if (num_states <= 0) return cudaSuccess;
    const int cap = tps_half_cap<A0, A1, MIN_BLOCKS>();
    if (cap <= 0) return cudaErrorLaunchOutOfResources;
    const int items = 2 * ((num_states + 31) / 32);
    tps_half_kernel<A0, A1, MIN_BLOCKS><<<items < cap ? items : cap, 32, sizeof(float) * S::WARP_WORDS, stream>>>(
        d_out, d_in0, stride0, num_states, gravity);
    return cudaGetLastError();
}

// ---- variant 2: per-column output flush + parked per-state data ------------------------------
// The gradient programs can park long-lived per-joint values in a lane-private shared-memory
// column (slot-major, conflict-free) and re-load them in every du-column, and they flush each
// finished column pair to global memory instead of staging the whole 2n*n output tile.  Both cut
// shared memory and registers per warp so that more warps are resident per scheduler.
template <int NJ>
__device__ __forceinline__ void flush_colpair(float *__restrict__ g_tile, const float *s_warp, int j, int cnt, int lane) {
    constexpr int W = 2 * NJ, PAD = W | 1, OUT = 2 * NJ * NJ;
    for (int e = lane; e < cnt * W; e += 32) {
        const int s = e / W, c = e - s * W;
        g_tile[s * OUT + (c < NJ ? NJ * j + c : NJ * NJ + NJ * j + c - NJ)] = s_warp[s * PAD + c];
    }
}

template <class A>
struct Tps2Shape {
    static constexpr int IN = A::IN0 + A::IN1 + A::IN2;
    static constexpr int IN_PAD = odd_pad(IN);
    static constexpr int COL_PAD = odd_pad(2 * A::NJ);
    static constexpr int STAGE_WORDS = 32 * cmax(IN_PAD, COL_PAD);
    static constexpr int WARP_WORDS = STAGE_WORDS + 32 * A::PARK_SLOTS;
};

template <class A, int WARPS, int MIN_BLOCKS>
__global__ void __launch_bounds__(32 * WARPS, MIN_BLOCKS)
tps2_kernel(float *__restrict__ d_out, const float *__restrict__ d_in0, int stride0,
            const float *__restrict__ d_in1, const float *__restrict__ d_in2, int num_states, float gravity) {
    using S = Tps2Shape<A>;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sw = smem + warp * S::WARP_WORDS;
    float *sp = sw + S::STAGE_WORDS + lane;
    const int ntiles = (num_states + 31) >> 5;
    for (int tile = blockIdx.x * WARPS + warp; tile < ntiles; tile += gridDim.x * WARPS) {
        const long long first = (long long)tile * 32;
        const int cnt = min(32, num_states - (int)first);
        tile_load<A::IN0, S::IN_PAD>(sw, 0, d_in0, first, stride0, cnt, lane);
        tile_load<A::IN1, S::IN_PAD>(sw, A::IN0, d_in1, first, A::IN1, cnt, lane);
        tile_load<A::IN2, S::IN_PAD>(sw, A::IN0 + A::IN1, d_in2, first, A::IN2, cnt, lane);
        __syncwarp();
        const int src = min(lane, cnt - 1);
        A::eval(sw + src * S::IN_PAD, sw + lane * S::COL_PAD, sp, d_out + first * A::OUT, cnt, lane, sw, gravity);
        __syncwarp();
    }
}

template <class A, int WARPS, int MIN_BLOCKS>
cudaError_t tps2_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, const float *d_in2,
                        int num_states, float gravity, cudaStream_t stream) {
    using S = Tps2Shape<A>;
    if (num_states <= 0) return cudaSuccess;
    auto kern = tps2_kernel<A, WARPS, MIN_BLOCKS>;
    constexpr size_t smem_bytes = sizeof(float) * S::WARP_WORDS * WARPS;
    static int caps[kMaxDevices];               // resident CTAs per device (benign race: idempotent)
    int dev = 0;
    if (cudaError_t e = current_device(dev)) return e;
    int &cap = caps[dev];
    if (cap == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * WARPS, smem_bytes);
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        cap = sms * per_sm;
    }
    const int ntiles = (num_states + 31) / 32;
    int blocks = (ntiles + WARPS - 1) / WARPS;
    if (blocks > cap) blocks = cap;
    kern<<<blocks, 32 * WARPS, smem_bytes, stream>>>(d_out, d_in0, stride0, d_in1, d_in2, num_states, gravity);
    return cudaGetLastError();
}

template <class A, int WARPS, int MIN_BLOCKS>
cudaError_t tps_launch(float *d_out, const float *d_in0, int stride0, const float *d_in1, const float *d_in2,
                       int num_states, float gravity, cudaStream_t stream, float dt = 0.f) {
    using S = TpsShape<A>;
    if (num_states <= 0) return cudaSuccess;
    auto kern = tps_kernel<A, WARPS, MIN_BLOCKS>;
    constexpr size_t smem_bytes = sizeof(float) * S::WARP_WORDS * WARPS;
    static int caps[kMaxDevices];               // resident CTAs per device; beyond it, CTAs loop (benign race)
    int dev = 0;
    if (cudaError_t e = current_device(dev)) return e;
    int &cap = caps[dev];
    if (cap == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * WARPS, smem_bytes);
        cap = sms * (per_sm > 0 ? per_sm : 1);
    }
    const int ntiles = (num_states + 31) / 32;
    int blocks = (ntiles + WARPS - 1) / WARPS;
    if (blocks > cap) blocks = cap;
    kern<<<blocks, 32 * WARPS, smem_bytes, stream>>>(d_out, d_in0, stride0, d_in1, d_in2, num_states, gravity, dt);
    return cudaGetLastError();
}

}  // namespace GRID_NS
