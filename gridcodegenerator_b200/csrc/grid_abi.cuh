// grid_abi.cuh - the C ABI of include/grid_b200.h on top of the generated launchers.
//
// The generated translation unit defines, before including this file:
//   GRID_ROBOT_NAME, GRID_ROBOT_HASH, GRID_N
//   GRID_NS::gen::launch_id / launch_minv / launch_fd / launch_id_grad / launch_fd_grad
//   GRID_NS::gen::kernel_kind(alg), GRID_NS::gen::traced_flops(alg)
// Host-side structure mirrors the reference's emitted host layer
// (GRiDCodeGenerator.py:87-203 gridData/init_gridData/init_grid/close_grid and the
// *_host generators, e.g. algorithms/_forward_dynamics_gradient.py:179-242) with three
// changes: pinned host buffers, per-chunk stream pipelining of H2D / kernel / D2H instead
// of three cudaDeviceSynchronize calls, and status codes instead of exit().
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>

#include "grid_b200.h"

namespace GRID_NS {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

static int fail(const char *where, cudaError_t e) {
    g_last_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return (int)e == 0 ? -1 : (int)e;
}
static int fail_msg(const char *msg) {
    g_last_error = msg;
    return -1;
}
static int check_args(const void *out, const void *in, int stride, int min_stride, int n_states) {
    if (n_states < 0) return fail_msg("num_timesteps must be >= 0");
    if (n_states == 0) return 0;
    if (!out || !in) return fail_msg("null device pointer");
    if (stride < min_stride) return fail_msg("stride is smaller than the words the algorithm reads per state");
    return 0;
}

constexpr int kStreams = 4;
constexpr int kMinChunk = 2048;      // states; below this, pipelining costs more than it hides
constexpr size_t kMinChunkBytes = 4u << 20;   // H2D + D2H bytes per chunk

}  // namespace GRID_NS

struct grid_data {
    int cap;
    int device;                      // the device the buffers and streams belong to
    float *h_q_qd_u, *h_q_qd, *h_q, *h_c, *h_Minv, *h_qdd, *h_dc_du, *h_df_du;
    float *d_q_qd_u, *d_q_qd, *d_q, *d_c, *d_Minv, *d_qdd, *d_dc_du, *d_df_du;
    // consumer buffers (fused FD-gradient consumers), allocated on first use
    float *h_lambda, *h_vjp, *h_lin;
    float *d_lambda, *d_vjp, *d_lin;
    cudaStream_t streams[GRID_NS::kStreams];
};

extern "C" {

int grid_abi_version(void) { return GRID_B200_ABI_VERSION; }
int grid_num_joints(void) { return GRID_N; }
const char *grid_robot_name(void) { return GRID_ROBOT_NAME; }
const char *grid_robot_hash(void) { return GRID_ROBOT_HASH; }
const char *grid_last_error(void) { return GRID_NS::g_last_error.c_str(); }
const char *grid_kernel_kind(const char *alg) { return GRID_NS::gen::kernel_kind(alg); }
long long grid_traced_flops(const char *alg) { return GRID_NS::gen::traced_flops(alg); }
/* GRID_FORCE_KERNEL / GRID_PIPE_MODE / GRID_PIPE_CHUNK after load (value NULL or "" = default) */
int grid_set_option(const char *key, const char *value) {
    if (!key) return GRID_NS::fail_msg("grid_set_option: null key");
    GRID_NS::Options &o = GRID_NS::options();
    const bool unset = !value || !*value;
    if (!strcmp(key, "GRID_FORCE_KERNEL")) {
        const int f = GRID_NS::parse_force(value);
        if (f < 0) return GRID_NS::fail_msg("GRID_FORCE_KERNEL must be tps, wps, cps, pipe, lps or empty");
        o.force_kernel = f;
    } else if (!strcmp(key, "GRID_PIPE_MODE")) {
        if (!unset && strcmp(value, "fused") && strcmp(value, "staged"))
            return GRID_NS::fail_msg("GRID_PIPE_MODE must be staged, fused or empty");
        o.pipe_fused = (!unset && !strcmp(value, "fused")) ? 1 : 0;
    } else if (!strcmp(key, "GRID_PIPE_CHUNK")) {
        o.pipe_chunk = unset ? -1 : atoi(value) / 32 * 32;
    } else if (!strcmp(key, "GRID_PIPE_WARPS")) {
        o.pipe_warps = unset ? 0 : atoi(value);
    } else if (!strcmp(key, "GRID_PIPE_ORDER_CHUNK")) {
        o.pipe_order_chunk = unset ? -1 : atoi(value);
    } else if (!strcmp(key, "GRID_PIPE_ONLY_TASK")) {
        o.pipe_only_task = unset ? -1 : atoi(value);
    } else if (!strcmp(key, "GRID_PIPE_STAGGER_NS")) {
        o.pipe_stagger_ns = unset ? 0 : atoi(value);
    } else {
        return GRID_NS::fail_msg("grid_set_option: unknown key");
    }
    return 0;
}

/* kernels launched: a call served by the phase-split kernels launches one kernel per stage */
long long grid_launch_count(void) {
    long long n = GRID_NS::g_launches.load();
#ifdef GRID_HAS_PIPE
    n += GRID_NS::pipe::g_kernel_launches.load() - GRID_NS::pipe::g_calls.load();
#endif
#ifdef GRID_HAS_LPS
    n += GRID_NS::lps::g_kernel_launches.load() - GRID_NS::lps::g_calls.load();
#endif
    return n;
}

#define GRID_LAUNCH(expr, name)                                         \
    do {                                                                \
        cudaError_t e__ = (expr);                                       \
        if (e__ != cudaSuccess) return GRID_NS::fail(name, e__);       \
        GRID_NS::g_launches.fetch_add(1);                              \
    } while (0)

int grid_inverse_dynamics_device(float *d_c, const float *d_q_qd, int stride, const float *d_qdd,
                                 int num_timesteps, float gravity, void *stream) {
    if (int rc = GRID_NS::check_args(d_c, d_q_qd, stride, 2 * GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    GRID_LAUNCH(GRID_NS::gen::launch_id(d_c, d_q_qd, stride, d_qdd, num_timesteps, gravity, (cudaStream_t)stream),
                "inverse_dynamics_kernel");
    return 0;
}

int grid_direct_minv_device(float *d_Minv, const float *d_q, int stride, int num_timesteps, void *stream) {
    if (int rc = GRID_NS::check_args(d_Minv, d_q, stride, GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    GRID_LAUNCH(GRID_NS::gen::launch_minv(d_Minv, d_q, stride, num_timesteps, (cudaStream_t)stream),
                "direct_minv_kernel");
    return 0;
}

int grid_forward_dynamics_device(float *d_qdd, const float *d_q_qd_u, int stride, int num_timesteps,
                                 float gravity, void *stream) {
    if (int rc = GRID_NS::check_args(d_qdd, d_q_qd_u, stride, 3 * GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    GRID_LAUNCH(GRID_NS::gen::launch_fd(d_qdd, d_q_qd_u, stride, num_timesteps, gravity, (cudaStream_t)stream),
                "forward_dynamics_kernel");
    return 0;
}

/* further algorithms: thread-per-state programs where the robot has them */
int grid_crba_device(float *d_M, const float *d_q, int stride, int num_timesteps, void *stream) {
    if (int rc = GRID_NS::check_args(d_M, d_q, stride, GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    cudaError_t e = GRID_NS::gen::launch_crba(d_M, d_q, stride, num_timesteps, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported)
        return GRID_NS::fail_msg("no mass-matrix (CRBA) program for this robot: grid_kernel_kind(\"crba\") is \"none\"");
    if (e != cudaSuccess) return GRID_NS::fail("crba_kernel", e);
    GRID_NS::g_launches.fetch_add(1);
    return 0;
}

int grid_aba_device(float *d_qdd, const float *d_q_qd_u, int stride, int num_timesteps, float gravity, void *stream) {
    if (int rc = GRID_NS::check_args(d_qdd, d_q_qd_u, stride, 3 * GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    cudaError_t e = GRID_NS::gen::launch_aba(d_qdd, d_q_qd_u, stride, num_timesteps, gravity, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported)
        return GRID_NS::fail_msg("no articulated-body (ABA) program for this robot: grid_kernel_kind(\"aba\") is \"none\"");
    if (e != cudaSuccess) return GRID_NS::fail("aba_kernel", e);
    GRID_NS::g_launches.fetch_add(1);
    return 0;
}

int grid_inverse_dynamics_gradient_device(float *d_dc_du, const float *d_q_qd, int stride, const float *d_qdd,
                                          int num_timesteps, float gravity, void *stream) {
    if (int rc = GRID_NS::check_args(d_dc_du, d_q_qd, stride, 2 * GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    GRID_LAUNCH(GRID_NS::gen::launch_id_grad(d_dc_du, d_q_qd, stride, d_qdd, num_timesteps, gravity,
                                              (cudaStream_t)stream),
                "inverse_dynamics_gradient_kernel");
    return 0;
}

int grid_forward_dynamics_gradient_device(float *d_df_du, const float *d_q_qd_u, int stride, const float *d_qdd,
                                          const float *d_Minv, int num_timesteps, float gravity, void *stream) {
    const bool pre = d_qdd != nullptr || d_Minv != nullptr;
    if (pre && !(d_qdd && d_Minv)) return GRID_NS::fail_msg("d_qdd and d_Minv must both be given or both be NULL");
    if (int rc = GRID_NS::check_args(d_df_du, d_q_qd_u, stride, (pre ? 2 : 3) * GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    GRID_LAUNCH(GRID_NS::gen::launch_fd_grad(d_df_du, d_q_qd_u, stride, d_qdd, d_Minv, num_timesteps, gravity,
                                              (cudaStream_t)stream),
                "forward_dynamics_gradient_kernel");
    return 0;
}

int grid_forward_dynamics_gradient_vjp_device(float *d_out, const float *d_q_qd_u, int stride, const float *d_lambda,
                                              int num_timesteps, float dt, float gravity, void *stream) {
    if (int rc = GRID_NS::check_args(d_out, d_q_qd_u, stride, 3 * GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    if (!d_lambda) return GRID_NS::fail_msg("null d_lambda");
    cudaError_t e = GRID_NS::gen::launch_fd_vjp(d_out, d_q_qd_u, stride, d_lambda, num_timesteps, gravity, dt,
                                                (cudaStream_t)stream);
    if (e == cudaErrorNotSupported)
        return GRID_NS::fail_msg("fused FD-gradient consumers need a thread-per-state or phase-split FD-gradient "
                                 "kernel; this robot has neither");
    if (e != cudaSuccess) return GRID_NS::fail("forward_dynamics_gradient_vjp_kernel", e);
    GRID_NS::g_launches.fetch_add(1);
    return 0;
}

int grid_forward_dynamics_linearize_device(float *d_out, const float *d_q_qd_u, int stride, int num_timesteps, float dt,
                                           float gravity, void *stream) {
    if (int rc = GRID_NS::check_args(d_out, d_q_qd_u, stride, 3 * GRID_N, num_timesteps)) return rc;
    if (num_timesteps == 0) return 0;
    cudaError_t e = GRID_NS::gen::launch_fd_lin(d_out, d_q_qd_u, stride, num_timesteps, gravity, dt, (cudaStream_t)stream);
    if (e == cudaErrorNotSupported)
        return GRID_NS::fail_msg("fused FD-gradient consumers need a thread-per-state or phase-split FD-gradient "
                                 "kernel; this robot has neither");
    if (e != cudaSuccess) return GRID_NS::fail("forward_dynamics_linearize_kernel", e);
    GRID_NS::g_launches.fetch_add(1);
    return 0;
}

/* ---- gridData handle ------------------------------------------------------------------ */
#define GRID_CU(expr, where)                                            \
    do {                                                                \
        cudaError_t e__ = (expr);                                       \
        if (e__ != cudaSuccess) return GRID_NS::fail(where, e__);      \
    } while (0)

static int grid_data_alloc(grid_data *hd, int T) {
    const size_t n = GRID_N, f = sizeof(float);
    GRID_CU(cudaGetDevice(&hd->device), "cudaGetDevice");
    struct { float **h, **d; size_t words; } bufs[] = {
        {&hd->h_q_qd_u, &hd->d_q_qd_u, 3 * n}, {&hd->h_q_qd, &hd->d_q_qd, 2 * n}, {&hd->h_q, &hd->d_q, n},
        {&hd->h_c, &hd->d_c, n}, {&hd->h_Minv, &hd->d_Minv, n * n}, {&hd->h_qdd, &hd->d_qdd, n},
        {&hd->h_dc_du, &hd->d_dc_du, 2 * n * n}, {&hd->h_df_du, &hd->d_df_du, 2 * n * n}};
    for (auto &b : bufs) {
        GRID_CU(cudaMallocHost((void **)b.h, b.words * T * f), "cudaMallocHost");
        GRID_CU(cudaMalloc((void **)b.d, b.words * T * f), "cudaMalloc");
        memset(*b.h, 0, b.words * T * f);
    }
    int lo = 0, hi = 0;
    GRID_CU(cudaDeviceGetStreamPriorityRange(&lo, &hi), "cudaDeviceGetStreamPriorityRange");
    for (int i = 0; i < GRID_NS::kStreams; i++)
        GRID_CU(cudaStreamCreateWithPriority(&hd->streams[i], cudaStreamNonBlocking, hi), "cudaStreamCreate");
    return 0;
}

grid_data *grid_data_create(int max_timesteps) {
    if (max_timesteps <= 0) {
        GRID_NS::fail_msg("max_timesteps must be positive");
        return nullptr;
    }
    grid_data *hd = new grid_data();
    memset(hd, 0, sizeof(*hd));
    hd->cap = max_timesteps;
    if (grid_data_alloc(hd, max_timesteps) != 0) {
        std::string keep = GRID_NS::g_last_error;
        grid_data_destroy(hd);
        GRID_NS::g_last_error = keep;
        return nullptr;
    }
    return hd;
}

void grid_data_destroy(grid_data *hd) {
    if (!hd) return;
    float *hs[] = {hd->h_q_qd_u, hd->h_q_qd, hd->h_q, hd->h_c, hd->h_Minv, hd->h_qdd, hd->h_dc_du, hd->h_df_du,
                   hd->h_lambda, hd->h_vjp, hd->h_lin};
    float *ds[] = {hd->d_q_qd_u, hd->d_q_qd, hd->d_q, hd->d_c, hd->d_Minv, hd->d_qdd, hd->d_dc_du, hd->d_df_du,
                   hd->d_lambda, hd->d_vjp, hd->d_lin};
    for (float *p : hs) if (p) cudaFreeHost(p);
    for (float *p : ds) if (p) cudaFree(p);
    for (auto s : hd->streams) if (s) cudaStreamDestroy(s);
    delete hd;
}

int grid_data_capacity(const grid_data *hd) { return hd ? hd->cap : 0; }

/* consumer buffers are allocated on first use: h_lin is (2n + 3n^2) floats per state */
static int grid_data_alloc_consumers(grid_data *hd) {
    if (hd->h_lambda) return 0;
    const size_t n = GRID_N, f = sizeof(float), T = hd->cap;
    struct { float **h, **d; size_t words; } bufs[] = {
        {&hd->h_lambda, &hd->d_lambda, 2 * n}, {&hd->h_vjp, &hd->d_vjp, 5 * n}, {&hd->h_lin, &hd->d_lin, 2 * n + 3 * n * n}};
    for (auto &b : bufs) {
        GRID_CU(cudaMallocHost((void **)b.h, b.words * T * f), "cudaMallocHost");
        GRID_CU(cudaMalloc((void **)b.d, b.words * T * f), "cudaMalloc");
        memset(*b.h, 0, b.words * T * f);
    }
    return 0;
}

float *grid_data_ptr(grid_data *hd, const char *field) {
    if (!hd || !field) return nullptr;
    if (!strcmp(field + (field[0] ? 2 : 0), "lambda") || !strcmp(field + (field[0] ? 2 : 0), "vjp") ||
        !strcmp(field + (field[0] ? 2 : 0), "lin")) {
        if (grid_data_alloc_consumers(hd) != 0) return nullptr;
        struct { const char *name; float *p; } ctab[] = {
            {"h_lambda", hd->h_lambda}, {"h_vjp", hd->h_vjp}, {"h_lin", hd->h_lin},
            {"d_lambda", hd->d_lambda}, {"d_vjp", hd->d_vjp}, {"d_lin", hd->d_lin}};
        for (auto &t : ctab) if (!strcmp(t.name, field)) return t.p;
        return nullptr;
    }
    struct { const char *name; float *p; } tab[] = {
        {"h_q_qd_u", hd->h_q_qd_u}, {"h_q_qd", hd->h_q_qd}, {"h_q", hd->h_q}, {"h_c", hd->h_c},
        {"h_Minv", hd->h_Minv}, {"h_qdd", hd->h_qdd}, {"h_dc_du", hd->h_dc_du}, {"h_df_du", hd->h_df_du},
        {"d_q_qd_u", hd->d_q_qd_u}, {"d_q_qd", hd->d_q_qd}, {"d_q", hd->d_q}, {"d_c", hd->d_c},
        {"d_Minv", hd->d_Minv}, {"d_qdd", hd->d_qdd}, {"d_dc_du", hd->d_dc_du}, {"d_df_du", hd->d_df_du}};
    for (auto &t : tab) if (!strcmp(t.name, field)) return t.p;
    return nullptr;
}

}  // extern "C"

namespace GRID_NS {

struct Span { const float *h; float *d; size_t words; };          // per-state words of an input
struct OutSpan { float *h; float *d; size_t words; };

// Pipelines [H2D inputs | kernel | D2H output] per chunk of states over kStreams streams.
template <class Launch>
static int run_pipelined(grid_data *hd, int T, const Span *ins, int n_ins, const OutSpan *outs, int n_outs,
                         Launch launch) {
    if (!hd) return fail_msg("null grid_data");
    if (T < 0 || T > hd->cap) return fail_msg("num_timesteps exceeds the grid_data capacity");
    if (T == 0) return 0;
    int dev = -1;
    GRID_CU(cudaGetDevice(&dev), "cudaGetDevice");
    if (dev != hd->device)
        return fail_msg("this grid_data was created on another device: cudaSetDevice() to it before calling");
    // chunks: at most 2 per stream, at least kMinChunk states and ~kMinChunkBytes of traffic each - every
    // cudaMemcpyAsync costs ~10 us of fixed DMA set-up, which 16 half-megabyte copies would not amortise
    // (fused consumers move 140 B per iiwa14 state: 4 chunks beat 8 by 15 %, profiles/r2_*)
    int chunks = (T + kMinChunk - 1) / kMinChunk;
    if (chunks > 2 * kStreams) chunks = 2 * kStreams;
    size_t words = 0;
    for (int i = 0; i < n_ins; i++) words += ins[i].words;
    for (int o = 0; o < n_outs; o++) words += outs[o].words;
    const size_t by_bytes = (size_t)T * words * sizeof(float) / kMinChunkBytes;
    if ((size_t)chunks > by_bytes) chunks = by_bytes < 1 ? 1 : (int)by_bytes;
    const int per = (T + chunks - 1) / chunks;
    // On the first failure: stop enqueueing, but still wait for every stream - copies that are already
    // in flight use the caller's buffers, which the caller is free to release once this returns.
    int rc = 0;
    auto cu = [&](cudaError_t e, const char *where) {
        if (e != cudaSuccess && rc == 0) rc = fail(where, e);
        return e == cudaSuccess;
    };
    for (int c = 0; c < chunks && rc == 0; c++) {
        const size_t first = (size_t)c * per;
        if (first >= (size_t)T) break;
        const int cnt = (int)((first + per <= (size_t)T) ? per : (T - first));
        cudaStream_t s = hd->streams[c % kStreams];
        bool ok = true;
        for (int i = 0; i < n_ins && ok; i++)
            ok = cu(cudaMemcpyAsync(ins[i].d + first * ins[i].words, ins[i].h + first * ins[i].words,
                                    ins[i].words * cnt * sizeof(float), cudaMemcpyHostToDevice, s), "H2D");
        if (!ok) break;
        if (int r = launch(first, cnt, s)) { rc = r; break; }
        for (int o = 0; o < n_outs && ok; o++)
            ok = cu(cudaMemcpyAsync(outs[o].h + first * outs[o].words, outs[o].d + first * outs[o].words,
                                    outs[o].words * cnt * sizeof(float), cudaMemcpyDeviceToHost, s), "D2H");
    }
    std::string keep = g_last_error;
    for (auto s : hd->streams) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess && rc == 0) { rc = fail("cudaStreamSynchronize", e); keep = g_last_error; }
    }
    if (rc != 0) g_last_error = keep;
    return rc;
}

template <class Launch>
static int run_pipelined(grid_data *hd, int T, const Span *ins, int n_ins, OutSpan out, Launch launch) {
    return run_pipelined(hd, T, ins, n_ins, &out, 1, launch);
}

}  // namespace GRID_NS

extern "C" {

int grid_inverse_dynamics(grid_data *hd, int T, float gravity, int use_qdd, int compressed) {
    if (!hd) return GRID_NS::fail_msg("null grid_data");
    const size_t n = GRID_N, st = compressed ? 2 * n : 3 * n;
    GRID_NS::Span ins[2] = {{compressed ? hd->h_q_qd : hd->h_q_qd_u, compressed ? hd->d_q_qd : hd->d_q_qd_u, st},
                             {hd->h_qdd, hd->d_qdd, n}};
    float *d_in = ins[0].d;
    return GRID_NS::run_pipelined(hd, T, ins, use_qdd ? 2 : 1, {hd->h_c, hd->d_c, n},
        [&](size_t first, int cnt, cudaStream_t s) {
            return grid_inverse_dynamics_device(hd->d_c + first * n, d_in + first * st, (int)st,
                                                use_qdd ? hd->d_qdd + first * n : nullptr, cnt, gravity, s);
        });
}

int grid_direct_minv(grid_data *hd, int T, int compressed) {
    if (!hd) return GRID_NS::fail_msg("null grid_data");
    const size_t n = GRID_N, st = compressed ? n : 3 * n;
    GRID_NS::Span ins[1] = {{compressed ? hd->h_q : hd->h_q_qd_u, compressed ? hd->d_q : hd->d_q_qd_u, st}};
    float *d_in = ins[0].d;
    return GRID_NS::run_pipelined(hd, T, ins, 1, {hd->h_Minv, hd->d_Minv, n * n},
        [&](size_t first, int cnt, cudaStream_t s) {
            return grid_direct_minv_device(hd->d_Minv + first * n * n, d_in + first * st, (int)st, cnt, s);
        });
}

int grid_forward_dynamics(grid_data *hd, int T, float gravity) {
    if (!hd) return GRID_NS::fail_msg("null grid_data");
    const size_t n = GRID_N, st = 3 * n;
    GRID_NS::Span ins[1] = {{hd->h_q_qd_u, hd->d_q_qd_u, st}};
    return GRID_NS::run_pipelined(hd, T, ins, 1, {hd->h_qdd, hd->d_qdd, n},
        [&](size_t first, int cnt, cudaStream_t s) {
            return grid_forward_dynamics_device(hd->d_qdd + first * n, hd->d_q_qd_u + first * st, (int)st, cnt,
                                                gravity, s);
        });
}

int grid_inverse_dynamics_gradient(grid_data *hd, int T, float gravity, int use_qdd, int compressed) {
    if (!hd) return GRID_NS::fail_msg("null grid_data");
    const size_t n = GRID_N, st = compressed ? 2 * n : 3 * n;
    GRID_NS::Span ins[2] = {{compressed ? hd->h_q_qd : hd->h_q_qd_u, compressed ? hd->d_q_qd : hd->d_q_qd_u, st},
                             {hd->h_qdd, hd->d_qdd, n}};
    float *d_in = ins[0].d;
    return GRID_NS::run_pipelined(hd, T, ins, use_qdd ? 2 : 1, {hd->h_dc_du, hd->d_dc_du, 2 * n * n},
        [&](size_t first, int cnt, cudaStream_t s) {
            return grid_inverse_dynamics_gradient_device(hd->d_dc_du + first * 2 * n * n, d_in + first * st, (int)st,
                                                         use_qdd ? hd->d_qdd + first * n : nullptr, cnt, gravity, s);
        });
}

int grid_forward_dynamics_gradient(grid_data *hd, int T, float gravity, int use_qdd_minv) {
    if (!hd) return GRID_NS::fail_msg("null grid_data");
    const size_t n = GRID_N, st = 3 * n;
    GRID_NS::Span ins[3] = {{hd->h_q_qd_u, hd->d_q_qd_u, st}, {hd->h_qdd, hd->d_qdd, n}, {hd->h_Minv, hd->d_Minv, n * n}};
    return GRID_NS::run_pipelined(hd, T, ins, use_qdd_minv ? 3 : 1, {hd->h_df_du, hd->d_df_du, 2 * n * n},
        [&](size_t first, int cnt, cudaStream_t s) {
            return grid_forward_dynamics_gradient_device(
                hd->d_df_du + first * 2 * n * n, hd->d_q_qd_u + first * st, (int)st,
                use_qdd_minv ? hd->d_qdd + first * n : nullptr, use_qdd_minv ? hd->d_Minv + first * n * n : nullptr,
                cnt, gravity, s);
        });
}

/* host forms of the fused consumers: O(n) words per state come back instead of 2n^2 */
int grid_forward_dynamics_gradient_vjp(grid_data *hd, int T, float dt, float gravity) {
    if (!hd) return GRID_NS::fail_msg("null grid_data");
    if (int rc = grid_data_alloc_consumers(hd)) return rc;
    const size_t n = GRID_N, st = 3 * n;
    GRID_NS::Span ins[2] = {{hd->h_q_qd_u, hd->d_q_qd_u, st}, {hd->h_lambda, hd->d_lambda, 2 * n}};
    return GRID_NS::run_pipelined(hd, T, ins, 2, {hd->h_vjp, hd->d_vjp, 5 * n},
        [&](size_t first, int cnt, cudaStream_t s) {
            return grid_forward_dynamics_gradient_vjp_device(hd->d_vjp + first * 5 * n, hd->d_q_qd_u + first * st, (int)st,
                                                             hd->d_lambda + first * 2 * n, cnt, dt, gravity, s);
        });
}

int grid_forward_dynamics_linearize(grid_data *hd, int T, float dt, float gravity) {
    if (!hd) return GRID_NS::fail_msg("null grid_data");
    if (int rc = grid_data_alloc_consumers(hd)) return rc;
    const size_t n = GRID_N, st = 3 * n, ow = 2 * n + 3 * n * n;
    GRID_NS::Span ins[1] = {{hd->h_q_qd_u, hd->d_q_qd_u, st}};
    return GRID_NS::run_pipelined(hd, T, ins, 1, {hd->h_lin, hd->d_lin, ow},
        [&](size_t first, int cnt, cudaStream_t s) {
            return grid_forward_dynamics_linearize_device(hd->d_lin + first * ow, hd->d_q_qd_u + first * st, (int)st, cnt,
                                                          dt, gravity, s);
        });
}

}  // extern "C"

namespace GRID_NS {
__global__ void noop_kernel() {}
}  // namespace GRID_NS

namespace GRID_NS {
// one launch of an algorithm by name; d_in1 = qdd (id, id_grad, fd_grad) or lambda (fd_vjp), d_in2 = Minv (fd_grad)
static int launch_by_name(const char *alg, float *d_out, const float *d_in, int stride, const float *d_in1,
                          const float *d_in2, int N, float dt, float gravity, cudaStream_t s) {
    if (!strcmp(alg, "id")) return grid_inverse_dynamics_device(d_out, d_in, stride, d_in1, N, gravity, s);
    if (!strcmp(alg, "minv")) return grid_direct_minv_device(d_out, d_in, stride, N, s);
    if (!strcmp(alg, "fd")) return grid_forward_dynamics_device(d_out, d_in, stride, N, gravity, s);
    if (!strcmp(alg, "aba")) return grid_aba_device(d_out, d_in, stride, N, gravity, s);
    if (!strcmp(alg, "crba")) return grid_crba_device(d_out, d_in, stride, N, s);
    if (!strcmp(alg, "id_grad")) return grid_inverse_dynamics_gradient_device(d_out, d_in, stride, d_in1, N, gravity, s);
    if (!strcmp(alg, "fd_grad")) return grid_forward_dynamics_gradient_device(d_out, d_in, stride, d_in1, d_in2, N, gravity, s);
    if (!strcmp(alg, "fd_vjp")) return grid_forward_dynamics_gradient_vjp_device(d_out, d_in, stride, d_in1, N, dt, gravity, s);
    if (!strcmp(alg, "fd_lin")) return grid_forward_dynamics_linearize_device(d_out, d_in, stride, N, dt, gravity, s);
    if (!strcmp(alg, "noop")) {                     // floor of a timing method: an empty kernel
        noop_kernel<<<1, 32, 0, s>>>();
        return cudaGetLastError() == cudaSuccess ? 0 : fail_msg("noop launch failed");
    }
    return fail_msg("unknown algorithm name");
}
}  // namespace GRID_NS

/* ---- CUDA-graph entry for repeated fixed-shape calls (trajectory optimisers call the same
 * shape every iteration): the launch - for phase-split kernels: scratch allocation, ticket memset,
 * two kernels, free - is captured once and replayed with one cudaGraphLaunch. ---------------- */
struct grid_graph {
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int device;
};

extern "C" grid_graph *grid_graph_create(const char *alg, float *d_out, const float *d_in, int stride, const float *d_in1,
                                         const float *d_in2, int num_timesteps, float dt, float gravity) {
    using namespace GRID_NS;
    if (!alg || num_timesteps <= 0) { fail_msg("grid_graph_create: bad arguments"); return nullptr; }
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { fail_msg("cudaStreamCreate failed"); return nullptr; }
    grid_graph *g = nullptr;
    // one eager launch first: per-device launcher state (shared-memory opt-in, occupancy) is set up outside capture
    int rc = launch_by_name(alg, d_out, d_in, stride, d_in1, d_in2, num_timesteps, dt, gravity, s);
    if (rc == 0 && cudaStreamSynchronize(s) != cudaSuccess) rc = fail_msg("warm-up launch failed");
    if (rc == 0) {
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            rc = launch_by_name(alg, d_out, d_in, stride, d_in1, d_in2, num_timesteps, dt, gravity, s);
            e = cudaStreamEndCapture(s, &graph);
        }
        if (e != cudaSuccess) rc = fail("grid_graph_create capture", e);
        if (rc == 0) {
            cudaGraphExec_t exec = nullptr;
            e = cudaGraphInstantiate(&exec, graph, 0);
            if (e != cudaSuccess) {
                rc = fail("cudaGraphInstantiate", e);
            } else {
                g = new grid_graph{graph, exec, 0};
                cudaGetDevice(&g->device);
                graph = nullptr;
            }
        }
        if (graph) cudaGraphDestroy(graph);
    }
    cudaStreamDestroy(s);
    return g;
}

extern "C" int grid_graph_launch(grid_graph *g, void *stream) {
    if (!g) return GRID_NS::fail_msg("null grid_graph");
    cudaError_t e = cudaGraphLaunch(g->exec, (cudaStream_t)stream);
    if (e != cudaSuccess) return GRID_NS::fail("cudaGraphLaunch", e);
    GRID_NS::g_launches.fetch_add(1);
    return 0;
}

extern "C" void grid_graph_destroy(grid_graph *g) {
    if (!g) return;
    cudaGraphExecDestroy(g->exec);
    cudaGraphDestroy(g->graph);
    delete g;
}

/* alg may carry the suffix "@graph": the launch is then captured once and the timed launches replay it */
extern "C" int grid_time_launches(const char *alg, float *d_out, const float *d_in, int stride, int num_timesteps,
                                  float gravity, int reps, float *h_us) {
    using namespace GRID_NS;
    if (!alg || !h_us || reps <= 0 || reps > 100000) return fail_msg("bad arguments to grid_time_launches");
    std::string name(alg);
    const size_t at = name.find("@graph");
    const bool use_graph = at != std::string::npos;
    if (use_graph) name.resize(at);
    grid_graph *gr = nullptr;
    if (use_graph) {
        gr = grid_graph_create(name.c_str(), d_out, d_in, stride, nullptr, nullptr, num_timesteps, 0.f, gravity);
        if (!gr) return -1;
    }
    cudaStream_t s = nullptr;
    GRID_CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    cudaEvent_t *ev = new cudaEvent_t[2 * (size_t)reps];
    for (int i = 0; i < 2 * reps; i++) cudaEventCreate(&ev[i]);
    int rc = 0;
    for (int i = -3; i < reps && rc == 0; i++) {          // 3 warm-up launches
        if (i >= 0) cudaEventRecord(ev[2 * i], s);
        rc = gr ? grid_graph_launch(gr, s)
                : launch_by_name(name.c_str(), d_out, d_in, stride, nullptr, nullptr, num_timesteps, 0.f, gravity, s);
        if (i >= 0) cudaEventRecord(ev[2 * i + 1], s);
    }
    if (rc == 0) {
        cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) rc = fail("grid_time_launches", e);
        for (int i = 0; i < reps && rc == 0; i++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[2 * i], ev[2 * i + 1]);
            h_us[i] = ms * 1e3f;
        }
    }
    for (int i = 0; i < 2 * reps; i++) cudaEventDestroy(ev[i]);
    delete[] ev;
    cudaStreamDestroy(s);
    if (gr) grid_graph_destroy(gr);
    return rc;
}

/* ---- FP32 roofline microbenchmark ------------------------------------------------------ */
namespace GRID_NS {
__global__ void __launch_bounds__(256) fp32_peak_kernel(float *out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; k++) x[k] = threadIdx.x * 1e-3f + k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int k = 0; k < 8; k++) x[k] = fmaf(x[k], a, b);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; k++) s += x[k];
    if (s == 123.456f) out[0] = s;
}
}  // namespace GRID_NS

extern "C" double grid_measure_fp32_tflops(int repeats) {
    using namespace GRID_NS;
    float *d = nullptr;
    if (cudaMalloc(&d, 4) != cudaSuccess) { fail_msg("cudaMalloc failed in grid_measure_fp32_tflops"); return -1.0; }
    cudaDeviceProp prop;
    int dev = 0;
    cudaGetDevice(&dev);
    cudaGetDeviceProperties(&prop, dev);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    if (repeats < 1) repeats = 1;
    for (int r = 0; r < repeats + 2; r++) {
        cudaEventRecord(e0);
        fp32_peak_kernel<<<blocks, threads>>>(d, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
        g_launches.fetch_add(1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (r >= 2 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    return best;
}
