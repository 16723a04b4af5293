// grid_lps.cuh - serial chains of any length: lane = state, ROLLED loops over the joints.
//
// The traced straight-line programs stop at ~20 k operations per thread; a 64-link chain's Minv or a single
// gradient column is longer than that, and round 1 served such robots with the CTA-per-state kernels
// (grid_wps.cuh): one CTA of 4-5 warps per SM, 176 KB of shared memory per state, 22.8 ms for the FD gradient
// of 16 384 states of the 64-link chain (3.4 % of the FP32 roofline).  These kernels keep the decomposition
// of the phase-split kernels - a per-state program followed by one program per du-column, 32 consecutive
// states per warp - but write the programs as loops over the joint index with the joint's constants read from
// constant memory (the index is warp-uniform), so the code is a few KB and stays in the instruction cache,
// every lane works on its own state (no shuffles, no barriers inside a program), and a batch exposes
// states x columns independent warps instead of states CTAs.
//
//   stage A (warp = 32 states): X_i(q), the composite base transforms, RNEA bias forces, the articulated-body
//            inertia pass (U_i, Dinv_i), qdd by an O(n) articulated-body solve (no Minv needed), RNEA at qdd;
//            per-joint results go to a scratch array [tile][joint][word][lane] (one 128-byte line per access).
//   Minv columns (warp = 32 states x column j): the column recursions of reference algorithms/_direct_minv.py
//            restricted to one column: backward j..0, forward 0..j.
//   gradient columns (warp = 32 states x du-column): the forward recursion of reference
//            algorithms/_inverse_dynamics_gradient.py:189-430 for one column.  The BACKWARD accumulation
//            (:477-541) needs every df_i of the column again in reverse order (the wide kernel stores them:
//            100 KB per state); here each df_i is moved to the BASE frame as soon as it exists
//            (f0_i = (iX0)^T df_i), so the accumulated force at joint i is a suffix sum,
//            dc_i = s0_i . (Total - sum_{k<i} f0_k), and one scalar per joint is all that is kept.
//            For the FD gradient the column is then multiplied by -Minv WITHOUT Minv: an O(n)
//            articulated-body solve with the U_i, Dinv_i of stage A replaces the n x n product of
//            reference algorithms/_forward_dynamics_gradient.py:48-57 (4 n^3 flops per state, the only
//            GEMM-shaped step of the path) - which is also this repo's answer to "tensor cores?" for chains.
//
// Needs, from the generated translation unit (GRID_NS::gen): struct WT (N) and the __constant__ tables wt_S,
// wt_E0, wt_r0, wt_I, wt_damping (emit_wps_tables); and grid_wps.cuh for the spatial-algebra helpers.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include "grid_wps.cuh"
#define GRID_HAS_LPS 1

namespace GRID_NS { namespace lps {

using namespace gen;
using wps::Xf;
using wps::xmotion;
using wps::xtforce;
using wps::mxS;
using wps::crossf;
using wps::imul;
using wps::pick;
using wps::add_at;
using wps::cross3;
using wps::mat3T;

constexpr int N = WT::N;
constexpr int W = 64;                  // scratch words per joint and state
// word offsets inside a joint's block
constexpr int wE = 0, wR = 9;          // X_i(q): E (row-major 3x3), r
constexpr int wV = 12, wIV = 18;       // v_i, I_i v_i
constexpr int wMXA = 24, wMF = 30;     // mxS(X a_parent), mxS(f_i)   (wMF doubles as temporary f_i storage)
constexpr int wE0 = 36, wR0 = 45;      // composite iX0: base -> joint i
constexpr int wS0 = 48;                // joint axis in base coordinates (motion vector)
constexpr int wU = 54, wDINV = 60;     // articulated-body U_i = IA_i S_i, 1 / (S_i^T U_i)
constexpr int wQD = 61, wY = 62, wQDD = 63;
constexpr int PITCH = 33;              // row pitch of the per-warp [joint][lane] staging tile
constexpr int kColWarps = 8;
// Stage A is latency-bound (one serial recursion per state; 218 us whether a launch holds 128 or 512 warps), so it
// runs over chunks of 16 384 states; the column kernels walk a chunk in SUB-CHUNKS of 4 096 states (128 tiles) in
// block order, so that the scratch lines they re-read ~20 times (67 MB per sub-chunk of the 64-link chain) stay in L2.
constexpr int kSubTiles = 128;
constexpr int kChunkStates = 4 * kSubTiles * 32;

static std::atomic<long long> g_kernel_launches{0};
static std::atomic<long long> g_calls{0};

__device__ __forceinline__ float *joint_ptr(float *tile_lane, int i) { return tile_lane + (size_t)i * (W * 32); }
__device__ __forceinline__ void ld6(const float *p, float *o) {
#pragma unroll
    for (int r = 0; r < 6; r++) o[r] = p[32 * r];
}
__device__ __forceinline__ void st6(float *p, const float *o) {
#pragma unroll
    for (int r = 0; r < 6; r++) p[32 * r] = o[r];
}
__device__ __forceinline__ Xf ldX(const float *p) {       // p = joint block + wE (or wE0)
    Xf x;
#pragma unroll
    for (int e = 0; e < 9; e++) x.E[e] = p[32 * e];
#pragma unroll
    for (int e = 0; e < 3; e++) x.r[e] = p[32 * (9 + e)];
    return x;
}
__device__ __forceinline__ void stX(float *p, const Xf &x) {
#pragma unroll
    for (int e = 0; e < 9; e++) p[32 * e] = x.E[e];
#pragma unroll
    for (int e = 0; e < 3; e++) p[32 * (9 + e)] = x.r[e];
}
__device__ __forceinline__ float dot6(const float *a, const float *b) {
    float s = a[0] * b[0];
#pragma unroll
    for (int r = 1; r < 6; r++) s = fmaf(a[r], b[r], s);
    return s;
}

// X_i(q) = X_joint(q) X_tree as (E, r): the same update the wide kernels do (wps::update_X), per thread
__device__ __forceinline__ Xf joint_X(int i, float q) {
    const int k = wt_S[i];
    const float *E0 = wt_E0 + 9 * i;
    Xf X;
#pragma unroll
    for (int e = 0; e < 9; e++) X.E[e] = E0[e];
    X.r[0] = wt_r0[3 * i]; X.r[1] = wt_r0[3 * i + 1]; X.r[2] = wt_r0[3 * i + 2];
    if (k < 3) {
        float sn, cs;
        sincosf(q, &sn, &cs);
        const int a = (k + 1) % 3, b = (k + 2) % 3;
#pragma unroll
        for (int col = 0; col < 3; col++) {
            const float ea = E0[3 * a + col], eb = E0[3 * b + col];
            const float na = cs * ea + sn * eb, nb = cs * eb - sn * ea;
#pragma unroll
            for (int row = 0; row < 3; row++) {
                if (row == a) X.E[3 * row + col] = na;
                if (row == b) X.E[3 * row + col] = nb;
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < 3; t++) X.r[t] += q * E0[3 * (k - 3) + t];
    }
    return X;
}

// IA <- X^T (IA - U Dinv U^T) X for a symmetric 6x6 (full storage), X = (E, r)
__device__ __forceinline__ void abi_to_parent(const Xf &X, const float *IA, const float *U, float Dinv, float *out) {
#pragma unroll
    for (int c = 0; c < 6; c++) {
        float e[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, x[6], t[6], o[6];
        e[c] = 1.0f;
        xmotion(X, e, x);                              // column c of X
        const float ux = dot6(U, x) * Dinv;
#pragma unroll
        for (int r = 0; r < 6; r++) {
            float acc = -U[r] * ux;
#pragma unroll
            for (int cc = 0; cc < 6; cc++) acc = fmaf(IA[6 * r + cc], x[cc], acc);
            t[r] = acc;
        }
        xtforce(X, t, o);
#pragma unroll
        for (int r = 0; r < 6; r++) out[6 * r + c] = o[r];
    }
}

// ---- stage A ------------------------------------------------------------------------------------
// FLAGS: bit 0 = bias forces c (RNEA at qdd = 0), bit 1 = articulated-body inertias (U, Dinv),
//        bit 2 = qdd = FD(q, qd, u) (needs bits 0 and 1), bit 3 = qdd given in d_qdd,
//        bit 4 = gradient exports (RNEA at qdd -> mxS(X a_parent), mxS(f)), bit 5 = write qdd to d_qdd_out,
//        bit 6 = fused VJP consumer: second solve w = Minv lam_v (lam = d_qdd: [lam_q | lam_v] per state), w kept in
//                the wY word for the column kernel, x+ and B^T lam = dt w written to d_qdd_out (5n words per state),
//        bit 7 = fused linearisation consumer: x+ written to d_qdd_out (2n + 3n^2 words per state)
template <int FLAGS>
__global__ void __launch_bounds__(128)
stage_a_kernel(const float *__restrict__ d_in, int stride, const float *__restrict__ d_qdd, float *__restrict__ scratch,
               float *__restrict__ d_qdd_out, int num_states, float gravity, float dt) {
    constexpr bool C0 = FLAGS & 1, ABI = FLAGS & 2, SOLVE = FLAGS & 4, QDD_IN = FLAGS & 8, GRAD = FLAGS & 16, QDD_OUT = FLAGS & 32;
    constexpr bool VJP = FLAGS & 64, LIN = FLAGS & 128;
    constexpr int CONS_WORDS = VJP ? 5 * N : 2 * N + 3 * N * N;        // output row of a consumer
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int ntiles = (num_states + 31) >> 5;
    if (tile >= ntiles) return;
    const long long st = min((long long)tile * 32 + lane, (long long)num_states - 1);   // ragged tile: recompute the last state
    const bool valid = (long long)tile * 32 + lane < num_states;
    const float *row = d_in + st * stride;
    float *s = scratch + (size_t)tile * (N * W * 32) + lane;

    // pass 1 (forward): X, composite base transform, joint axis in base coordinates, v, I v, f at qdd = 0
    {
        Xf Xc;
        float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < N; i++) {
            const int k = wt_S[i];
            const float q = __ldg(row + i), qd = __ldg(row + N + i);
            const Xf X = joint_X(i, q);
            float *sj = joint_ptr(s, i);
            stX(sj + 32 * wE, X);
            if (i == 0) {
                Xc = X;
            } else {                                   // iX0 = X_i (i-1)X0: E = E_i E_c, r = r_c + E_c^T r_i
                Xf Y;
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = 0; c < 3; c++)
                        Y.E[3 * r + c] = X.E[3 * r] * Xc.E[c] + X.E[3 * r + 1] * Xc.E[3 + c] + X.E[3 * r + 2] * Xc.E[6 + c];
                float t[3];
                mat3T(Xc.E, X.r, t);
                Y.r[0] = Xc.r[0] + t[0]; Y.r[1] = Xc.r[1] + t[1]; Y.r[2] = Xc.r[2] + t[2];
                Xc = Y;
            }
            stX(sj + 32 * wE0, Xc);
            {   // s0 = 0X_i S_i: revolute [w ; r x w] with w = E_c^T e_k, prismatic [0 ; E_c^T e_k]
                const int ax = k < 3 ? k : k - 3;
                const float w[3] = {Xc.E[3 * ax], Xc.E[3 * ax + 1], Xc.E[3 * ax + 2]};
                float s0[6], t[3];
                cross3(Xc.r, w, t);
                const bool rev = k < 3;
                s0[0] = rev ? w[0] : 0.f; s0[1] = rev ? w[1] : 0.f; s0[2] = rev ? w[2] : 0.f;
                s0[3] = rev ? t[0] : w[0]; s0[4] = rev ? t[1] : w[1]; s0[5] = rev ? t[2] : w[2];
                st6(sj + 32 * wS0, s0);
            }
            float vn[6];
            if (i == 0) {
#pragma unroll
                for (int r = 0; r < 6; r++) vn[r] = 0.f;
            } else {
                xmotion(X, v, vn);
            }
            add_at(vn, k, qd);
            float iv[6];
            imul(i, vn, iv);
            st6(sj + 32 * wV, vn);
            st6(sj + 32 * wIV, iv);
            sj[32 * wQD] = qd;
            if (C0) {
                float an[6], t[6], f[6];
                if (i == 0) {
                    float e[6] = {0.f, 0.f, 0.f, 0.f, 0.f, gravity};
                    xmotion(X, e, an);                 // X[:,5] * gravity (algorithms/_inverse_dynamics.py:123)
                } else {
                    xmotion(X, a, an);
                    mxS(k, vn, t);
#pragma unroll
                    for (int r = 0; r < 6; r++) an[r] = fmaf(t[r], qd, an[r]);
                }
                imul(i, an, f);
                crossf(vn, iv, t);
#pragma unroll
                for (int r = 0; r < 6; r++) { f[r] += t[r]; a[r] = an[r]; }
                st6(sj + 32 * wMF, f);
            }
#pragma unroll
            for (int r = 0; r < 6; r++) v[r] = vn[r];
        }
    }
    // pass 2 (backward): c_i, articulated-body inertias, first half of the solve qdd = Minv (u - c)
    if (C0 || ABI) {
        float fc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, Ft[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float Fw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};          // second right-hand side (VJP): lam_v
        float IAc[36];
#pragma unroll
        for (int e = 0; e < 36; e++) IAc[e] = 0.f;
        for (int i = N - 1; i >= 0; i--) {
            const int k = wt_S[i];
            float *sj = joint_ptr(s, i);
            const Xf X = ldX(sj + 32 * wE);
            float c = 0.f;
            if (C0) {
                float f[6];
                ld6(sj + 32 * wMF, f);
#pragma unroll
                for (int r = 0; r < 6; r++) f[r] += fc[r];
                c = pick(f, k) + wt_damping[i] * sj[32 * wQD];
                xtforce(X, f, fc);
            }
            if (ABI) {
                float IA[36], U[6];
#pragma unroll
                for (int e = 0; e < 36; e++) IA[e] = wt_I[36 * i + e] + IAc[e];
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    float u = IA[6 * r];
#pragma unroll
                    for (int t = 1; t < 6; t++) u = (k == t) ? IA[6 * r + t] : u;
                    U[r] = u;
                }
                const float Dinv = 1.0f / pick(U, k);
                st6(sj + 32 * wU, U);
                sj[32 * wDINV] = Dinv;
                if (SOLVE) {
                    const float tau = __ldg(row + 2 * N + i) - c;
                    const float y = Dinv * (tau - pick(Ft, k));
                    sj[32 * wY] = y;
                    float t[6];
#pragma unroll
                    for (int r = 0; r < 6; r++) t[r] = fmaf(U[r], y, Ft[r]);
                    xtforce(X, t, Ft);
                    if (VJP) {
                        const float y2 = Dinv * (__ldg(d_qdd + st * 2 * N + N + i) - pick(Fw, k));
                        sj[32 * wMXA] = y2;                    // free until pass 4
#pragma unroll
                        for (int r = 0; r < 6; r++) t[r] = fmaf(U[r], y2, Fw[r]);
                        xtforce(X, t, Fw);
                    }
                }
                if (i > 0) abi_to_parent(X, IA, U, Dinv, IAc);
            }
        }
    }
    // pass 3 (forward): second half of the solve
    if (SOLVE) {
        float ap[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, aw[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < N; i++) {
            const int k = wt_S[i];
            float *sj = joint_ptr(s, i);
            float U[6], t[6];
            ld6(sj + 32 * wU, U);
            if (i > 0) {
                const Xf X = ldX(sj + 32 * wE);
                xmotion(X, ap, t);
#pragma unroll
                for (int r = 0; r < 6; r++) ap[r] = t[r];
                if (VJP) {
                    xmotion(X, aw, t);
#pragma unroll
                    for (int r = 0; r < 6; r++) aw[r] = t[r];
                }
            }
            const float Dinv = sj[32 * wDINV];
            const float qdd = sj[32 * wY] - Dinv * dot6(U, ap);
            add_at(ap, k, qdd);
            sj[32 * wQDD] = qdd;
            if (QDD_OUT && valid) d_qdd_out[st * N + i] = qdd;
            if (VJP) {
                const float w = sj[32 * wMXA] - Dinv * dot6(U, aw);
                add_at(aw, k, w);
                sj[32 * wY] = w;                                // the column kernel reads w = Minv lam_v here
                if (valid) d_qdd_out[st * CONS_WORDS + 4 * N + i] = dt * w;
            }
            if ((VJP || LIN) && valid) {                        // x+ = [q + dt qd ; qd + dt qdd]
                const float q = __ldg(row + i), qd = sj[32 * wQD];
                d_qdd_out[st * CONS_WORDS + i] = fmaf(dt, qd, q);
                d_qdd_out[st * CONS_WORDS + N + i] = fmaf(dt, qdd, qd);
            }
        }
    }
    // pass 4: RNEA at qdd -> mxS(X a_parent) and mxS(f)
    if (GRAD) {
        float a[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < N; i++) {
            const int k = wt_S[i];
            float *sj = joint_ptr(s, i);
            const Xf X = ldX(sj + 32 * wE);
            float v[6], iv[6], Xa[6], t[6], f[6];
            ld6(sj + 32 * wV, v);
            ld6(sj + 32 * wIV, iv);
            const float qd = sj[32 * wQD];
            const float qdd = QDD_IN ? __ldg(d_qdd + st * N + i) : (SOLVE ? sj[32 * wQDD] : 0.f);
            if (i == 0) {
                float e[6] = {0.f, 0.f, 0.f, 0.f, 0.f, gravity};
                xmotion(X, e, Xa);
            } else {
                xmotion(X, a, Xa);
            }
            mxS(k, Xa, t);
            st6(sj + 32 * wMXA, t);
#pragma unroll
            for (int r = 0; r < 6; r++) a[r] = Xa[r];
            add_at(a, k, qdd);
            if (i > 0) {
                mxS(k, v, t);
#pragma unroll
                for (int r = 0; r < 6; r++) a[r] = fmaf(t[r], qd, a[r]);
            }
            imul(i, a, f);
            crossf(v, iv, t);
#pragma unroll
            for (int r = 0; r < 6; r++) f[r] += t[r];
            st6(sj + 32 * wMF, f);
        }
        float fc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int i = N - 1; i >= 0; i--) {
            const int k = wt_S[i];
            float *sj = joint_ptr(s, i);
            float f[6], t[6];
            ld6(sj + 32 * wMF, f);
#pragma unroll
            for (int r = 0; r < 6; r++) f[r] += fc[r];
            mxS(k, f, t);
            st6(sj + 32 * wMF, t);
            if (i > 0) {
                const Xf X = ldX(sj + 32 * wE);
                xtforce(X, f, fc);
            }
        }
    }
}

// per-warp [joint][lane] tile -> LEN contiguous words per state at g_tile + state * out_words + off
__device__ __forceinline__ void store_rows(float *__restrict__ g_tile, long long out_words, int off, const float *sa,
                                           int len, int cnt, int lane) {
    __syncwarp();
    for (int st = 0; st < cnt; st++)
        for (int i = lane; i < len; i += 32) g_tile[(long long)st * out_words + off + i] = sa[i * PITCH + st];
    __syncwarp();
}

// ---- Minv columns: warp = (32 states, column j) -----------------------------------------------------
// replaces direct_minv_inner (algorithms/_direct_minv.py:23-382) for one column: F column backward from joint j
// to the root, then forward from the root to j; rows below the diagonal are written as zeros.
// LIN = false: d_Minv gets the upper-triangular, column-major Minv of the reference contract.
// LIN = true (fused linearisation consumer): the block B2 = dt Minv, full symmetric, inside rows of 2n + 3n^2 words.
template <int WARPS, bool LIN>
__global__ void __launch_bounds__(32 * WARPS)
minv_columns_kernel(float *__restrict__ d_Minv, const float *__restrict__ scratch, int num_states, int ntiles, float dt) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sa = smem + warp * (N * PITCH);
    // sub-chunk major, then column-major (long columns = large j first), WARPS consecutive tiles per CTA
    long long task = (long long)blockIdx.x * WARPS + warp;
    if (task >= (long long)N * ntiles) return;
    int sub = 0, sub_tiles = min(kSubTiles, ntiles);
    while (task >= (long long)N * sub_tiles) {
        task -= (long long)N * sub_tiles;
        sub++;
        sub_tiles = min(kSubTiles, ntiles - sub * kSubTiles);
    }
    const int j = N - 1 - (int)(task / sub_tiles), tile = sub * kSubTiles + (int)(task % sub_tiles);
    const int cnt = min(32, num_states - tile * 32);
    const float *s = scratch + (size_t)tile * (N * W * 32) + lane;
    float F[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = j; i >= 0; i--) {
        const int k = wt_S[i];
        const float *sj = s + (size_t)i * (W * 32);
        const float Dinv = sj[32 * wDINV];
        const float m = (i == j ? Dinv : 0.f) - Dinv * pick(F, k);
        sa[i * PITCH + lane] = m;
        if (i > 0) {
            float U[6], t[6];
            ld6(sj + 32 * wU, U);
#pragma unroll
            for (int r = 0; r < 6; r++) t[r] = fmaf(U[r], m, F[r]);
            const Xf X = ldX(sj + 32 * wE);
            xtforce(X, t, F);
        }
    }
    float G[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i <= j; i++) {
        const int k = wt_S[i];
        const float *sj = s + (size_t)i * (W * 32);
        float m = sa[i * PITCH + lane];
        if (i > 0) {
            float U[6], t[6];
            ld6(sj + 32 * wU, U);
            const Xf X = ldX(sj + 32 * wE);
            xmotion(X, G, t);
#pragma unroll
            for (int r = 0; r < 6; r++) G[r] = t[r];
            m -= sj[32 * wDINV] * dot6(U, G);
            sa[i * PITCH + lane] = m;
        }
        add_at(G, k, m);
    }
    if (!LIN) {
        for (int i = j + 1; i < N; i++) sa[i * PITCH + lane] = 0.f;
        store_rows(d_Minv + (long long)tile * 32 * N * N, (long long)N * N, j * N, sa, N, cnt, lane);
    } else {
        constexpr long long OUTW = 2 * N + 3 * N * N;
        constexpr int B2 = 2 * N + 2 * N * N;
        float *o = d_Minv + ((long long)tile * 32 + lane) * OUTW + B2;
        if (lane < cnt)
            for (int i = 0; i < j; i++) o[i * N + j] = dt * sa[i * PITCH + lane];   // mirror: row j of the columns i < j
        for (int i = 0; i <= j; i++) sa[i * PITCH + lane] *= dt;
        store_rows(d_Minv + (long long)tile * 32 * OUTW, OUTW, B2 + j * N, sa, j + 1, cnt, lane);
    }
}

// everything the gradient recursion reads per joint and state: 43 scratch words
struct JointData {
    Xf X, X0;
    float v[6], iv[6], s0[6], qd;
};
__device__ __forceinline__ JointData load_joint(const float *sj) {
    JointData d;
    d.X = ldX(sj + 32 * wE);
    ld6(sj + 32 * wV, d.v);
    ld6(sj + 32 * wIV, d.iv);
    d.qd = sj[32 * wQD];
    ld6(sj + 32 * wS0, d.s0);
    d.X0 = ldX(sj + 32 * wE0);
    return d;
}

// ---- gradient columns: warp = (32 states, du-column) ---------------------------------------------------
// MODE 0: dc_du column; 1: df_du column = -Minv dc_du column through the articulated-body solve;
//      2: fused VJP consumer: (A^T lam)[j] or (A^T lam)[n + j] = lam terms - dt dc_du[:, col] . w, w = Minv lam_v from stage A;
//      3: fused linearisation consumer: column of A21 = dt dqdd/dq or A22 = I + dt dqdd/dqd.
// Measured on the 64-link chain (profiles/r2_lps_column_kernel_variants.jsonl, FD gradient of 16 384 states): occupancy
// beats hand-made software pipelining - 3 CTAs per SM at 80 registers with the loads inside the iteration 2 614 us;
// joint i + 1 requested before joint i is computed: 2 852 us at 2 CTAs / 128 registers (4 556 us when squeezed into 80),
// unrolled by two 2 792 us, at 1 CTA / 158 registers 3 866 us.
#ifndef GRID_LPS_PIPELINE
#define GRID_LPS_PIPELINE 0            // 0: load inside the iteration, 1: request joint i + 1 before computing joint i, 2: + unroll by 2
#endif
#ifndef GRID_LPS_MINB
#define GRID_LPS_MINB 3                // resident CTAs per SM the column kernel is compiled for (register cap 65536 / (256 * MINB))
#endif
template <int WARPS, int MODE>
__global__ void __launch_bounds__(32 * WARPS, GRID_LPS_MINB)
grad_columns_kernel(float *__restrict__ d_out, const float *__restrict__ scratch, const float *__restrict__ d_lam,
                    int num_states, int ntiles, float dt) {
    constexpr bool SOLVE = MODE == 1 || MODE == 3;
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sa = smem + warp * (N * PITCH);
    // task order: joint-major, the two sides of a joint adjacent, long columns (small j) first; the WARPS warps
    // of a CTA take adjacent columns of ONE tile, so they read the same scratch lines at about the same time
    constexpr int cgroups = (2 * N + WARPS - 1) / WARPS;
    int blk = blockIdx.x, sub = 0, sub_tiles = min(kSubTiles, ntiles);
    while (blk >= cgroups * sub_tiles) {            // sub-chunks of kSubTiles tiles, the last one may be shorter
        blk -= cgroups * sub_tiles;
        sub++;
        sub_tiles = min(kSubTiles, ntiles - sub * kSubTiles);
    }
    const int tile = sub * kSubTiles + blk % sub_tiles, cg = blk / sub_tiles;
    const int cc = cg * WARPS + warp;
    if (cc >= 2 * N) return;
    const int j = cc >> 1, side = cc & 1;
    const int cnt = min(32, num_states - tile * 32);
    const float *s = scratch + (size_t)tile * (N * W * 32) + lane;

    // (GRID_LPS_PIPELINE > 0 requests the 43 scratch words of joint i + 1 before joint i is computed; the kernel is
    // bound by exposed L1/L2 latency - long_scoreboard 4.4 stalls per issue, profiles/r2_ncu_lps_fdgrad_chain64_N4096.md -
    // but the registers that costs lose more occupancy than the prefetch gains, see above.)
    float dv[6], da[6], P[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    JointData cur = load_joint(s + (size_t)j * (W * 32));
    {   // joint j starts the column
        const int k = wt_S[j];
        const float *sj = s + (size_t)j * (W * 32);
        float t[6];
        if (side == 0) {
            mxS(k, cur.v, dv);                              // == mxS(X v_parent)
            float mxa[6];
            ld6(sj + 32 * wMXA, mxa);
            mxS(k, dv, t);
#pragma unroll
            for (int r = 0; r < 6; r++) da[r] = fmaf(t[r], cur.qd, mxa[r]);
        } else {
#pragma unroll
            for (int r = 0; r < 6; r++) dv[r] = 0.f;
            add_at(dv, k, 1.0f);
            mxS(k, cur.v, da);
        }
    }
#if GRID_LPS_PIPELINE == 2
#pragma unroll 2
#else
#pragma unroll 1
#endif
    for (int i = j; i < N; i++) {
        const int k = wt_S[i];
#if GRID_LPS_PIPELINE == 0
        if (i > j) cur = load_joint(s + (size_t)i * (W * 32));
        JointData nxt = cur;
#else
        JointData nxt = cur;
        if (i + 1 < N) nxt = load_joint(s + (size_t)(i + 1) * (W * 32));
#endif
        float t[6];
        if (i > j) {
            float n6[6];
            xmotion(cur.X, dv, n6);
#pragma unroll
            for (int r = 0; r < 6; r++) dv[r] = n6[r];
            xmotion(cur.X, da, n6);
            mxS(k, dv, t);
#pragma unroll
            for (int r = 0; r < 6; r++) da[r] = fmaf(t[r], cur.qd, n6[r]);
        }
        // df = I da + dv x* (I v) + v x* (I dv)
        float df[6], idv[6];
        imul(i, da, df);
        crossf(dv, cur.iv, t);
#pragma unroll
        for (int r = 0; r < 6; r++) df[r] += t[r];
        imul(i, dv, idv);
        crossf(cur.v, idv, t);
#pragma unroll
        for (int r = 0; r < 6; r++) df[r] += t[r];
        // base-frame bookkeeping: keep s0_i . (forces of the joints before i), add this joint's force
        sa[i * PITCH + lane] = dot6(cur.s0, P);
        xtforce(cur.X0, df, t);
#pragma unroll
        for (int r = 0; r < 6; r++) P[r] += t[r];
        cur = nxt;
    }
    // rows: i >= j: s0_i . (Total - prefix_i); i < j: s0_i . (Total - [dq] mxS(f_j) in the base frame)
    float T2[6];
#pragma unroll
    for (int r = 0; r < 6; r++) T2[r] = P[r];
    if (side == 0) {
        const float *sj = s + (size_t)j * (W * 32);
        float mf[6], t[6];
        ld6(sj + 32 * wMF, mf);
        const Xf X0 = ldX(sj + 32 * wE0);
        xtforce(X0, mf, t);
#pragma unroll
        for (int r = 0; r < 6; r++) T2[r] -= t[r];
    }
#pragma unroll 4
    for (int i = 0; i < N; i++) {
        const float *sj = s + (size_t)i * (W * 32);
        float s0[6];
        ld6(sj + 32 * wS0, s0);
        float dc = i < j ? dot6(s0, T2) : dot6(s0, P) - sa[i * PITCH + lane];
        if (side == 1 && i == j) dc += wt_damping[i];
        sa[i * PITCH + lane] = dc;
    }
    if (SOLVE) {
        // x = Minv dc through the articulated-body recursions, then the column is -x
        float F[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int i = N - 1; i >= 0; i--) {
            const int k = wt_S[i];
            const float *sj = s + (size_t)i * (W * 32);
            const float y = sj[32 * wDINV] * (sa[i * PITCH + lane] - pick(F, k));
            sa[i * PITCH + lane] = y;
            if (i > 0) {
                float U[6], t[6];
                ld6(sj + 32 * wU, U);
#pragma unroll
                for (int r = 0; r < 6; r++) t[r] = fmaf(U[r], y, F[r]);
                const Xf X = ldX(sj + 32 * wE);
                xtforce(X, t, F);
            }
        }
        float ap[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int i = 0; i < N; i++) {
            const int k = wt_S[i];
            const float *sj = s + (size_t)i * (W * 32);
            float x = sa[i * PITCH + lane];
            if (i > 0) {
                float U[6], t[6];
                ld6(sj + 32 * wU, U);
                const Xf X = ldX(sj + 32 * wE);
                xmotion(X, ap, t);
#pragma unroll
                for (int r = 0; r < 6; r++) ap[r] = t[r];
                x -= sj[32 * wDINV] * dot6(U, ap);
            }
            add_at(ap, k, x);
            sa[i * PITCH + lane] = MODE == 3 ? fmaf(-dt, x, (side == 1 && i == j) ? 1.0f : 0.0f) : -x;
        }
    }
    if (MODE == 2) {
        float g = 0.f;
        for (int i = 0; i < N; i++) g = fmaf(sa[i * PITCH + lane], s[(size_t)i * (W * 32) + 32 * wY], g);
        if (lane < cnt) {
            const long long st = (long long)tile * 32 + lane;
            const float lq = __ldg(d_lam + st * 2 * N + j);
            float *o = d_out + st * 5 * N;
            if (side == 0) o[2 * N + j] = fmaf(-dt, g, lq);
            else o[3 * N + j] = __ldg(d_lam + st * 2 * N + N + j) + dt * (lq - g);
        }
    } else if (MODE == 3) {
        constexpr long long OUTW = 2 * N + 3 * N * N;
        store_rows(d_out + (long long)tile * 32 * OUTW, OUTW, 2 * N + side * N * N + j * N, sa, N, cnt, lane);
    } else {
        store_rows(d_out + (long long)tile * 32 * 2 * N * N, (long long)2 * N * N, side * N * N + j * N, sa, N, cnt, lane);
    }
}

// ---- gradient columns, CTA-cooperative: joint blocks streamed into shared memory by the bulk-copy engine (TMA) ----
// The column kernel above waits for L1/L2 (long_scoreboard 4.4 stalls per issue): every warp loads the same 43 words
// per joint for itself.  Here the WARPS warps of a CTA (one tile, adjacent columns) share ONE copy of each joint's
// 8 KB scratch block ([word][lane], contiguous), which a single thread requests STAGES - 1 joints ahead with
// cp.async.bulk (global -> shared, completion on an mbarrier); the warps then read shared memory with a fixed latency.
// One __syncthreads() per joint step hands the consumed buffer back to the producer.  The schedule of a CTA is one
// linear sequence of joint blocks: main recursion j0..N-1, finishing pass 0..N-1, and for the FD gradient the two
// passes of the articulated-body solve N-1..0, 0..N-1.  Same arithmetic, same order, same results as the kernel above.
#ifndef GRID_LPS_BULK
#define GRID_LPS_BULK 0                // 1: the column kernels of ALG >= 2 use this variant
#endif
#ifndef GRID_LPS_BULK_STAGES
#define GRID_LPS_BULK_STAGES 3
#endif
__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

template <int WARPS, int MODE>
__global__ void __launch_bounds__(32 * WARPS, 2)
grad_columns_bulk_kernel(float *__restrict__ d_out, const float *__restrict__ scratch, const float *__restrict__ d_lam,
                         int num_states, int ntiles, float dt) {
    constexpr bool SOLVE = MODE == 1 || MODE == 3;
    constexpr int STAGES = GRID_LPS_BULK_STAGES, JB = W * 32;          // floats per joint block (8 KB)
    constexpr int cgroups = (2 * N + WARPS - 1) / WARPS;
    extern __shared__ __align__(128) float smem[];
    __shared__ __align__(8) unsigned long long full[STAGES];
    float *buf = smem;                                                  // STAGES joint blocks
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *sa = smem + STAGES * JB + warp * (N * PITCH);
    int blk = blockIdx.x, sub = 0, sub_tiles = min(kSubTiles, ntiles);
    while (blk >= cgroups * sub_tiles) {
        blk -= cgroups * sub_tiles;
        sub++;
        sub_tiles = min(kSubTiles, ntiles - sub * kSubTiles);
    }
    const int tile = sub * kSubTiles + blk % sub_tiles, cg = blk / sub_tiles;
    const int cc = cg * WARPS + warp;
    const bool active = cc < 2 * N;                                     // a CTA past the last column pair idles its spare warps
    const int j = active ? cc >> 1 : N - 1, side = cc & 1;
    const int j0 = (cg * WARPS) >> 1;                                   // first joint any warp of this CTA starts at
    const int cnt = min(32, num_states - tile * 32);
    const float *tile_base = scratch + (size_t)tile * (N * W * 32);
    // schedule: [0, nA) main j0..N-1 | [nA, nA+N) finishing 0..N-1 | solve backward N-1..0 | solve forward 0..N-1
    const int nA = N - j0, T = nA + N + (SOLVE ? 2 * N : 0);
    auto joint_of = [&](int st) { return st < nA ? j0 + st : st < nA + N ? st - nA : st < nA + 2 * N ? nA + 2 * N - 1 - st : st - nA - 2 * N; };
    if (threadIdx.x == 0) {
        for (int b = 0; b < STAGES; b++) mbar_init(&full[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int st = 0; st < STAGES - 1 && st < T; st++) {
            mbar_expect_tx(&full[st], JB * 4);
            bulk_g2s(buf + st * JB, tile_base + (size_t)joint_of(st) * JB, JB * 4, &full[st]);
        }
    float dv[6], da[6], P[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, T2[6], F[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, g = 0.f;
    for (int st = 0; st < T; st++) {
        __syncthreads();                                                // everyone is done with the buffer of step st - 1
        if (threadIdx.x == 0 && st + STAGES - 1 < T) {
            const int b = (st + STAGES - 1) % STAGES;
            mbar_expect_tx(&full[b], JB * 4);
            bulk_g2s(buf + b * JB, tile_base + (size_t)joint_of(st + STAGES - 1) * JB, JB * 4, &full[b]);
        }
        mbar_wait(&full[st % STAGES], (st / STAGES) & 1);
        const float *sj = buf + (st % STAGES) * JB + lane;              // this joint's block, [word][lane]
        const int i = joint_of(st), k = wt_S[i];
        if (!active) continue;
        if (st < nA) {                                                  // ---- main recursion
            if (i < j) continue;
            float v[6], iv[6], t[6];
            ld6(sj + 32 * wV, v);
            ld6(sj + 32 * wIV, iv);
            const float qd = sj[32 * wQD];
            if (i == j) {
                if (side == 0) {
                    mxS(k, v, dv);
                    float mxa[6], mf[6];
                    ld6(sj + 32 * wMXA, mxa);
                    mxS(k, dv, t);
#pragma unroll
                    for (int r = 0; r < 6; r++) da[r] = fmaf(t[r], qd, mxa[r]);
                    ld6(sj + 32 * wMF, mf);                              // leaves the column at joint j (rows i < j)
                    const Xf X0 = ldX(sj + 32 * wE0);
                    xtforce(X0, mf, T2);
                } else {
#pragma unroll
                    for (int r = 0; r < 6; r++) { dv[r] = 0.f; T2[r] = 0.f; }
                    add_at(dv, k, 1.0f);
                    mxS(k, v, da);
                }
            } else {
                const Xf X = ldX(sj + 32 * wE);
                float n6[6];
                xmotion(X, dv, n6);
#pragma unroll
                for (int r = 0; r < 6; r++) dv[r] = n6[r];
                xmotion(X, da, n6);
                mxS(k, dv, t);
#pragma unroll
                for (int r = 0; r < 6; r++) da[r] = fmaf(t[r], qd, n6[r]);
            }
            float df[6], idv[6], s0[6];
            imul(i, da, df);
            crossf(dv, iv, t);
#pragma unroll
            for (int r = 0; r < 6; r++) df[r] += t[r];
            imul(i, dv, idv);
            crossf(v, idv, t);
#pragma unroll
            for (int r = 0; r < 6; r++) df[r] += t[r];
            ld6(sj + 32 * wS0, s0);
            sa[i * PITCH + lane] = dot6(s0, P);
            const Xf X0 = ldX(sj + 32 * wE0);
            xtforce(X0, df, t);
#pragma unroll
            for (int r = 0; r < 6; r++) P[r] += t[r];
            if (i == N - 1) {                                            // Total is known: T2 = Total - [dq] mxS(f_j) in the base frame
#pragma unroll
                for (int r = 0; r < 6; r++) T2[r] = P[r] - T2[r];
            }
        } else if (st < nA + N) {                                       // ---- finishing pass: rows of the column
            float s0[6];
            ld6(sj + 32 * wS0, s0);
            float dc = i < j ? dot6(s0, T2) : dot6(s0, P) - sa[i * PITCH + lane];
            if (side == 1 && i == j) dc += wt_damping[i];
            sa[i * PITCH + lane] = dc;
            if (MODE == 2) g = fmaf(dc, sj[32 * wY], g);
        } else if (st < nA + 2 * N) {                                   // ---- solve, backward
            const float y = sj[32 * wDINV] * (sa[i * PITCH + lane] - pick(F, k));
            sa[i * PITCH + lane] = y;
            if (i > 0) {
                float U[6], t[6];
                ld6(sj + 32 * wU, U);
#pragma unroll
                for (int r = 0; r < 6; r++) t[r] = fmaf(U[r], y, F[r]);
                const Xf X = ldX(sj + 32 * wE);
                xtforce(X, t, F);
            } else {
#pragma unroll
                for (int r = 0; r < 6; r++) F[r] = 0.f;                  // F becomes the forward accumulator
            }
        } else {                                                        // ---- solve, forward
            float x = sa[i * PITCH + lane];
            if (i > 0) {
                float U[6], t[6];
                ld6(sj + 32 * wU, U);
                const Xf X = ldX(sj + 32 * wE);
                xmotion(X, F, t);
#pragma unroll
                for (int r = 0; r < 6; r++) F[r] = t[r];
                x -= sj[32 * wDINV] * dot6(U, F);
            }
            add_at(F, k, x);
            sa[i * PITCH + lane] = MODE == 3 ? fmaf(-dt, x, (side == 1 && i == j) ? 1.0f : 0.0f) : -x;
        }
    }
    if (!active) return;
    if (MODE == 2) {
        if (lane < cnt) {
            const long long stt = (long long)tile * 32 + lane;
            const float lq = __ldg(d_lam + stt * 2 * N + j);
            float *o = d_out + stt * 5 * N;
            if (side == 0) o[2 * N + j] = fmaf(-dt, g, lq);
            else o[3 * N + j] = __ldg(d_lam + stt * 2 * N + N + j) + dt * (lq - g);
        }
    } else if (MODE == 3) {
        constexpr long long OUTW = 2 * N + 3 * N * N;
        store_rows(d_out + (long long)tile * 32 * OUTW, OUTW, 2 * N + side * N * N + j * N, sa, N, cnt, lane);
    } else {
        store_rows(d_out + (long long)tile * 32 * 2 * N * N, (long long)2 * N * N, side * N * N + j * N, sa, N, cnt, lane);
    }
}

// ---- df_du = -Minv dc_du with a caller-supplied Minv (USE_QDD_MINV_FLAG overload) -------------------------------
// The reference's overload takes qdd and Minv from the caller (algorithms/_forward_dynamics_gradient.py:22-25,
// 202-220) and must use THAT Minv, so the O(n) solve of the column kernel does not apply: this is the one place
// where the path has a real batched product, (n x n)(n x 2n) per state, and the one place tensor cores are used.
// CTA = one state; Minv (column-major, upper triangle read symmetrically) and the dc_du block the column kernel left
// in d_out are staged in shared memory with a padded pitch; N % 16 == 0: mma.sync.m16n8k8 TF32 with the 3xTF32 split
// (FP32-class accuracy: 1.4e-6 relative at n = 64, 2x the FFMA rate with operands on chip,
// profiles/r2_micro_tc_minv_gemm.jsonl); other N: FP32 FFMA, one thread per output column.
template <int NN>
__global__ void __launch_bounds__(NN % 16 == 0 ? 32 * (NN / 16) : 64)
minv_product_kernel(float *__restrict__ d_out, const float *__restrict__ d_Minv, int num_states) {
    constexpr int P = NN + 4;
    extern __shared__ float smem[];
    float *sA = smem, *sB = smem + NN * P;                       // A(row, k) = Minv, B(k, col) = dc at col * P + k
    const int nthr = blockDim.x;
    for (long long st = blockIdx.x; st < num_states; st += gridDim.x) {
        const float *M = d_Minv + st * NN * NN;
        float *o = d_out + st * 2 * NN * NN;
        for (int e = threadIdx.x; e < NN * NN; e += nthr) {
            const int c = e / NN, r = e - c * NN;                // column-major upper triangle -> full symmetric
            if (r <= c) {
                const float m = __ldg(M + e);
                sA[r * P + c] = m;
                sA[c * P + r] = m;
            }
        }
        for (int e = threadIdx.x; e < 2 * NN * NN; e += nthr) sB[(e / NN) * P + e % NN] = o[e];
        __syncthreads();
        if constexpr (NN % 16 == 0) {
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
            float acc[2 * NN / 8][4];
#pragma unroll
            for (int j = 0; j < 2 * NN / 8; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
            for (int k0 = 0; k0 < NN; k0 += 8) {
                const float af[4] = {sA[(16 * warp + g) * P + k0 + t], sA[(16 * warp + g + 8) * P + k0 + t],
                                     sA[(16 * warp + g) * P + k0 + t + 4], sA[(16 * warp + g + 8) * P + k0 + t + 4]};
                unsigned ah[4], al[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    ah[i] = wps::to_tf32(af[i]);
                    al[i] = wps::to_tf32(af[i] - __uint_as_float(ah[i]));
                }
#pragma unroll
                for (int j = 0; j < 2 * NN / 8; j++) {
                    const float bf[2] = {sB[(8 * j + g) * P + k0 + t], sB[(8 * j + g) * P + k0 + t + 4]};
                    const unsigned bh[2] = {wps::to_tf32(bf[0]), wps::to_tf32(bf[1])};
                    const unsigned bl[2] = {wps::to_tf32(bf[0] - __uint_as_float(bh[0])),
                                            wps::to_tf32(bf[1] - __uint_as_float(bh[1]))};
                    wps::mma_m16n8k8_tf32(acc[j], al, bh);
                    wps::mma_m16n8k8_tf32(acc[j], ah, bl);
                    wps::mma_m16n8k8_tf32(acc[j], ah, bh);
                }
            }
#pragma unroll
            for (int j = 0; j < 2 * NN / 8; j++) {
                const int c0 = 8 * j + 2 * t, r0 = 16 * warp + g;
                o[c0 * NN + r0] = -acc[j][0];
                o[(c0 + 1) * NN + r0] = -acc[j][1];
                o[c0 * NN + r0 + 8] = -acc[j][2];
                o[(c0 + 1) * NN + r0 + 8] = -acc[j][3];
            }
        } else {
            for (int col = threadIdx.x; col < 2 * NN; col += nthr)
                for (int r = 0; r < NN; r++) {
                    float a = 0.f;
                    for (int k = 0; k < NN; k++) a = fmaf(sA[r * P + k], sB[col * P + k], a);
                    o[col * NN + r] = -a;
                }
        }
        __syncthreads();
    }
}

// ---- launchers -------------------------------------------------------------------------------------------

// shared-memory opt-in of a column kernel on the current device, once per (kernel, device).  Keyed by the
// kernel's ADDRESS: the two gradient kernels have the same function type, a per-type static would be shared.
static cudaError_t opt_in_smem(const void *kern, size_t bytes) {
    constexpr int kMaxKernels = 8;
    static const void *seen[kMaxDevices][kMaxKernels];
    int dev = 0;
    if (cudaError_t e = current_device(dev)) return e;
    for (int i = 0; i < kMaxKernels; i++) {
        if (seen[dev][i] == kern) return cudaSuccess;
        if (seen[dev][i] == nullptr) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
            if (e == cudaSuccess) seen[dev][i] = kern;       // benign race: idempotent
            return e;
        }
    }
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// ALG: 0 = Minv, 1 = FD, 2 = ID gradient (HAS_QDD: qdd given), 3 = FD gradient,
//      4 = fused VJP consumer (d_qdd = lambda, 2n per state; d_out rows of 5n), 5 = fused linearisation consumer
//      ALG 3 with HAS_QDD: the USE_QDD_MINV_FLAG overload - dc_du columns at the given qdd, then the product with the
//      caller's Minv (d_Minv) on the tensor cores (minv_product_kernel)
template <int ALG, bool HAS_QDD>
cudaError_t lps_launch(float *d_out, const float *d_in, int stride, const float *d_qdd, int num_states, float gravity,
                       cudaStream_t stream, float dt = 0.f, const float *d_Minv = nullptr) {
    if (num_states <= 0) return cudaSuccess;
    g_calls.fetch_add(1);
    constexpr bool PRE = ALG == 3 && HAS_QDD;
    constexpr int FLAGS = ALG == 0 ? 2 : ALG == 1 ? (1 | 2 | 4 | 32) : (ALG == 2 || PRE) ? (16 | (HAS_QDD ? 8 : 0))
                        : ALG == 3 ? (1 | 2 | 4 | 16) : ALG == 4 ? (1 | 2 | 4 | 16 | 64) : (1 | 2 | 4 | 16 | 128);
    constexpr long long OUTW = ALG == 0 ? N * N : ALG == 1 ? N : ALG == 4 ? 5 * N : ALG == 5 ? 2 * N + 3 * N * N : 2 * N * N;
    constexpr int QW = ALG == 4 ? 2 * N : N;              // words per state behind d_qdd (qdd or lambda)
    constexpr size_t col_smem = sizeof(float) * N * PITCH * kColWarps;
    constexpr int GMODE = PRE ? 0 : ALG == 3 ? 1 : ALG == 4 ? 2 : ALG == 5 ? 3 : 0;
    constexpr size_t prod_smem = sizeof(float) * 3 * N * (N + 4);
    auto prod_kern = minv_product_kernel<N>;
    if (PRE && !d_Minv) return cudaErrorInvalidValue;
    auto minv_kern = minv_columns_kernel<kColWarps, ALG == 5>;
#if GRID_LPS_BULK
    auto grad_kern = grad_columns_bulk_kernel<kColWarps, GMODE>;
    constexpr size_t grad_smem = col_smem + sizeof(float) * GRID_LPS_BULK_STAGES * W * 32;
#else
    auto grad_kern = grad_columns_kernel<kColWarps, GMODE>;
    constexpr size_t grad_smem = col_smem;
#endif
    cudaError_t e = cudaSuccess;
    if (ALG == 0 || ALG == 5) e = opt_in_smem((const void *)minv_kern, col_smem);
    if (e == cudaSuccess && ALG >= 2) e = opt_in_smem((const void *)grad_kern, grad_smem);
    if (e == cudaSuccess && PRE) e = opt_in_smem((const void *)prod_kern, prod_smem);
    if (e != cudaSuccess) return e;
    const int chunk = num_states < kChunkStates ? num_states : kChunkStates;
    const size_t sc_bytes = (size_t)((chunk + 31) / 32) * N * W * 32 * sizeof(float);
    float *scratch = nullptr;
    keep_pool_memory();
    e = cudaMallocAsync((void **)&scratch, sc_bytes, stream);
    if (e != cudaSuccess) return e;
    for (int first = 0; first < num_states && e == cudaSuccess; first += chunk) {
        const int n = num_states - first < chunk ? num_states - first : chunk;
        const int ntiles = (n + 31) / 32;
        const float *in = d_in + (long long)first * stride;
        const float *qdd = d_qdd ? d_qdd + (long long)first * QW : nullptr;
        float *out = d_out + (long long)first * OUTW;
        stage_a_kernel<FLAGS><<<(ntiles + 3) / 4, 128, 0, stream>>>(in, stride, qdd, scratch,
                                                                    (ALG == 1 || ALG >= 4) ? out : nullptr, n, gravity, dt);
        g_kernel_launches.fetch_add(1);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        if (ALG == 0 || ALG == 5) {
            const long long tasks = (long long)N * ntiles;
            minv_kern<<<(unsigned)((tasks + kColWarps - 1) / kColWarps), 32 * kColWarps, col_smem, stream>>>(out, scratch, n, ntiles, dt);
            g_kernel_launches.fetch_add(1);
            if ((e = cudaGetLastError()) != cudaSuccess) break;
        }
        if (ALG >= 2) {
            const long long blocks = (long long)((2 * N + kColWarps - 1) / kColWarps) * ntiles;
            grad_kern<<<(unsigned)blocks, 32 * kColWarps, grad_smem, stream>>>(out, scratch, qdd, n, ntiles, dt);
            g_kernel_launches.fetch_add(1);
            if (PRE && (e = cudaGetLastError()) == cudaSuccess) {
                int dev = 0, sms = 148;
                if (current_device(dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
                const int pb = n < 8 * sms ? n : 8 * sms;
                prod_kern<<<pb, N % 16 == 0 ? 32 * (N / 16) : 64, prod_smem, stream>>>(out, d_Minv + (long long)first * N * N, n);
                g_kernel_launches.fetch_add(1);
            }
        }
        e = cudaGetLastError();
    }
    cudaError_t e2 = cudaFreeAsync(scratch, stream);
    return e != cudaSuccess ? e : e2;
}

}}  // namespace GRID_NS::lps
