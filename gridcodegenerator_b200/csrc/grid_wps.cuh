// grid_wps.cuh - "wide" kernels: one CTA per state, lanes = columns.
//
// Used where one thread per state is not viable: robots whose traced program is too
// large (Atlas, chain-64: Minv / FD / ID-gradient / FD-gradient) and, for small robots,
// small batches where latency matters (N = 128 knot points).  Replaces the reference's
// block-per-state kernels (algorithms/_direct_minv.py:23-382,
// _inverse_dynamics_gradient.py:27-650, _forward_dynamics_gradient.py:7-57) with a
// different decomposition:
//   * every du-column (dq_j or dqd_j) of the gradient and every column of Minv/F is owned
//     by ONE lane for the whole recursion, so there are no shared-memory atomics and no
//     barrier inside a gradient pass (the reference has 26-197); the Minv
//     passes need one named barrier per joint;
//   * the two RNEA sweeps run on their own warp (6 row-lanes per joint, a tree level at a time),
//     the first one concurrently with the Minv passes;
//   * X_i is kept as (E, r) (12 floats) and applied as two 3x3 products and a cross
//     product (27 FMA) instead of a dense 6x6 (36);
//   * F and the per-column df storage are slot-allocated at generation time from the tree
//     (F: live ancestors only; df: 6*(|anc|+1) per joint), which is what makes the 64-link
//     chain fit: the reference needs 431-519 KB of shared memory per block for its
//     gradients, this layout needs ~160 KB.
// The generated translation unit provides, in GRID_NS::gen, `struct WT` (sizes) and the
// __constant__ tables wt_* before including this file.
#pragma once
#include <cuda_runtime.h>

namespace GRID_NS { namespace wps {

using namespace gen;
constexpr int N = WT::N;
constexpr int NT = WT::NT;

// ---- shared memory layout (floats) ----------------------------------------------------
struct L {
    static constexpr int q = 0, qd = q + N, u = qd + N, qdd = u + N, c = qdd + N;
    static constexpr int Er = ((c + N + 3) / 4) * 4;          // 12 per joint, float4 aligned
    static constexpr int v = Er + 12 * N, a = v + 6 * N, f = a + 6 * N, Xa = f + 6 * N, Iv = Xa + 6 * N;
    static constexpr int Ic = Iv + 6 * N;                      // 36 per joint: link inertias, loaded once per CTA
    static constexpr int Minv = Ic + 36 * N;                   // N*N row-major [row][col], upper
    static constexpr int IA = Minv + N * N;                    // 36 per joint, row-major
    static constexpr int U = IA + 36 * N, W = U + 6 * N, Dinv = W + 6 * N;
    static constexpr int F = Dinv + N;                         // NSLOT * 6 * N, [slot][row][col]
    static constexpr int minv_end = F + WT::NSLOT * 6 * N;
    // the gradient region aliases the Minv scratch (IA, U, W, Dinv, F are dead by then)
    static constexpr int df = IA;                              // 2 * DF_WORDS
    static constexpr int sv = df + 2 * WT::DF_WORDS;           // NSAVE * 12 * 2N  [slot][12][col]
    static constexpr int dc = sv + WT::NSAVE * 12 * 2 * N;     // 2N * N  [col][row] == output layout
    static constexpr int grad_end = dc + 2 * N * N;
    static constexpr int total = ((minv_end > grad_end ? minv_end : grad_end) + 3) / 4 * 4;
};

// ---- per-thread spatial algebra ---------------------------------------------------------
struct Xf { float E[9]; float r[3]; };

__device__ __forceinline__ Xf load_X(const float *s, int i) {
    Xf x;
    const float4 *p = reinterpret_cast<const float4 *>(s + L::Er + 12 * i);
    float4 a = p[0], b = p[1], c = p[2];
    x.E[0] = a.x; x.E[1] = a.y; x.E[2] = a.z; x.E[3] = a.w; x.E[4] = b.x; x.E[5] = b.y; x.E[6] = b.z; x.E[7] = b.w;
    x.E[8] = c.x; x.r[0] = c.y; x.r[1] = c.z; x.r[2] = c.w;
    return x;
}
__device__ __forceinline__ void load6(const float *p, float *o) {
#pragma unroll
    for (int r = 0; r < 6; r++) o[r] = p[r];
}
__device__ __forceinline__ void store6(float *p, const float *o) {
#pragma unroll
    for (int r = 0; r < 6; r++) p[r] = o[r];
}
__device__ __forceinline__ void cross3(const float *a, const float *b, float *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ void mat3(const float *E, const float *x, float *o) {
    o[0] = E[0] * x[0] + E[1] * x[1] + E[2] * x[2];
    o[1] = E[3] * x[0] + E[4] * x[1] + E[5] * x[2];
    o[2] = E[6] * x[0] + E[7] * x[1] + E[8] * x[2];
}
__device__ __forceinline__ void mat3T(const float *E, const float *x, float *o) {
    o[0] = E[0] * x[0] + E[3] * x[1] + E[6] * x[2];
    o[1] = E[1] * x[0] + E[4] * x[1] + E[7] * x[2];
    o[2] = E[2] * x[0] + E[5] * x[1] + E[8] * x[2];
}
// X v (motion vector): [E w ; E (l - r x w)]
__device__ __forceinline__ void xmotion(const Xf &X, const float *v, float *o) {
    float t[3];
    cross3(X.r, v, t);
    t[0] = v[3] - t[0]; t[1] = v[4] - t[1]; t[2] = v[5] - t[2];
    mat3(X.E, v, o);
    mat3(X.E, t, o + 3);
}
// X^T f (force vector): [E^T n + r x (E^T fl) ; E^T fl]
__device__ __forceinline__ void xtforce(const Xf &X, const float *f, float *o) {
    float t[3];
    mat3T(X.E, f + 3, o + 3);
    mat3T(X.E, f, o);
    cross3(X.r, o + 3, t);
    o[0] += t[0]; o[1] += t[1]; o[2] += t[2];
}
// w x e_a, branch-free (a is warp-uniform): component a is 0, (a+1)%3 gets w[(a+2)%3], (a+2)%3 gets -w[(a+1)%3]
__device__ __forceinline__ void wxe(int a, const float *w, float *o) {
    o[0] = a == 1 ? -w[2] : (a == 2 ? w[1] : 0.f);
    o[1] = a == 2 ? -w[0] : (a == 0 ? w[2] : 0.f);
    o[2] = a == 0 ? -w[1] : (a == 1 ? w[0] : 0.f);
}
// (v x) e_k : the reference's mx0..mx5 (helpers/_spatial_algebra_helpers.py:62-147)
__device__ __forceinline__ void mxS(int k, const float *v, float *o) {
    const int a = k < 3 ? k : k - 3;
    float top[3], bot[3];
    wxe(a, v, top);
    wxe(a, v + 3, bot);
    const bool rev = k < 3;
    o[0] = rev ? top[0] : 0.f; o[1] = rev ? top[1] : 0.f; o[2] = rev ? top[2] : 0.f;
    o[3] = rev ? bot[0] : top[0]; o[4] = rev ? bot[1] : top[1]; o[5] = rev ? bot[2] : top[2];
}
__device__ __forceinline__ float pick(const float *v, int k) {
    float r = v[0];
#pragma unroll
    for (int t = 1; t < 6; t++) r = (k == t) ? v[t] : r;
    return r;
}
__device__ __forceinline__ void add_at(float *v, int k, float x) {
#pragma unroll
    for (int t = 0; t < 6; t++) v[t] += (k == t) ? x : 0.f;
}
// v x* f : fx_times_v (helpers/_spatial_algebra_helpers.py:181-256)
__device__ __forceinline__ void crossf(const float *v, const float *f, float *o) {
    float t[3];
    cross3(v, f, o);
    cross3(v + 3, f + 3, t);
    o[0] += t[0]; o[1] += t[1]; o[2] += t[2];
    cross3(v, f + 3, o + 3);
}
// I_i v with the inertia in constant memory (i is warp-uniform)
__device__ __forceinline__ void imul(int i, const float *v, float *o) {
    const float *I = wt_I + 36 * i;
#pragma unroll
    for (int r = 0; r < 6; r++) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 6; c++) acc = fmaf(I[6 * r + c], v[c], acc);
        o[r] = acc;
    }
}

// ---- X(q): E = E_joint(q) E0, r = r0 (+ q * E0[k-3,:] for prismatic) ---------------------
// replaces load_update_XImats_helpers (helpers/_topology_helpers.py:90-182)
__device__ __forceinline__ void update_X(float *s, int i) {
    const int k = wt_S[i];
    const float *E0 = wt_E0 + 9 * i;
    float E[9], r[3] = {wt_r0[3 * i], wt_r0[3 * i + 1], wt_r0[3 * i + 2]};
#pragma unroll
    for (int e = 0; e < 9; e++) E[e] = E0[e];
    const float qi = s[L::q + i];
    if (k < 3) {
        float sn, cs;
        sincosf(qi, &sn, &cs);
        const int a = (k + 1) % 3, b = (k + 2) % 3;
#pragma unroll
        for (int col = 0; col < 3; col++) {
            const float ea = E0[3 * a + col], eb = E0[3 * b + col];
            const float na = cs * ea + sn * eb, nb = cs * eb - sn * ea;
            // rows a and b are warp-divergent indices only across joints; write through selects
#pragma unroll
            for (int row = 0; row < 3; row++) {
                if (row == a) E[3 * row + col] = na;
                if (row == b) E[3 * row + col] = nb;
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < 3; t++) r[t] += qi * E0[3 * (k - 3) + t];
    }
    float *dst = s + L::Er + 12 * i;
#pragma unroll
    for (int e = 0; e < 9; e++) dst[e] = E[e];
    dst[9] = r[0]; dst[10] = r[1]; dst[11] = r[2];
}

// ---- RNEA on one warp: 6 lanes (rows) per joint, up to 5 joints of a tree level at a time ----
// replaces inverse_dynamics_inner / _vaf (algorithms/_inverse_dynamics.py:33-304).  Row r of X is
// rebuilt from (E, r): rows 0-2 = [E_r | 0], rows 3-5 = [r x E_(r-3) | E_(r-3)].
__device__ __forceinline__ void x_row(const Xf &X, int row, float *c) {
    const int a = row < 3 ? row : row - 3;
    const float e[3] = {X.E[3 * a], X.E[3 * a + 1], X.E[3 * a + 2]};
    if (row < 3) {
        c[0] = e[0]; c[1] = e[1]; c[2] = e[2]; c[3] = c[4] = c[5] = 0.f;
    } else {
        cross3(X.r, e, c);
        c[3] = e[0]; c[4] = e[1]; c[5] = e[2];
    }
}
// column `col` of X (for X^T f): X[c][col], c = 0..5
__device__ __forceinline__ void x_col(const Xf &X, int col, float *c) {
    if (col < 3) {
        c[0] = X.E[col]; c[1] = X.E[3 + col]; c[2] = X.E[6 + col];
        float t[3];
#pragma unroll
        for (int a = 0; a < 3; a++) {
            const float e[3] = {X.E[3 * a], X.E[3 * a + 1], X.E[3 * a + 2]};
            cross3(X.r, e, t);
            c[3 + a] = col == 0 ? t[0] : (col == 1 ? t[1] : t[2]);
        }
    } else {
        c[0] = c[1] = c[2] = 0.f;
        c[3] = X.E[col - 3]; c[4] = X.E[3 + col - 3]; c[5] = X.E[6 + col - 3];
    }
}

__device__ void rnea_rows(float *s, int lane, bool use_qdd, float gravity) {
    const int slot = lane / 6, row = lane - 6 * slot;
    const bool lane_ok = lane < 30;
    for (int lvl = 0; lvl < WT::NLEVELS; lvl++) {
        const int first = wt_level_start[lvl], count = wt_level_start[lvl + 1] - first;
        for (int base = 0; base < count; base += 5) {
            const bool act = lane_ok && (base + slot < count);
            const int i = act ? wt_level_joints[first + base + slot] : 0;
            const int par = wt_parent[i], k = wt_S[i];
            const float qdi = s[L::qd + i];
            float vi = 0.f, xa = 0.f;
            if (act) {
                const Xf X = load_X(s, i);
                float c[6];
                x_row(X, row, c);
                if (par < 0) {
                    xa = c[5] * gravity;
                } else {
#pragma unroll
                    for (int t = 0; t < 6; t++) {
                        vi = fmaf(c[t], s[L::v + 6 * par + t], vi);
                        xa = fmaf(c[t], s[L::a + 6 * par + t], xa);
                    }
                }
                if (row == k) vi += qdi;
                s[L::v + 6 * i + row] = vi;
                s[L::Xa + 6 * i + row] = xa;
            }
            __syncwarp();
            float ai = xa;
            if (act) {
                float v6[6], t[6];
                load6(s + L::v + 6 * i, v6);
                if (use_qdd && row == k) ai += s[L::qdd + i];
                if (par >= 0) {
                    mxS(k, v6, t);
                    ai = fmaf(pick(t, row), qdi, ai);
                }
                float iv = 0.f;
#pragma unroll
                for (int t2 = 0; t2 < 6; t2++) iv = fmaf(s[L::Ic + 36 * i + 6 * row + t2], v6[t2], iv);
                s[L::a + 6 * i + row] = ai;
                s[L::Iv + 6 * i + row] = iv;
            }
            __syncwarp();
            if (act) {
                float v6[6], a6[6], iv6[6], t[6];
                load6(s + L::v + 6 * i, v6);
                load6(s + L::a + 6 * i, a6);
                load6(s + L::Iv + 6 * i, iv6);
                float f = 0.f;
#pragma unroll
                for (int t2 = 0; t2 < 6; t2++) f = fmaf(s[L::Ic + 36 * i + 6 * row + t2], a6[t2], f);
                crossf(v6, iv6, t);
                s[L::f + 6 * i + row] = f + pick(t, row);
            }
            __syncwarp();
        }
    }
    for (int lvl = WT::NLEVELS - 1; lvl >= 0; lvl--) {
        const int first = wt_level_start[lvl], count = wt_level_start[lvl + 1] - first;
        for (int base = 0; base < count; base += 5) {
            const bool act = lane_ok && (base + slot < count);
            const int i = act ? wt_level_joints[first + base + slot] : 0;
            const int par = wt_parent[i], k = wt_S[i];
            if (act) {
                float f6[6];
                load6(s + L::f + 6 * i, f6);
                if (row == k) s[L::c + i] = f6[row] + wt_damping[i] * s[L::qd + i];
                if (par >= 0) {
                    const Xf X = load_X(s, i);
                    float c[6], acc = 0.f;
                    x_col(X, row, c);
#pragma unroll
                    for (int t = 0; t < 6; t++) acc = fmaf(c[t], f6[t], acc);
                    atomicAdd(s + L::f + 6 * par + row, acc);      // siblings of one level share a parent
                }
            }
            __syncwarp();
        }
    }
}

// barrier among the threads that run the Minv passes (every warp except the RNEA warp)
__device__ __forceinline__ void minv_bar() {
    asm volatile("bar.sync 1, %0;" ::"n"(NT - 32) : "memory");
}

// ---- Minv: both passes.  Called by all threads with tid < NT-32 ------------------------------
// replaces direct_minv_inner (algorithms/_direct_minv.py:23-382; oracle _test.py:117-226)
__device__ void minv_passes(float *s, int tid) {
    // backward pass, children before parents
    for (int i = N - 1; i >= 0; i--) {
        const int par = wt_parent[i], k = wt_S[i], nsub = wt_nsub[i];
        const Xf X = load_X(s, i);
        const float *IAi = s + L::IA + 36 * i;
        float U[6];
#pragma unroll
        for (int r = 0; r < 6; r++) U[r] = IAi[6 * r + k];
        const float Dinv = 1.0f / pick(U, k);
        for (int t = tid; t < nsub; t += NT - 32) {
            const int j = i + t;
            float Fij[6];
            const float *Fs = s + L::F + wt_fslot_b[i] * 6 * N + j;
#pragma unroll
            for (int r = 0; r < 6; r++) Fij[r] = (t == 0) ? 0.f : Fs[r * N];
            const float m = (t == 0 ? Dinv : 0.f) - Dinv * pick(Fij, k);
            s[L::Minv + i * N + j] = m;
            if (par >= 0) {
                float o[6];
#pragma unroll
                for (int r = 0; r < 6; r++) Fij[r] = fmaf(U[r], m, Fij[r]);
                xtforce(X, Fij, o);
                float *Fp = s + L::F + wt_fslot_b[par] * 6 * N + j;
#pragma unroll
                for (int r = 0; r < 6; r++) Fp[r * N] = o[r];
            }
            if (t == 0) {
                float w[6];
                xtforce(X, U, w);
                s[L::Dinv + i] = Dinv;
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    s[L::U + 6 * i + r] = U[r];
                    s[L::W + 6 * i + r] = w[r] * Dinv;
                }
            }
        }
        // articulated inertia: IA[par] += X^T (IA - U Dinv U^T) X, one lane per column
        const int cl = tid - WT::IA_LANE0;
        if (par >= 0 && cl >= 0 && cl < 6) {
            float e[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, x[6], t[6], o[6];
            add_at(e, cl, 1.0f);
            xmotion(X, e, x);
            float ux = 0.f;
#pragma unroll
            for (int r = 0; r < 6; r++) ux = fmaf(U[r], x[r], ux);
            ux *= Dinv;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                float acc = -U[r] * ux;
#pragma unroll
                for (int c = 0; c < 6; c++) acc = fmaf(IAi[6 * r + c], x[c], acc);
                t[r] = acc;
            }
            xtforce(X, t, o);
            float *IAp = s + L::IA + 36 * par;
#pragma unroll
            for (int r = 0; r < 6; r++) IAp[6 * r + cl] += o[r];
        }
        minv_bar();
    }
    // forward pass, serial over joints ("CANNOT BE IN PARALLEL BY BFS_LEVEL", _test.py:191)
    for (int i = 0; i < N; i++) {
        const int par = wt_parent[i], k = wt_S[i];
        const Xf X = load_X(s, i);
        float w[6];
        load6(s + L::W + 6 * i, w);
        for (int t = tid; t < N - i; t += NT - 32) {
            const int j = i + t;
            float m = s[L::Minv + i * N + j];
            float Fp[6], Fij[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (par >= 0) {
                const float *Fs = s + L::F + wt_fslot_f[par] * 6 * N + j;
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    Fp[r] = Fs[r * N];
                    m = fmaf(-w[r], Fp[r], m);
                }
                s[L::Minv + i * N + j] = m;
            }
            if (wt_fslot_f[i] >= 0) {
                if (par >= 0) xmotion(X, Fp, Fij);
                add_at(Fij, k, m);
                float *Fd = s + L::F + wt_fslot_f[i] * 6 * N + j;
#pragma unroll
                for (int r = 0; r < 6; r++) Fd[r * N] = Fij[r];
            }
        }
        minv_bar();
    }
}

__device__ __forceinline__ float minv_sym(const float *s, int r, int c) {
    return r <= c ? s[L::Minv + r * N + c] : s[L::Minv + c * N + r];
}

// ---- ID gradient: one lane per du-column, no barriers inside ---------------------------------
// replaces inverse_dynamics_gradient_inner (algorithms/_inverse_dynamics_gradient.py:27-650;
// oracle _test.py:229-488).  Lane ct owns column (j, sd) = (ct / 2, ct % 2): d/dq_j or d/dqd_j.  Joint-major
// lane order keeps the columns of one warp inside a contiguous range of (DFS-numbered) joints, so a
// warp skips every joint step outside the union of its subtrees (Atlas: 32 instead of 60 warp-steps).
__device__ void grad_column(float *s, int ct, bool valid) {
    // lanes without a column (valid == false) only take part in the warp votes
    const int sd = ct & 1;
    const int j = valid ? (ct >> 1) : N;
    const int col = sd * N + (valid ? j : 0);             // column of dc_du / df_du
    const int jend = valid ? j + wt_nsub[j] : j;
    const int lj = valid ? wt_level[j] : 0;
    float *dfp = s + L::df + sd * WT::DF_WORDS;
    float dv[6], da[6];
    // forward: joints of subtree(j) in id order, all lanes of the warp on the same joint
    for (int i = 0; i < N; i++) {
        const bool active = (i >= j) && (i < jend);
        if (!__any_sync(0xffffffffu, active)) continue;
        if (!active) continue;
        const Xf X = load_X(s, i);
        const int k = wt_S[i], par = wt_parent[i];
        const float qdi = s[L::qd + i];
        float vi[6], t[6];
        load6(s + L::v + 6 * i, vi);
        if (i == j) {
            if (sd == 0) {
                float xa[6];
                load6(s + L::Xa + 6 * i, xa);
                mxS(k, vi, dv);                       // == mxS(X v_parent)
                mxS(k, dv, da);
                mxS(k, xa, t);
#pragma unroll
                for (int r = 0; r < 6; r++) da[r] = fmaf(da[r], qdi, t[r]);
            } else {
#pragma unroll
                for (int r = 0; r < 6; r++) dv[r] = 0.f;
                add_at(dv, k, 1.0f);
                mxS(k, vi, da);
            }
        } else {
            if (par != i - 1) {                       // first joint of a later branch: reload the parent's dv, da
                const float *p = s + L::sv + (wt_saveslot[par] * 12) * (2 * N) + ct;
#pragma unroll
                for (int r = 0; r < 6; r++) { dv[r] = p[r * 2 * N]; da[r] = p[(6 + r) * 2 * N]; }
            }
            float dvn[6], dan[6];
            xmotion(X, dv, dvn);
            xmotion(X, da, dan);
            mxS(k, dvn, t);
#pragma unroll
            for (int r = 0; r < 6; r++) { dv[r] = dvn[r]; da[r] = fmaf(t[r], qdi, dan[r]); }
        }
        if (wt_saveslot[i] >= 0) {
            float *p = s + L::sv + (wt_saveslot[i] * 12) * (2 * N) + ct;
#pragma unroll
            for (int r = 0; r < 6; r++) { p[r * 2 * N] = dv[r]; p[(6 + r) * 2 * N] = da[r]; }
        }
        // df = I da + dv x* (I v) + v x* (I dv)
        float ivv[6], df[6], idv[6];
        load6(s + L::Iv + 6 * i, ivv);
        imul(i, da, df);
        imul(i, dv, idv);
        crossf(dv, ivv, t);
#pragma unroll
        for (int r = 0; r < 6; r++) df[r] += t[r];
        crossf(vi, idv, t);
        const int w = wt_level[i] + 1;
        float *d = dfp + wt_dfbase[i] + lj;
#pragma unroll
        for (int r = 0; r < 6; r++) d[r * w] = df[r] + t[r];
    }
    // backward: subtree accumulation (i > j), leave the subtree at i == j, then up the ancestors
    float up[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int i = N - 1; i >= 0; i--) {
        const bool in_sub = (i >= j) && (i < jend);
        const bool is_anc = valid && (i < j) && (j < i + wt_nsub[i]);
        const bool active = in_sub || is_anc;
        if (!__any_sync(0xffffffffu, active)) continue;
        if (!active) continue;
        const int k = wt_S[i], par = wt_parent[i];
        float vec[6];
        if (in_sub) {
            const int w = wt_level[i] + 1;
            const float *d = dfp + wt_dfbase[i] + lj;
#pragma unroll
            for (int r = 0; r < 6; r++) vec[r] = d[r * w];
        } else {
#pragma unroll
            for (int r = 0; r < 6; r++) vec[r] = up[r];
        }
        float dcv = pick(vec, k);
        if (sd == 1 && i == j) dcv += wt_damping[i];
        s[L::dc + col * N + i] = dcv;
        if (par >= 0) {
            if (i == j && sd == 0) {
                float fi[6], t[6];
                load6(s + L::f + 6 * i, fi);
                mxS(k, fi, t);
#pragma unroll
                for (int r = 0; r < 6; r++) vec[r] -= t[r];
            }
            const Xf X = load_X(s, i);
            float o[6];
            xtforce(X, vec, o);
            if (i > j) {
                const int wp = wt_level[par] + 1;
                float *d = dfp + wt_dfbase[par] + lj;
#pragma unroll
                for (int r = 0; r < 6; r++) d[r * wp] += o[r];
            } else {
#pragma unroll
                for (int r = 0; r < 6; r++) up[r] = o[r];
            }
        }
    }
}

// ---- df_du = -Minv dc_du on the tensor cores (3xTF32), measured variant ----------------------------
// BASELINE.json north_star: "tensor cores only if a batched product in the gradient path measurably wins at
// the stated tolerance".  The product is (N x N) (N x 2N) per state.  mma.sync.m16n8k8 TF32 with the 3xTF32
// split (hi*hi + hi*lo + lo*hi) keeps FP32-class accuracy (1.4e-6 relative at N = 64, 5e-4 for one TF32
// pass - too close to the 1e-3 bar; profiles/r2_micro_tc_minv_gemm.jsonl).  Operands are first copied into
// row-padded staging (pitch N + 4: conflict-free fragment loads) in the df/sv region, which is dead once
// grad_column has run.  Warp w < N/16 owns rows [16w, 16w + 16).  Selected by WT::TC_MATMUL
// (KernelPlan(wps_tc_matmul=True)); needs N % 16 == 0.
__device__ __forceinline__ unsigned to_tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void mma_m16n8k8_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NN>
__device__ __forceinline__ void matmul_tc3(float *s) {
    constexpr int P = NN + 4, MT = NN / 16, NTL = 2 * NN / 8;
    static_assert(NN % 16 == 0, "tensor-core product needs N % 16 == 0");
    static_assert(3 * NN * P <= 2 * WT::DF_WORDS + WT::NSAVE * 12 * 2 * NN, "staging does not fit the dead df/sv region");
    static_assert(MT * 32 <= NT, "not enough warps");
    float *pA = s + L::df, *pB = pA + NN * P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    for (int e = tid; e < NN * NN; e += NT) pA[(e / NN) * P + e % NN] = s[L::Minv + e];
    for (int e = tid; e < 2 * NN * NN; e += NT) pB[(e / NN) * P + e % NN] = s[L::dc + e];
    __syncthreads();
    if (warp < MT) {
        float acc[NTL][4];
#pragma unroll
        for (int j = 0; j < NTL; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
#pragma unroll
        for (int k0 = 0; k0 < NN; k0 += 8) {
            const float af[4] = {pA[(16 * warp + g) * P + k0 + t], pA[(16 * warp + g + 8) * P + k0 + t],
                                 pA[(16 * warp + g) * P + k0 + t + 4], pA[(16 * warp + g + 8) * P + k0 + t + 4]};
            unsigned ah[4], al[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                ah[i] = to_tf32(af[i]);
                al[i] = to_tf32(af[i] - __uint_as_float(ah[i]));
            }
#pragma unroll
            for (int j = 0; j < NTL; j++) {
                const float bf[2] = {pB[(8 * j + g) * P + k0 + t], pB[(8 * j + g) * P + k0 + t + 4]};
                const unsigned bh[2] = {to_tf32(bf[0]), to_tf32(bf[1])};
                const unsigned bl[2] = {to_tf32(bf[0] - __uint_as_float(bh[0])), to_tf32(bf[1] - __uint_as_float(bh[1]))};
                mma_m16n8k8_tf32(acc[j], al, bh);                // small terms first
                mma_m16n8k8_tf32(acc[j], ah, bl);
                mma_m16n8k8_tf32(acc[j], ah, bh);
            }
        }
#pragma unroll
        for (int j = 0; j < NTL; j++) {
            const int c0 = 8 * j + 2 * t, r0 = 16 * warp + g;
            s[L::dc + c0 * NN + r0] = -acc[j][0];
            s[L::dc + (c0 + 1) * NN + r0] = -acc[j][1];
            s[L::dc + c0 * NN + r0 + 8] = -acc[j][2];
            s[L::dc + (c0 + 1) * NN + r0 + 8] = -acc[j][3];
        }
    }
    __syncthreads();
}

// ALG: 0 = Minv, 1 = FD, 2 = ID gradient, 3 = FD gradient.
// EXTRA: ALG 2 -> qdd given (USE_QDD_FLAG); ALG 3 -> qdd and Minv given (USE_QDD_MINV_FLAG).
//
// wps_compute: one state, everything in the shared-memory block `s` (struct L).  On entry the inputs
// are in place - s[L::q..] (q | qd | u as far as the algorithm reads them), for EXTRA s[L::qdd..] and,
// ALG 3, the full symmetric Minv in s[L::Minv..] - and s[L::Ic..] holds the link inertias.  On exit:
//   ALG 0: s[L::Minv + row * N + col], upper triangle (row <= col) valid
//   ALG 1: s[L::qdd + i]
//   ALG 2: s[L::dc + col * N + row] = dc_du       ALG 3: the same block holds df_du
// All NT threads of the CTA must call; ends with a barrier.
template <int ALG, bool EXTRA>
__device__ __forceinline__ void wps_compute(float *s, float gravity) {
    const int tid = threadIdx.x;
    constexpr bool need_minv = (ALG == 0) || (ALG == 1) || (ALG == 3 && !EXTRA);
    constexpr bool need_c0 = (ALG == 1) || (ALG == 3 && !EXTRA);
    constexpr int RW = WT::RNEA_TID;                 // first thread of the RNEA warp
    for (int i = tid; i < N; i += NT) update_X(s, i);
    if (need_minv) {
        for (int e = tid; e < 36 * N; e += NT) s[L::IA + e] = s[L::Ic + e];
        for (int e = tid; e < N * N; e += NT) s[L::Minv + e] = 0.f;
    }
    __syncthreads();
    // Minv passes on warps [0, RW/32), bias forces (RNEA with qdd = 0) concurrently on the RNEA warp
    if (tid < RW) {
        if (need_minv) minv_passes(s, tid);
    } else {
        if (need_c0) rnea_rows(s, tid - RW, false, gravity);
    }
    __syncthreads();
    if (need_minv && ALG != 0) {                 // mirror the upper triangle: later reads are plain loads
        for (int e = tid; e < N * N; e += NT) {
            const int rr = e / N, cc = e - rr * N;
            if (rr > cc) s[L::Minv + e] = s[L::Minv + cc * N + rr];
        }
        __syncthreads();
    }
    if (need_c0) {        // forward_dynamics_finish (algorithms/_forward_dynamics.py:21-49)
        for (int r = tid; r < N; r += NT) {
            float acc = 0.f;
            for (int k = 0; k < N; k++) acc = fmaf(s[L::Minv + r * N + k], s[L::u + k] - s[L::c + k], acc);
            s[L::qdd + r] = acc;
        }
        __syncthreads();
    }
    if (ALG >= 2) {
        if (tid >= RW) rnea_rows(s, tid - RW, ALG == 3 ? true : EXTRA, gravity);
        else
            for (int e = tid; e < 2 * N * N; e += RW) s[L::dc + e] = 0.f;
        __syncthreads();
        if (tid < WT::COL_WARPS * 32)            // whole warps: grad_column votes with a full mask
            grad_column(s, tid < 2 * N ? tid : 2 * N - 1, tid < 2 * N);
        __syncthreads();
        if constexpr (ALG == 3 && WT::TC_MATMUL && N % 16 == 0) {
            matmul_tc3<N>(s);
        } else if (ALG == 3) {
            // df_du[:, col] = -Minv dc_du[:, col]  (algorithms/_forward_dynamics_gradient.py:48-57)
            // each lane owns one column: read it into registers, overwrite it in place
            for (int col = tid; col < 2 * N; col += NT) {
                float dcol[N];
#pragma unroll
                for (int k = 0; k < N; k++) dcol[k] = s[L::dc + col * N + k];
                for (int r = 0; r < N; r++) {
                    float acc = 0.f;
#pragma unroll
                    for (int k = 0; k < N; k++) acc = fmaf(s[L::Minv + r * N + k], dcol[k], acc);
                    s[L::dc + col * N + r] = -acc;
                }
            }
            __syncthreads();
        }
    }
}

template <int ALG, bool EXTRA>
__device__ __forceinline__ void wps_body(float *__restrict__ d_out, const float *__restrict__ d_in, int stride,
                                         const float *__restrict__ d_qdd, const float *__restrict__ d_Minv,
                                         int num_states, float gravity) {
    extern __shared__ float4 smem4[];
    float *s = reinterpret_cast<float *>(smem4);
    const int tid = threadIdx.x;
    constexpr int n_in = (ALG == 0) ? N : ((ALG == 1 || (ALG == 3 && !EXTRA)) ? 3 * N : 2 * N);

    for (int e = tid; e < 36 * N; e += NT) s[L::Ic + e] = __ldg(wt_I_g + e);       // once per CTA
    for (long long st = blockIdx.x; st < num_states; st += gridDim.x) {
        for (int e = tid; e < n_in; e += NT) s[L::q + e] = __ldg(d_in + st * stride + e);
        if (EXTRA) {
            for (int e = tid; e < N; e += NT) s[L::qdd + e] = __ldg(d_qdd + st * N + e);
            if (ALG == 3)
                for (int e = tid; e < N * N; e += NT) {
                    const int cc = e / N, rr = e - cc * N;      // global is column-major, upper triangle
                    if (rr <= cc) {
                        const float m = __ldg(d_Minv + st * N * N + e);
                        s[L::Minv + rr * N + cc] = m;
                        s[L::Minv + cc * N + rr] = m;
                    }
                }
        }
        __syncthreads();
        wps_compute<ALG, EXTRA>(s, gravity);
        if (ALG == 0) {
            float *o = d_out + st * N * N;
            for (int e = tid; e < N * N; e += NT) {
                const int cc = e / N, rr = e - cc * N;
                o[e] = rr <= cc ? s[L::Minv + rr * N + cc] : 0.f;
            }
        } else if (ALG == 1) {
            for (int r = tid; r < N; r += NT) d_out[st * N + r] = s[L::qdd + r];
        } else {
            float *o = d_out + st * 2 * N * N;
            for (int e = tid; e < 2 * N * N; e += NT) o[e] = s[L::dc + e];
        }
        __syncthreads();
    }
}

// The reference's *_inner contract for robots whose single-thread program is too large (Atlas, 64-link
// chain): inputs and output already in SHARED memory, scratch supplied by the caller
// (GRiDCodeGenerator.py:249-276).  s_work = L::total floats of 16-byte aligned shared memory (what the
// facade reports as gen_<alg>_inner_temp_mem_size()); blockDim.x must be NT.  s_x = u (ALG 1, 3 without
// EXTRA) or qdd (EXTRA) or unused; s_Minv_in = column-major n x n, upper triangle read (ALG 3 with EXTRA).
// Outputs in the reference's layouts: Minv column-major upper (strict lower 0), qdd[n], dc_du / df_du
// column-major n x 2n.
template <int ALG, bool EXTRA>
__device__ __forceinline__ void wps_inner(float *s_out, const float *s_q, const float *s_qd, const float *s_x,
                                          const float *s_Minv_in, float *s_work, float gravity) {
    float *s = s_work;
    const int tid = threadIdx.x;
    __syncthreads();                                 // the caller's writes to its inputs are visible
    for (int e = tid; e < 36 * N; e += NT) s[L::Ic + e] = __ldg(wt_I_g + e);
    for (int e = tid; e < N; e += NT) {
        s[L::q + e] = s_q[e];
        if (ALG != 0) s[L::qd + e] = s_qd[e];
        if (ALG == 1 || (ALG == 3 && !EXTRA)) s[L::u + e] = s_x[e];
        if (EXTRA) s[L::qdd + e] = s_x[e];
    }
    if (ALG == 3 && EXTRA)
        for (int e = tid; e < N * N; e += NT) {
            const int cc = e / N, rr = e - cc * N;
            if (rr <= cc) {
                const float m = s_Minv_in[e];
                s[L::Minv + rr * N + cc] = m;
                s[L::Minv + cc * N + rr] = m;
            }
        }
    __syncthreads();
    wps_compute<ALG, EXTRA>(s, gravity);
    if (ALG == 0) {
        for (int e = tid; e < N * N; e += NT) {
            const int cc = e / N, rr = e - cc * N;
            s_out[e] = rr <= cc ? s[L::Minv + rr * N + cc] : 0.f;
        }
    } else if (ALG == 1) {
        for (int r = tid; r < N; r += NT) s_out[r] = s[L::qdd + r];
    } else {
        for (int e = tid; e < 2 * N * N; e += NT) s_out[e] = s[L::dc + e];
    }
    __syncthreads();
}

// inverse_dynamics_gradient_inner of the reference takes the RNEA results (v, a, f of every joint,
// s_vaf = [v(6n) | a(6n) | f(6n)]) instead of computing them (algorithms/_inverse_dynamics_gradient.py:
// 27-41).  X a_parent is recovered from a_i: mxS(X a_p) = mxS(a_i - mxS(v_i) qd_i) because mxS_k(e_k) = 0.
__device__ __forceinline__ void wps_grad_inner_vaf(float *s_out, const float *s_q, const float *s_qd,
                                                   const float *s_vaf, float *s_work) {
    float *s = s_work;
    const int tid = threadIdx.x;
    __syncthreads();
    for (int e = tid; e < 36 * N; e += NT) s[L::Ic + e] = __ldg(wt_I_g + e);
    for (int e = tid; e < N; e += NT) { s[L::q + e] = s_q[e]; s[L::qd + e] = s_qd[e]; }
    for (int e = tid; e < 6 * N; e += NT) { s[L::v + e] = s_vaf[e]; s[L::f + e] = s_vaf[12 * N + e]; }
    __syncthreads();
    for (int i = tid; i < N; i += NT) {
        update_X(s, i);
        const int k = wt_S[i];
        float v6[6], a6[6], t[6], iv[6];
        load6(s + L::v + 6 * i, v6);
        load6(s_vaf + 6 * N + 6 * i, a6);
        if (wt_parent[i] >= 0) {
            mxS(k, v6, t);
            const float qdi = s[L::qd + i];
#pragma unroll
            for (int r = 0; r < 6; r++) a6[r] -= t[r] * qdi;
        }
        store6(s + L::Xa + 6 * i, a6);
#pragma unroll
        for (int r = 0; r < 6; r++) {
            float acc = 0.f;
#pragma unroll
            for (int c = 0; c < 6; c++) acc = fmaf(s[L::Ic + 36 * i + 6 * r + c], v6[c], acc);
            iv[r] = acc;
        }
        store6(s + L::Iv + 6 * i, iv);
    }
    for (int e = tid; e < 2 * N * N; e += NT) s[L::dc + e] = 0.f;
    __syncthreads();
    if (tid < WT::COL_WARPS * 32) grad_column(s, tid < 2 * N ? tid : 2 * N - 1, tid < 2 * N);
    __syncthreads();
    for (int e = tid; e < 2 * N * N; e += NT) s_out[e] = s[L::dc + e];
    __syncthreads();
}

}}  // namespace GRID_NS::wps

namespace GRID_NS { namespace wps {

// needs blockDim.x == WT::NT threads and L::total floats of dynamic shared memory
template <int ALG, bool EXTRA>
__global__ void __launch_bounds__(NT)
wps_kernel(float *__restrict__ d_out, const float *__restrict__ d_in, int stride, const float *__restrict__ d_qdd,
           const float *__restrict__ d_Minv, int num_states, float gravity) {
    wps_body<ALG, EXTRA>(d_out, d_in, stride, d_qdd, d_Minv, num_states, gravity);
}

}}  // namespace GRID_NS::wps

namespace GRID_NS { namespace wps {

template <int ALG, bool EXTRA>
cudaError_t wps_launch(float *d_out, const float *d_in, int stride, const float *d_qdd, const float *d_Minv,
                       int num_states, float gravity, cudaStream_t stream) {
    if (num_states <= 0) return cudaSuccess;
    auto kern = wps_kernel<ALG, EXTRA>;
    constexpr size_t smem_bytes = sizeof(float) * L::total;
    static int caps[kMaxDevices];
    int dev = 0;
    if (cudaError_t e = current_device(dev)) return e;
    int &cap = caps[dev];
    if (cap == 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
        int sms = 0, per_sm = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem_bytes);
        if (e != cudaSuccess) return e;
        if (per_sm < 1) return cudaErrorLaunchOutOfResources;
        cap = sms * per_sm;
    }
    const int blocks = num_states < cap ? num_states : cap;
    kern<<<blocks, NT, smem_bytes, stream>>>(d_out, d_in, stride, d_qdd, d_Minv, num_states, gravity);
    return cudaGetLastError();
}

}}  // namespace GRID_NS::wps
