/*
 * rbd_oracle.c - plain-C float64 restatement of the reference's numpy algorithms.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/rbd_numpy.py for the rules).  Same algorithms as the
 * reference's _test.py, written for any robot table so whole batches can be checked in seconds:
 *   orc_rnea       <- test_rnea_fpass 5-76, test_rnea_bpass 78-107
 *   orc_minv       <- test_minv_bpass 117-184, test_minv_fpass 186-202
 *   orc_rnea_grad  <- test_rnea_grad_inner 229-488
 *   fd / fd_grad   <- test_fd_grad 496-520
 * Dense 6x6 arithmetic throughout (no sparsity tricks): it is the checker, not the product.
 * Validated against oracle/rbd_numpy.py and the reference goldens in tests/test_oracle.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n;
    const int *parent;      /* n */
    const int *S;           /* n, 0..5 */
    const double *E0;       /* 9n row-major */
    const double *r0;       /* 3n */
    const double *I;        /* 36n row-major */
    const double *damping;  /* n */
} orc_robot;

static void xmat(const orc_robot *R, int i, double q, double *X /*36 row-major*/) {
    const double *E0 = R->E0 + 9 * i;
    double E[9], r[3] = {R->r0[3 * i], R->r0[3 * i + 1], R->r0[3 * i + 2]};
    int k = R->S[i];
    memcpy(E, E0, sizeof(E));
    if (k < 3) {
        int a = (k + 1) % 3, b = (k + 2) % 3;
        double c = cos(q), s = sin(q);
        for (int col = 0; col < 3; col++) {
            E[3 * a + col] = c * E0[3 * a + col] + s * E0[3 * b + col];
            E[3 * b + col] = c * E0[3 * b + col] - s * E0[3 * a + col];
        }
    } else {
        for (int t = 0; t < 3; t++) r[t] += q * E0[3 * (k - 3) + t];
    }
    double rx[9] = {0, -r[2], r[1], r[2], 0, -r[0], -r[1], r[0], 0};
    memset(X, 0, 36 * sizeof(double));
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
            X[6 * a + b] = E[3 * a + b];
            X[6 * (a + 3) + b + 3] = E[3 * a + b];
            double acc = 0;
            for (int t = 0; t < 3; t++) acc += E[3 * a + t] * rx[3 * t + b];
            X[6 * (a + 3) + b] = -acc;
        }
}
static void mv6(const double *M, const double *v, double *o) {   /* o = M v */
    for (int r = 0; r < 6; r++) {
        double acc = 0;
        for (int c = 0; c < 6; c++) acc += M[6 * r + c] * v[c];
        o[r] = acc;
    }
}
static void mtv6(const double *M, const double *v, double *o) {  /* o = M^T v */
    for (int c = 0; c < 6; c++) {
        double acc = 0;
        for (int r = 0; r < 6; r++) acc += M[6 * r + c] * v[r];
        o[c] = acc;
    }
}
static void cross3(const double *a, const double *b, double *o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
static void mxS(int k, const double *v, double *o) {     /* (v x) e_k  (_test.py:522-608) */
    double e[3] = {0, 0, 0};
    e[k % 3] = 1.0;
    if (k < 3) {
        cross3(v, e, o);
        cross3(v + 3, e, o + 3);
    } else {
        o[0] = o[1] = o[2] = 0;
        cross3(v, e, o + 3);
    }
}
static void crossf(const double *v, const double *f, double *o) {   /* v x* f (_test.py:649-664) */
    double t[3];
    cross3(v, f, o);
    cross3(v + 3, f + 3, t);
    o[0] += t[0]; o[1] += t[1]; o[2] += t[2];
    cross3(v, f + 3, o + 3);
}

/* workspace: X 36n | v a f Xa Iv 6n each */
static void orc_rnea(const orc_robot *R, const double *X, const double *qd, const double *qdd, double g,
                     double *c, double *v, double *a, double *f, double *Xa, double *Iv) {
    int n = R->n;
    for (int i = 0; i < n; i++) {
        int p = R->parent[i], k = R->S[i];
        double t[6], base[6] = {0, 0, 0, 0, 0, g};
        if (p < 0) {
            memset(v + 6 * i, 0, 6 * sizeof(double));
            mv6(X + 36 * i, base, Xa + 6 * i);
        } else {
            mv6(X + 36 * i, v + 6 * p, v + 6 * i);
            mv6(X + 36 * i, a + 6 * p, Xa + 6 * i);
        }
        v[6 * i + k] += qd[i];
        memcpy(a + 6 * i, Xa + 6 * i, 6 * sizeof(double));
        if (qdd) a[6 * i + k] += qdd[i];
        if (p >= 0) {
            mxS(k, v + 6 * i, t);
            for (int r = 0; r < 6; r++) a[6 * i + r] += t[r] * qd[i];
        }
        mv6(R->I + 36 * i, v + 6 * i, Iv + 6 * i);
        mv6(R->I + 36 * i, a + 6 * i, f + 6 * i);
        crossf(v + 6 * i, Iv + 6 * i, t);
        for (int r = 0; r < 6; r++) f[6 * i + r] += t[r];
    }
    for (int i = n - 1; i >= 0; i--) {
        int p = R->parent[i];
        c[i] = f[6 * i + R->S[i]] + R->damping[i] * qd[i];
        if (p >= 0) {
            double t[6];
            mtv6(X + 36 * i, f + 6 * i, t);
            for (int r = 0; r < 6; r++) f[6 * p + r] += t[r];
        }
    }
}

static int in_sub(const orc_robot *R, int j, int i) {    /* j in subtree(i) */
    while (j >= 0) {
        if (j == i) return 1;
        j = R->parent[j];
    }
    return 0;
}

/* Minv: n*n row-major [row][col], upper triangle.  ws: F 6n*n*n? -> F[i][r][j] (n*6*n), IA 36n, U 6n, Dinv n */
static void orc_minv(const orc_robot *R, const double *X, double *Minv, double *F, double *IA, double *U, double *Dinv) {
    int n = R->n;
    memset(Minv, 0, sizeof(double) * n * n);
    memset(F, 0, sizeof(double) * n * 6 * n);
    memcpy(IA, R->I, sizeof(double) * 36 * n);
    for (int i = n - 1; i >= 0; i--) {
        int p = R->parent[i], k = R->S[i];
        double *Fi = F + (size_t)i * 6 * n;
        for (int r = 0; r < 6; r++) U[6 * i + r] = IA[36 * i + 6 * r + k];
        Dinv[i] = 1.0 / U[6 * i + k];
        Minv[i * n + i] = Dinv[i];
        for (int j = i; j < n; j++) {
            if (!in_sub(R, j, i)) continue;
            Minv[i * n + j] -= Dinv[i] * Fi[k * n + j];
            if (p >= 0) {
                double col[6], t[6];
                for (int r = 0; r < 6; r++) {
                    Fi[r * n + j] += U[6 * i + r] * Minv[i * n + j];
                    col[r] = Fi[r * n + j];
                }
                mtv6(X + 36 * i, col, t);
                for (int r = 0; r < 6; r++) F[(size_t)p * 6 * n + r * n + j] += t[r];
            }
        }
        if (p >= 0) {
            double Ia[36], T[36];
            for (int r = 0; r < 6; r++)
                for (int c = 0; c < 6; c++) Ia[6 * r + c] = IA[36 * i + 6 * r + c] - U[6 * i + r] * U[6 * i + c] * Dinv[i];
            for (int r = 0; r < 6; r++)         /* T = Ia X */
                for (int c = 0; c < 6; c++) {
                    double acc = 0;
                    for (int t = 0; t < 6; t++) acc += Ia[6 * r + t] * X[36 * i + 6 * t + c];
                    T[6 * r + c] = acc;
                }
            for (int r = 0; r < 6; r++)         /* IA[p] += X^T T */
                for (int c = 0; c < 6; c++) {
                    double acc = 0;
                    for (int t = 0; t < 6; t++) acc += X[36 * i + 6 * t + r] * T[6 * t + c];
                    IA[36 * p + 6 * r + c] += acc;
                }
        }
    }
    for (int i = 0; i < n; i++) {
        int p = R->parent[i], k = R->S[i];
        double *Fi = F + (size_t)i * 6 * n;
        double w[6];
        if (p >= 0) mtv6(X + 36 * i, U + 6 * i, w);
        for (int j = i; j < n; j++) {
            double col[6] = {0, 0, 0, 0, 0, 0}, t[6] = {0, 0, 0, 0, 0, 0};
            if (p >= 0) {
                double acc = 0;
                for (int r = 0; r < 6; r++) {
                    col[r] = F[(size_t)p * 6 * n + r * n + j];
                    acc += w[r] * col[r];
                }
                Minv[i * n + j] -= Dinv[i] * acc;
                mv6(X + 36 * i, col, t);
            }
            for (int r = 0; r < 6; r++) Fi[r * n + j] = t[r];
            Fi[k * n + j] += Minv[i * n + j];
        }
    }
}

/* dc: n x 2n row-major.  ws: dv, da, df each 2*n*6*n : [s][i][r][col] */
static void orc_rnea_grad(const orc_robot *R, const double *X, const double *qd, const double *v, const double *f,
                          const double *Xa, const double *Iv, double *dc, double *dv, double *da, double *df) {
    int n = R->n;
    size_t blk = (size_t)n * 6 * n;
    memset(dv, 0, 2 * blk * sizeof(double));
    memset(da, 0, 2 * blk * sizeof(double));
    memset(df, 0, 2 * blk * sizeof(double));
    memset(dc, 0, sizeof(double) * n * 2 * n);
    for (int i = 0; i < n; i++) {
        int p = R->parent[i], k = R->S[i];
        for (int s = 0; s < 2; s++) {
            double *dvi = dv + s * blk + (size_t)i * 6 * n, *dai = da + s * blk + (size_t)i * 6 * n;
            double *dfi = df + s * blk + (size_t)i * 6 * n;
            for (int col = 0; col <= i; col++) {
                if (!in_sub(R, i, col)) continue;          /* col in anc(i) or col == i */
                double a6[6] = {0, 0, 0, 0, 0, 0}, b6[6] = {0, 0, 0, 0, 0, 0}, x[6], t[6], u6[6];
                if (p >= 0 && col != i) {
                    for (int r = 0; r < 6; r++) {
                        x[r] = dv[s * blk + (size_t)p * 6 * n + r * n + col];
                        u6[r] = da[s * blk + (size_t)p * 6 * n + r * n + col];
                    }
                    mv6(X + 36 * i, x, a6);
                    mv6(X + 36 * i, u6, b6);
                }
                if (col == i) {
                    if (s == 0) {
                        mxS(k, v + 6 * i, t);
                        for (int r = 0; r < 6; r++) a6[r] += t[r];
                        mxS(k, Xa + 6 * i, t);
                        for (int r = 0; r < 6; r++) b6[r] += t[r];
                    } else {
                        a6[k] += 1.0;
                        mxS(k, v + 6 * i, t);
                        for (int r = 0; r < 6; r++) b6[r] += t[r];
                    }
                }
                mxS(k, a6, t);
                for (int r = 0; r < 6; r++) b6[r] += t[r] * qd[i];
                double Ida[6], Idv[6], c1[6], c2[6];
                mv6(R->I + 36 * i, b6, Ida);
                mv6(R->I + 36 * i, a6, Idv);
                crossf(a6, Iv + 6 * i, c1);
                crossf(v + 6 * i, Idv, c2);
                for (int r = 0; r < 6; r++) {
                    dvi[r * n + col] = a6[r];
                    dai[r * n + col] = b6[r];
                    dfi[r * n + col] = Ida[r] + c1[r] + c2[r];
                }
            }
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        int p = R->parent[i], k = R->S[i];
        for (int s = 0; s < 2; s++) {
            double *dfi = df + s * blk + (size_t)i * 6 * n;
            for (int col = 0; col < n; col++) {
                if (!(in_sub(R, i, col) || in_sub(R, col, i))) continue;
                dc[i * 2 * n + s * n + col] = dfi[k * n + col] + ((s == 1 && col == i) ? R->damping[i] : 0.0);
                if (p >= 0) {
                    double x[6], t[6];
                    for (int r = 0; r < 6; r++) x[r] = dfi[r * n + col];
                    if (s == 0 && col == i) {
                        mxS(k, f + 6 * i, t);
                        for (int r = 0; r < 6; r++) x[r] -= t[r];
                    }
                    mtv6(X + 36 * i, x, t);
                    for (int r = 0; r < 6; r++) df[s * blk + (size_t)p * 6 * n + r * n + col] += t[r];
                }
            }
        }
    }
}

typedef struct {
    const orc_robot *R;
    int alg, first, last;
    const double *q, *qd, *x;
    double g;
    double *out;
} job_t;

/* alg: 0 id (x = qdd or NULL), 1 minv (upper, col-major), 2 fd (x = u), 3 id_grad (x = qdd or NULL), 4 fd_grad (x = u) */
static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    const orc_robot *R = J->R;
    int n = R->n;
    size_t blk = (size_t)n * 6 * n;
    double *ws = (double *)malloc(sizeof(double) * (36 * n + 31 * n + 2 * (size_t)n * n + blk + 36 * n + 7 * n + 6 * blk +
                                                   2 * (size_t)n * n + 3 * n));
    double *X = ws, *v = X + 36 * n, *a = v + 6 * n, *f = a + 6 * n, *Xa = f + 6 * n, *Iv = Xa + 6 * n, *c = Iv + 6 * n;
    double *Minv = c + n, *Md = Minv + (size_t)n * n, *F = Md + (size_t)n * n, *IA = F + blk, *U = IA + 36 * n, *Dinv = U + 6 * n;
    double *dv = Dinv + n, *da = dv + 2 * blk, *df = da + 2 * blk, *dc = df + 2 * blk, *qdd = dc + 2 * (size_t)n * n, *tmp = qdd + n;
    for (int st = J->first; st < J->last; st++) {
        const double *q = J->q + (size_t)st * n, *qd = J->qd ? J->qd + (size_t)st * n : NULL;
        const double *x = J->x ? J->x + (size_t)st * n : NULL;
        for (int i = 0; i < n; i++) xmat(R, i, q[i], X + 36 * i);
        if (J->alg == 0) {
            orc_rnea(R, X, qd, x, J->g, J->out + (size_t)st * n, v, a, f, Xa, Iv);
        } else if (J->alg == 1) {
            orc_minv(R, X, Minv, F, IA, U, Dinv);
            double *o = J->out + (size_t)st * n * n;
            for (int col = 0; col < n; col++)
                for (int row = 0; row < n; row++) o[col * n + row] = row <= col ? Minv[row * n + col] : 0.0;
        } else if (J->alg == 3) {
            orc_rnea(R, X, qd, x, J->g, c, v, a, f, Xa, Iv);
            orc_rnea_grad(R, X, qd, v, f, Xa, Iv, dc, dv, da, df);
            double *o = J->out + (size_t)st * 2 * n * n;
            for (int col = 0; col < 2 * n; col++)
                for (int row = 0; row < n; row++) o[col * n + row] = dc[row * 2 * n + col];
        } else {
            orc_rnea(R, X, qd, NULL, J->g, c, v, a, f, Xa, Iv);
            orc_minv(R, X, Minv, F, IA, U, Dinv);
            for (int r = 0; r < n; r++)
                for (int k = 0; k < n; k++) Md[r * n + k] = r <= k ? Minv[r * n + k] : Minv[k * n + r];
            for (int r = 0; r < n; r++) {
                double acc = 0;
                for (int k = 0; k < n; k++) acc += Md[r * n + k] * (x[k] - c[k]);
                qdd[r] = acc;
            }
            if (J->alg == 2) {
                memcpy(J->out + (size_t)st * n, qdd, n * sizeof(double));
                continue;
            }
            orc_rnea(R, X, qd, qdd, J->g, tmp, v, a, f, Xa, Iv);
            orc_rnea_grad(R, X, qd, v, f, Xa, Iv, dc, dv, da, df);
            double *o = J->out + (size_t)st * 2 * n * n;
            for (int col = 0; col < 2 * n; col++)
                for (int row = 0; row < n; row++) {
                    double acc = 0;
                    for (int k = 0; k < n; k++) acc += Md[row * n + k] * dc[k * 2 * n + col];
                    o[col * n + row] = -acc;
                }
        }
    }
    free(ws);
    return NULL;
}

int orc_batch(int n, const int *parent, const int *S, const double *E0, const double *r0, const double *I,
              const double *damping, int alg, int N, const double *q, const double *qd, const double *x, double gravity,
              double *out, int threads) {
    orc_robot R = {n, parent, S, E0, r0, I, damping};
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > N) threads = N > 0 ? N : 1;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < threads; t++) {
        jobs[t] = (job_t){&R, alg, (int)((long long)N * t / threads), (int)((long long)N * (t + 1) / threads), q, qd, x,
                          gravity, out};
        if (threads == 1) worker(&jobs[t]);
        else pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    if (threads > 1)
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    return 0;
}
